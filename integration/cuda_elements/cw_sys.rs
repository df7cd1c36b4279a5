// cw_sys.rs — Rust FFI of include/circom_b200.h (the run-time side of the cuda_elements back end).
// Link: `cargo:rustc-link-lib=dylib=circom_b200`.  Every function returns CW_OK (0) or a negative code;
// `cw_last_error()` has the message (thread-local).
use std::os::raw::{c_char, c_float, c_int, c_void};

#[repr(C)] pub struct cw_circuit { _p: [u8; 0] }
#[repr(C)] pub struct cw_batch { _p: [u8; 0] }
#[repr(C)] pub struct cw_r1cs { _p: [u8; 0] }
#[repr(C)] pub struct cw_comm { _p: [u8; 0] }

pub const CW_FLAG_COMPACT: u32 = 16 | 32;

extern "C" {
    pub fn cw_last_error() -> *const c_char;
    pub fn cw_circuit_load(path: *const c_char, flags: u32, out: *mut *mut cw_circuit) -> c_int;
    pub fn cw_circuit_destroy(c: *mut cw_circuit);
    pub fn cw_get_size_of_witness(c: *const cw_circuit) -> u32;
    pub fn cw_get_main_input_signal_no(c: *const cw_circuit) -> u32;
    pub fn cw_fnv1a(name: *const c_char) -> u64;
    pub fn cw_circuit_write_dat(c: *const cw_circuit, path: *const c_char) -> c_int;
    pub fn cw_circuit_write_sym(c: *const cw_circuit, path: *const c_char) -> c_int;   // needs the symbols section of the .cb2c
    pub fn cw_batch_create(c: *const cw_circuit, batch: u32, device: c_int, out: *mut *mut cw_batch) -> c_int;
    pub fn cw_batch_destroy(b: *mut cw_batch);
    pub fn cw_batch_set_input(b: *mut cw_batch, instance: u32, name_hash: u64, idx: u32, limbs: *const u64) -> c_int;
    pub fn cw_batch_set_inputs(b: *mut cw_batch, inputs: *const u64, is_device_ptr: c_int) -> c_int;
    pub fn cw_batch_run(b: *mut cw_batch) -> c_int;
    pub fn cw_batch_sync(b: *mut cw_batch) -> c_int;
    pub fn cw_batch_status(b: *mut cw_batch, status: *mut i32) -> c_int;
    pub fn cw_batch_get_witness(b: *mut cw_batch, out: *mut u64) -> c_int;
    pub fn cw_batch_get_witness_async(b: *mut cw_batch, out: *mut u64) -> c_int;
    pub fn cw_batch_get_witness_wait(b: *mut cw_batch) -> c_int;
    pub fn cw_batch_write_wtns(b: *mut cw_batch, instance: u32, path: *const c_char) -> c_int;
    /// what the circuit's log() calls print for an instance (LogBucket), and the reference's failed-assert message
    pub fn cw_batch_log(b: *mut cw_batch, instance: u32, buf: *mut c_char, cap: usize, len: *mut usize) -> c_int;
    pub fn cw_circuit_assert_info(c: *const cw_circuit, assert_no: u32, buf: *mut c_char, cap: usize, len: *mut usize) -> c_int;
    pub fn cw_batch_stream(b: *mut cw_batch) -> *mut c_void;   // cudaStream_t
    pub fn cw_r1cs_from_circuit(c: *const cw_circuit, out: *mut *mut cw_r1cs) -> c_int;
    pub fn cw_r1cs_check_batch(r: *mut cw_r1cs, b: *mut cw_batch, first_bad: *mut i64, kernel_ms: *mut c_float) -> c_int;
    pub fn cw_r1cs_destroy(r: *mut cw_r1cs);
    pub fn cw_comm_unique_id(id: *mut u8) -> c_int;             // 128 bytes
    pub fn cw_comm_init(id: *const u8, rank: c_int, world: c_int, device: c_int, out: *mut *mut cw_comm) -> c_int;
    pub fn cw_comm_from_nccl(nccl_comm: *mut c_void, rank: c_int, world: c_int, device: c_int, out: *mut *mut cw_comm) -> c_int;
    pub fn cw_circuit_broadcast(cm: *mut cw_comm, c: *mut *mut cw_circuit, root: c_int) -> c_int;
    pub fn cw_batch_gather_witness_packed(cm: *mut cw_comm, b: *mut cw_batch, first: u32, count: u32, root: c_int,
                                          recv_device: *mut u32, send_scratch_device: *mut u32, ms: *mut c_float) -> c_int;
    pub fn cw_comm_destroy(c: *mut cw_comm);
}

/// calculateWitness for a slice of inputs (each: n_inputs canonical field elements, 4 x u64 limbs), dense rows out
pub fn calculate_witness_batch(cb2c: &std::ffi::CStr, inputs: &[u64], batch: u32, device: i32) -> Result<Vec<u64>, String> {
    unsafe {
        let err = || std::ffi::CStr::from_ptr(cw_last_error()).to_string_lossy().into_owned();
        let mut c = std::ptr::null_mut();
        if cw_circuit_load(cb2c.as_ptr(), CW_FLAG_COMPACT, &mut c) != 0 { return Err(err()); }
        let mut b = std::ptr::null_mut();
        if cw_batch_create(c, batch, device, &mut b) != 0 { cw_circuit_destroy(c); return Err(err()); }
        let w = cw_get_size_of_witness(c) as usize;
        let mut out = vec![0u64; batch as usize * w * 4];
        let rc = cw_batch_set_inputs(b, inputs.as_ptr(), 0) | cw_batch_run(b) | cw_batch_get_witness(b, out.as_mut_ptr());
        let mut st = vec![0i32; batch as usize];
        let rc = rc | cw_batch_status(b, st.as_mut_ptr());
        cw_batch_destroy(b);
        cw_circuit_destroy(c);
        if rc != 0 { return Err(err()); }
        if let Some(i) = st.iter().position(|s| *s != 0) { return Err(format!("instance {}: status {}", i, st[i])); }   // (cw_circuit_assert_info(c, st[i] - 1, ..) before the destroy calls gives the reference's message)
        Ok(out)
    }
}
