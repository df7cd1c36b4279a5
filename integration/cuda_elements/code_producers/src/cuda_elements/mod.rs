// code_producers/src/cuda_elements/mod.rs  (new module, beside c_elements / wasm_elements: lib.rs gains `pub mod cuda_elements;`)
//
// The CUDA producer does not print source text: it fills a `Cb2cFile` (docs/CB2C.md of circom_b200) that
// libcircom_b200.so lowers to an instruction tape.  Written against circom 2.2.3; NOT compiled in the build image of
// circom_b200 (no Rust toolchain there) - the byte layout is pinned by tests/test_cb2c_spec_cpu.py through an
// independent C writer, the semantics by the Python DSL that emits the same records.
use crate::components::{FieldMap, TemplateInstanceIOMap};
use num_bigint_dig::BigInt;
use std::collections::HashMap;
use std::io::Write;

pub use crate::components::*; // InputList, TemplateInstanceIOMap, ... (as c_elements/mod.rs:1)

/// What `CProducer` (c_elements/mod.rs:6-39) carries, reduced to what a .cb2c needs.
pub struct CUDAProducer {
    pub prime: String,                 // "--prime" value (program_structure/src/utils/constants.rs:3-13); unknown names: Err at `prime_id`
    pub prime_str: String,             // decimal modulus, as CProducer::get_prime()
    pub main_header: String,
    pub main_signal_offset: usize,     // = 1 (signal 0 is the constant one), c_elements/mod.rs:52
    pub number_of_main_outputs: usize,
    pub number_of_main_inputs: usize,
    pub main_input_list: InputList,    // (name, start, size): the source of the .dat hash map, mod.rs:17
    pub field_tracking: Vec<String>,   // the constant table the IR's ValueBucket{BigInt} indexes, mod.rs:27
    pub sanity_check_style: usize,     // 0: drop `===` asserts (assert_bucket.rs:73)
    pub function_ids: HashMap<String, u32>, // function header -> index in Cb2cFile::functions (Circuit::functions order)
    pub io_map: TemplateInstanceIOMap,      // as CProducer::io_map (c_elements/mod.rs:24): resolves LocationRule::Mapped now, travels to the .dat
    pub busid_field_info: FieldMap,         // as CProducer::busid_field_info: the fields of each bus (flattened by the producer)
    pub string_table: Vec<String>,          // as CProducer::get_string_table(): the literals of log()
}

impl CUDAProducer {
    pub fn prime_id(&self) -> Result<u32, ()> {
        match self.prime.as_str() {
            "bn128" => Ok(0),
            "bls12381" => Ok(1),
            "grumpkin" => Ok(2),
            "pallas" => Ok(3),
            "vesta" => Ok(4),
            "secq256r1" => Ok(5),
            "bls12377" => Ok(6),
            "goldilocks" => Ok(7),   // 64-bit values in the same 32-byte constants / elements
            _ => Err(()),
        }
    }
}

// ---- references (docs/CB2C.md "reference") -------------------------------------------------------------------------
#[derive(Clone, Copy, PartialEq, Eq, Hash, Debug)]
pub enum Ref {
    None,
    Imm(u32),                  // immediates of function code (NONE kind with an index)
    Own(u32),
    Sub { sub: u32, idx: u32 },
    Const(u32),
    Tmp(u32),
    One,
}
impl Ref {
    pub fn pack(self) -> u64 {
        let (k, s, i): (u64, u64, u64) = match self {
            Ref::None => (0, 0, 0),
            Ref::Imm(i) => (0, 0, i as u64),
            Ref::Own(i) => (1, 0, i as u64),
            Ref::Sub { sub, idx } => (2, sub as u64, idx as u64),
            Ref::Const(i) => (3, 0, i as u64),
            Ref::Tmp(i) => (4, 0, i as u64),
            Ref::One => (5, 0, 0),
        };
        (k << 56) | (s << 32) | i
    }
}

/// opcodes: OperatorType order (compute_bucket.rs:7-34) + the producer's own
#[allow(non_camel_case_types)]
#[derive(Clone, Copy)]
#[repr(u64)]
pub enum Op {
    MUL = 1, DIV, ADD, SUB, POW, IDIV, MOD, SHL, SHR, LEQ, GEQ, LT, GT, EQ, NEQ, LOR, LAND, LNOT, BOR, BAND, BXOR, BNOT,
    NEG = 23, COPY = 24, SELECT = 25, ASSERT = 26, ASSERT_EQ = 27, INV = 28, LOG = 29,
    JMP = 40, JZ = 41, RET = 42, LOADX = 43, STOREX = 44, CALL = 45, ARG = 46,
}

#[derive(Clone)]
pub struct OpRec { pub op: Op, pub d: Ref, pub a: Ref, pub b: Ref, pub c: Ref }

#[derive(Default)]
pub struct TemplateRecord {
    pub name: String,
    pub n_out: u32, pub n_in: u32, pub n_inter: u32, pub n_tmp: u32,
    pub subs: Vec<u32>,                                   // template index per sub-component, creation order
    pub ops: Vec<OpRec>,
    pub constraints: Vec<[Vec<(Ref, u32)>; 3]>,           // A, B, C: (signal reference, constant id)
    // symbols section (docs/CB2C.md): names of the own signals in numbering order (array elements spelled out, as
    // TemplateInstance::signals / dag::Node::signal_correspondence give them) and of the sub-components (the `symbol`
    // of each CreateCmpBucket, with its indices for component arrays) - what `--sym` prints (dag/src/sym_porting.rs)
    pub signal_names: Vec<String>,
    pub sub_names: Vec<String>,
}
#[derive(Default)]
pub struct FunctionRecord { pub name: String, pub n_params: u32, pub n_regs: u32, pub code: Vec<OpRec> }

#[derive(Default)]
pub struct Cb2cFile {
    pub prime: u32,
    pub consts: Vec<BigInt>,                              // canonical, < q
    const_index: HashMap<BigInt, u32>,
    pub templates: Vec<TemplateRecord>,
    pub main: u32,
    pub names: Vec<(String, u32, u32)>,                   // (qualified name, global signal id, size)
    pub functions: Vec<FunctionRecord>,
    pub io_map: TemplateInstanceIOMap,                    // copied from the producer by produce_cb2c: the IOMP section
    pub strings: Vec<String>,                             // the strings log() uses: the LOGS section
}

impl Cb2cFile {
    pub fn string_id(&mut self, text: &str) -> u32 {
        if let Some(p) = self.strings.iter().position(|s| s == text) { return p as u32; }
        self.strings.push(text.to_string());
        (self.strings.len() - 1) as u32
    }
    pub fn const_id(&mut self, v: &BigInt, q: &BigInt) -> u32 {
        let mut n = v % q;
        if n < BigInt::from(0) { n += q; }
        if let Some(i) = self.const_index.get(&n) { return *i; }
        let i = self.consts.len() as u32;
        self.const_index.insert(n.clone(), i);
        self.consts.push(n);
        i
    }

    fn w_u32<W: Write>(w: &mut W, v: u32) -> Result<(), ()> { w.write_all(&v.to_le_bytes()).map_err(|_| {}) }
    fn w_u64<W: Write>(w: &mut W, v: u64) -> Result<(), ()> { w.write_all(&v.to_le_bytes()).map_err(|_| {}) }
    fn w_str<W: Write>(w: &mut W, s: &str) -> Result<(), ()> {
        Self::w_u32(w, s.len() as u32)?;
        w.write_all(s.as_bytes()).map_err(|_| {})?;
        w.write_all(&[0u8; 3][..(4 - s.len() % 4) % 4]).map_err(|_| {})
    }
    fn w_ops<W: Write>(w: &mut W, ops: &[OpRec]) -> Result<(), ()> {
        for o in ops {
            Self::w_u64(w, o.op as u64)?;
            for r in [o.d, o.a, o.b, o.c] { Self::w_u64(w, r.pack())?; }
        }
        Ok(())
    }

    /// docs/CB2C.md "File layout", field by field
    pub fn write<W: Write>(&self, w: &mut W) -> Result<(), ()> {
        w.write_all(b"CB2C").map_err(|_| {})?;
        for v in [1u32, self.prime, self.consts.len() as u32, self.templates.len() as u32, self.main,
                  self.names.len() as u32, self.functions.len() as u32] { Self::w_u32(w, v)?; }
        for c in &self.consts {
            let (_, mut b) = c.to_bytes_le();
            b.resize(32, 0);
            w.write_all(&b).map_err(|_| {})?;
        }
        for t in &self.templates {
            Self::w_str(w, &t.name)?;
            let n_terms: usize = t.constraints.iter().map(|c| c[0].len() + c[1].len() + c[2].len()).sum();
            for v in [t.n_out, t.n_in, t.n_inter, t.subs.len() as u32, t.n_tmp, t.ops.len() as u32,
                      t.constraints.len() as u32, n_terms as u32] { Self::w_u32(w, v)?; }
            for s in &t.subs { Self::w_u32(w, *s)?; }
            Self::w_ops(w, &t.ops)?;
            for con in &t.constraints {
                for lc in con {
                    let mut terms = lc.clone();
                    terms.sort_by_key(|(r, _)| r.pack());          // canonical order
                    Self::w_u64(w, terms.len() as u64)?;
                    for (r, cid) in terms { Self::w_u64(w, r.pack())?; Self::w_u64(w, cid as u64)?; }
                }
            }
        }
        for (name, id, size) in &self.names { Self::w_str(w, name)?; Self::w_u32(w, *id)?; Self::w_u32(w, *size)?; }
        for f in &self.functions {
            Self::w_str(w, &f.name)?;
            for v in [f.n_params, f.n_regs, f.code.len() as u32] { Self::w_u32(w, v)?; }
            Self::w_ops(w, &f.code)?;
        }
        if !self.strings.is_empty() {                              // optional string table of log()
            w.write_all(b"LOGS").map_err(|_| {})?;
            Self::w_u32(w, self.strings.len() as u32)?;
            for s in &self.strings { Self::w_str(w, s)?; }
        }
        // optional io-map section (docs/CB2C.md): what generate_dat_io_signals_info puts into the .dat, handed to the library
        if !self.io_map.is_empty() {
            w.write_all(b"IOMP").map_err(|_| {})?;
            Self::w_u32(w, self.io_map.len() as u32)?;
            for (template_id, defs) in &self.io_map {              // BTreeMap: ascending template ids
                Self::w_u32(w, *template_id as u32)?;
                Self::w_u32(w, defs.len() as u32)?;
                for d in defs {
                    Self::w_u32(w, d.offset as u32)?;
                    Self::w_u32(w, d.lengths.len() as u32)?;
                    for l in &d.lengths { Self::w_u32(w, *l as u32)?; }
                    Self::w_u32(w, d.size as u32)?;
                    Self::w_u32(w, d.bus_id.unwrap_or(0) as u32)?;
                }
            }
        }
        // optional symbols section: only when every template carries a complete set of names
        if self.templates.iter().all(|t| t.signal_names.len() as u32 == t.n_out + t.n_in + t.n_inter
                                          && t.sub_names.len() == t.subs.len()) {
            w.write_all(b"SYMS").map_err(|_| {})?;
            for t in &self.templates {
                for n in &t.signal_names { Self::w_str(w, n)?; }
                for n in &t.sub_names { Self::w_str(w, n)?; }
            }
        }
        Ok(())
    }
}
