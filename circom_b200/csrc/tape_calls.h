// The interpreter builds that contain the function machine (tape_exec_kernel<.., HAS_CALLS = true, ..>) live in their
// own translation unit, tape_calls.cu: ptxas does not survive one module with every build of the interpreter.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace cw {
// this unit's copy of the constant field-parameter table (every .cu has its own c_fr without relocatable device code)
cudaError_t tape_calls_set_params(const FrParams *table, size_t bytes);
// prime: 0 / 1 select the specialised builds, anything else the generic-prime build (prime index from tp.prime)
void launch_tape_calls(int prime, const TapeDev &tp, uint4 *slots, u32 *plane, u32 bt_log2, u32 *first_assert, int *err,
                       u32 batch, u32 tiles, u32 threads, bool bit_plane, bool fused, cudaStream_t stream);
}  // namespace cw
