// ffa_kernels.h — internal launcher declarations shared by the kernel translation units and the C ABI.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "ffa_common.cuh"

namespace ffa {
struct FwdArgs;
struct BwdArgs;

cudaError_t launch_plan(bool fwd, const SliceMeta& meta, int n_tiles, int capacity, int* plan, cudaStream_t stream);
cudaError_t launch_fwd(bool bf16, bool out_f32, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                       const FwdArgs& a, cudaStream_t stream);
}  // namespace ffa
