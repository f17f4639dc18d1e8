// ffa_kernels.h — internal launcher declarations shared by the kernel translation units and the C ABI.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "ffa_common.cuh"

namespace ffa {
struct FwdArgs;
struct BwdArgs;

cudaError_t launch_plan(bool fwd, const SliceMeta& meta, int n_tiles, int capacity, int* plan, cudaStream_t stream);
cudaError_t launch_plan_fwd_rows(const SliceMeta& meta, int n_tiles, int tile_rows, int capacity, int* plan, cudaStream_t stream);
cudaError_t launch_fwd(int head_dim, bool bf16, bool out_f32, const CUtensorMap& tmQ, const CUtensorMap& tmK,
                       const CUtensorMap& tmV, const FwdArgs& a, cudaStream_t stream);
cudaError_t launch_bwd_preprocess(int head_dim, bool bf16, const void* out, const void* dout, const float* lse, float* delta,
                                  float* lse2, int total_q, int num_heads, int tq_ld, int64_t o_stride_t,
                                  int64_t o_stride_h, int64_t do_stride_t, int64_t do_stride_h, int64_t lse_stride_t,
                                  int64_t lse_stride_h, cudaStream_t stream);
cudaError_t launch_bwd(int head_dim, bool bf16, bool dkv_f32, const CUtensorMap& tmQ, const CUtensorMap& tmDO,
                       const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmDQ, const BwdArgs& a,
                       cudaStream_t stream);
// attention-sink post-passes (ffa_aux.cu); layout 0 = "sh" [s_sink, heads], 1 = "ssh" [rows, s_sink, heads]
cudaError_t launch_sink_fwd(int out_dtype, void* out, float* lse, const float* sink, int64_t rows, int heads, int dim, int s_sink,
                            int layout, int64_t o_stride_t, int64_t o_stride_h, int64_t lse_stride_t, int64_t lse_stride_h,
                            cudaStream_t stream);
cudaError_t launch_sink_bwd(float* dsink, const float* sink, const float* lse, const float* delta, int rows, int heads, int tq_ld,
                            int s_sink, int layout, int64_t lse_stride_t, int64_t lse_stride_h, cudaStream_t stream);
}  // namespace ffa
