// ffa_aux.cu — small HBM-bound helpers of the CP runtime, exported through the C ABI:
//  * b200_lse_merge: in-place LSE merge of two partial attention results
//    (reference: correct_attn_out_lse Triton kernel, magi_attention/functional/utils.py:364-558);
//  * b200_range_reduce_sum: row-span scatter-add used by group_reduce's receive side
//    (reference: range_reduce Triton kernel, magi_attention/common/range_op/_range_reduce.py:360).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/b200_ffa.h"

namespace {

template <typename T> __device__ __forceinline__ float to_f(T x);
template <> __device__ __forceinline__ float to_f<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f<__half>(__half x) { return __half2float(x); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __half from_f<__half>(float x) { return __float2half_rn(x); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// one warp per (row, head); dim elements strided over lanes
template <typename TO, typename T2>
__global__ void lse_merge_kernel(TO* __restrict__ out, float* __restrict__ lse, const T2* __restrict__ out2,
                                 const float* __restrict__ lse2, int64_t n_rh, int dim) {
  int64_t rh = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (rh >= n_rh) return;
  float l1 = lse[rh], l2 = lse2[rh];
  float mx = fmaxf(l1, l2), mn = fminf(l1, l2);
  float w1, w2, l;
  if (mx == -INFINITY) {
    w1 = 0.f; w2 = 0.f; l = -INFINITY;
  } else {
    l = mx + log1pf(__expf(mn - mx));
    w1 = l1 == -INFINITY ? 0.f : __expf(l1 - l);
    w2 = l2 == -INFINITY ? 0.f : __expf(l2 - l);
  }
  TO* o = out + rh * dim;
  const T2* o2 = out2 + rh * dim;
  for (int i = lane; i < dim; i += 32) o[i] = from_f<TO>(w1 * to_f(o[i]) + w2 * to_f(o2[i]));
  if (lane == 0) lse[rh] = l;
}

// Row-span accumulate.  Several entries may target the SAME dst span (an output split of group_reduce that
// receives partial dK/dV from several ranks); such entries must be ADJACENT in the lists: the first entry of a
// run owns the span and folds every source of the run in one pass (fp32 accumulate, no read-modify-write race).
template <typename TD, typename TS>
__global__ void range_reduce_kernel(TD* __restrict__ dst, const TS* __restrict__ src, const int2* __restrict__ dr,
                                    const int2* __restrict__ sr, int n, int64_t row_elems) {
  const int r = blockIdx.y;
  if (r >= n) return;
  const int2 d = dr[r];
  if (r > 0) {
    const int2 p = dr[r - 1];
    if (p.x == d.x && p.y == d.y) return;  // not the run leader
  }
  int run = 1;
  while (r + run < n) {
    const int2 nx = dr[r + run];
    if (nx.x != d.x || nx.y != d.y) break;
    ++run;
  }
  const int64_t len = int64_t(d.y - d.x) * row_elems;
  TD* dp = dst + int64_t(d.x) * row_elems;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += int64_t(gridDim.x) * blockDim.x) {
    float acc = to_f(dp[i]);
    for (int k = 0; k < run; ++k) acc += to_f(src[int64_t(sr[r + k].x) * row_elems + i]);
    dp[i] = from_f<TD>(acc);
  }
}

template <typename F>
int dispatch2(int a, int b, F&& f) {
  auto inner = [&](auto ta) {
    switch (b) {
      case B200_F16: return f(ta, __half{});
      case B200_BF16: return f(ta, __nv_bfloat16{});
      case B200_F32: return f(ta, float{});
    }
    return int(B200_ERR_INVALID);
  };
  switch (a) {
    case B200_F16: return inner(__half{});
    case B200_BF16: return inner(__nv_bfloat16{});
    case B200_F32: return inner(float{});
  }
  return int(B200_ERR_INVALID);
}

}  // namespace

extern "C" {

int b200_lse_merge(void* out, float* lse, const void* out2, const float* lse2, int64_t rows, int32_t heads, int32_t dim,
                   int32_t out_dtype, int32_t out2_dtype, b200_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int64_t n_rh = rows * heads;
  if (n_rh == 0) return 0;
  if (!out || !lse || !out2 || !lse2) return B200_ERR_INVALID;
  return dispatch2(out_dtype, out2_dtype, [&](auto ta, auto tb) {
    using TO = decltype(ta);
    using T2 = decltype(tb);
    int threads = 256;
    int64_t blocks = (n_rh * 32 + threads - 1) / threads;
    lse_merge_kernel<TO, T2><<<unsigned(blocks), threads, 0, stream>>>(static_cast<TO*>(out), lse, static_cast<const T2*>(out2),
                                                                      lse2, n_rh, dim);
    return cudaGetLastError() == cudaSuccess ? int(B200_OK) : int(B200_ERR_CUDA);
  });
}

int b200_range_reduce_sum(void* dst, int32_t dst_dtype, const void* src, int32_t src_dtype, const int32_t* dst_ranges,
                          const int32_t* src_ranges, int32_t num_ranges, int64_t row_elems, b200_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (num_ranges == 0) return 0;
  if (!dst || !src || !dst_ranges || !src_ranges) return B200_ERR_INVALID;
  return dispatch2(dst_dtype, src_dtype, [&](auto ta, auto tb) {
    using TD = decltype(ta);
    using TS = decltype(tb);
    dim3 grid(64, num_ranges);
    range_reduce_kernel<TD, TS><<<grid, 256, 0, stream>>>(static_cast<TD*>(dst), static_cast<const TS*>(src),
                                                         reinterpret_cast<const int2*>(dst_ranges),
                                                         reinterpret_cast<const int2*>(src_ranges), num_ranges, row_elems);
    return cudaGetLastError() == cudaSuccess ? int(B200_OK) : int(B200_ERR_CUDA);
  });
}

}  // extern "C"
