// ffa_aux.cu — small HBM-bound helpers of the CP runtime, exported through the C ABI:
//  * b200_lse_merge: in-place LSE merge of two partial attention results
//    (reference: correct_attn_out_lse Triton kernel, magi_attention/functional/utils.py:364-558);
//  * b200_range_reduce_sum: row-span scatter-add used by group_reduce's receive side
//    (reference: range_reduce Triton kernel, magi_attention/common/range_op/_range_reduce.py:360).
//  * attention-sink post-passes of b200_ffa_fwd / b200_ffa_bwd (launch_sink_fwd / launch_sink_bwd):
//    reference: sink columns of the fwd postprocess kernel (flash_fwd_postprocess_kernel.h:187-275),
//    calc_lse_sink / sink_bwd (magi_attention/functional/utils.py:117-206,234-280).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/b200_ffa.h"
#include "ffa_kernels.h"

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
    l = mx + log1pf(expf(mn - mx));
    w1 = l1 == -INFINITY ? 0.f : expf(l1 - l);
    w2 = l2 == -INFINITY ? 0.f : expf(l2 - l);
  }
  TO* o = out + rh * dim;
  const T2* o2 = out2 + rh * dim;
  for (int i = lane; i < dim; i += 32) o[i] = from_f<TO>(w1 * to_f(o[i]) + w2 * to_f(o2[i]));
  if (lane == 0) lse[rh] = l;
}

// Row-span accumulate.  Several entries may target the SAME dst span (an output split of group_reduce that
// receives partial dK/dV from several ranks); such entries must be ADJACENT in the lists: the first entry of a
// run owns the span and folds every source of the run in one pass (fp32 accumulate, no read-modify-write race).
template <typename T> struct Vec4;  // four consecutive elements moved with one load / store
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<__half> { using type = uint2; };
template <> struct Vec4<__nv_bfloat16> { using type = uint2; };
template <typename T> __device__ __forceinline__ void load4(const T* p, float (&x)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&x)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}
template <> __device__ __forceinline__ void load4<__half>(const __half* p, float (&x)[4]) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  const __half2 a = *reinterpret_cast<const __half2*>(&v.x), b = *reinterpret_cast<const __half2*>(&v.y);
  x[0] = __low2float(a); x[1] = __high2float(a); x[2] = __low2float(b); x[3] = __high2float(b);
}
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&x)[4]) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  x[0] = __uint_as_float(v.x << 16); x[1] = __uint_as_float(v.x & 0xffff0000u);
  x[2] = __uint_as_float(v.y << 16); x[3] = __uint_as_float(v.y & 0xffff0000u);
}
template <typename T> __device__ __forceinline__ void store4(T* p, const float (&x)[4]);
template <> __device__ __forceinline__ void store4<float>(float* p, const float (&x)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
}
template <> __device__ __forceinline__ void store4<__half>(__half* p, const float (&x)[4]) {
  const __half2 a = __floats2half2_rn(x[0], x[1]), b = __floats2half2_rn(x[2], x[3]);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&x)[4]) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(x[0], x[1]), b = __floats2bfloat162_rn(x[2], x[3]);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

// kVec: four elements per thread with 8 / 16 byte accesses (row_elems % 4 == 0 and 16-byte aligned bases, checked by the
// launcher); otherwise the scalar form.  HBM-bound: one read of every source span, one read-modify-write of the dst span.
template <typename TD, typename TS, bool kVec>
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
  if constexpr (kVec) {
    for (int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < len; i += int64_t(gridDim.x) * blockDim.x * 4) {
      float acc[4], x[4];
      load4<TD>(dp + i, acc);
      for (int k = 0; k < run; ++k) {
        load4<TS>(src + int64_t(sr[r + k].x) * row_elems + i, x);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] += x[e];
      }
      store4<TD>(dp + i, acc);
    }
  } else {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += int64_t(gridDim.x) * blockDim.x) {
      float acc = to_f(dp[i]);
      for (int k = 0; k < run; ++k) acc += to_f(src[int64_t(sr[r + k].x) * row_elems + i]);
      dp[i] = from_f<TD>(acc);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Attention sink.  The sink is `s_sink` extra logit columns per (row, head) whose value vectors are zero:
//   lse' = logaddexp(lse, logsumexp_j sink_j),  out' = out * exp(lse - lse').
// One warp per (row, head) pair; lanes < s_sink hold the sink logits, every lane rescales dim/32 output
// elements.  layout 0 ("sh"): sink [s_sink, heads]; 1 ("ssh"): sink [rows, s_sink, heads].  HBM-bound.
// ---------------------------------------------------------------------------------------------------------
template <typename TO>
__global__ void sink_fwd_kernel(TO* __restrict__ out, float* __restrict__ lse, const float* __restrict__ sink, int64_t rows,
                                int heads, int dim, int s_sink, int layout, int64_t o_stride_t, int64_t o_stride_h,
                                int64_t lse_stride_t, int64_t lse_stride_h) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t rh = warp; rh < rows * heads; rh += n_warps) {
    const int64_t row = rh / heads;
    const int h = int(rh - row * heads);
    float sv = -INFINITY;
    if (lane < s_sink) sv = sink[layout == 0 ? int64_t(lane) * heads + h : (row * s_sink + lane) * heads + h];
    float mx = sv;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = (lane < s_sink && mx != -INFINITY) ? expf(sv - mx) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    const float lse_sink = mx == -INFINITY ? -INFINITY : mx + logf(se);
    float* lp = lse + row * lse_stride_t + h * lse_stride_h;
    const float l = *lp;
    const float hi = fmaxf(l, lse_sink), lo = fminf(l, lse_sink);
    const float ln = hi == -INFINITY ? -INFINITY : hi + log1pf(expf(lo - hi));
    const float w = (l == -INFINITY) ? 0.f : expf(l - ln);
    TO* op = out + row * o_stride_t + h * o_stride_h;
    for (int i = lane; i < dim; i += 32) op[i] = from_f<TO>(to_f<TO>(op[i]) * w);
    if (lane == 0) *lp = ln;
  }
}

// dsink = -sum_rows exp(sink - lse) * delta  ("sh": reduced over the rows, fp32 atomicAdd into dsink [s_sink, heads];
// "ssh": written per row into dsink [rows, s_sink, heads]).  delta [heads, tq_ld] is the backward preprocess
// output (rowsum(O o dO)); lse is the global lse INCLUDING the sink.  grid (row blocks, heads).
constexpr int kMaxSink = 8;
__global__ void sink_bwd_kernel(float* __restrict__ dsink, const float* __restrict__ sink, const float* __restrict__ lse,
                                const float* __restrict__ delta, int rows, int heads, int tq_ld, int s_sink, int layout,
                                int64_t lse_stride_t, int64_t lse_stride_h) {
  const int h = blockIdx.y;
  float acc[kMaxSink];
  float sv[kMaxSink];
#pragma unroll
  for (int j = 0; j < kMaxSink; ++j) {
    acc[j] = 0.f;
    sv[j] = (layout == 0 && j < s_sink) ? sink[int64_t(j) * heads + h] : -INFINITY;
  }
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < rows; row += gridDim.x * blockDim.x) {
    const float l = lse[row * lse_stride_t + h * lse_stride_h];
    const float nd = -delta[int64_t(h) * tq_ld + row];
#pragma unroll
    for (int j = 0; j < kMaxSink; ++j) {
      if (j < s_sink) {
        const float s = layout == 0 ? sv[j] : sink[(int64_t(row) * s_sink + j) * heads + h];
        const float pj = (l == -INFINITY || s == -INFINITY) ? 0.f : expf(s - l);
        if (layout == 0) acc[j] = fmaf(pj, nd, acc[j]);
        else dsink[(int64_t(row) * s_sink + j) * heads + h] += pj * nd;
      }
    }
  }
  if (layout != 0) return;
  __shared__ float red[kMaxSink][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < kMaxSink; ++j) {
    float x = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[j][wid] = x;
  }
  __syncthreads();
  if (threadIdx.x < s_sink) {
    float x = 0.f;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) x += red[threadIdx.x][w];
    atomicAdd(&dsink[int64_t(threadIdx.x) * heads + h], x);
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

namespace ffa {

cudaError_t launch_sink_fwd(int out_dtype, void* out, float* lse, const float* sink, int64_t rows, int heads, int dim, int s_sink,
                            int layout, int64_t o_stride_t, int64_t o_stride_h, int64_t lse_stride_t, int64_t lse_stride_h,
                            cudaStream_t stream) {
  if (rows * heads == 0 || s_sink == 0) return cudaSuccess;
  const int threads = 256;
  const int64_t want = (rows * heads * 32 + threads - 1) / threads;
  const unsigned blocks = unsigned(want < 148 * 32 ? want : 148 * 32);  // grid-stride beyond 32 CTAs per SM
#define FFA_SINK_FWD(T_)                                                                                                  \
  sink_fwd_kernel<T_><<<blocks, threads, 0, stream>>>(static_cast<T_*>(out), lse, sink, rows, heads, dim, s_sink, layout, \
                                                      o_stride_t, o_stride_h, lse_stride_t, lse_stride_h)
  if (out_dtype == B200_F32) FFA_SINK_FWD(float);
  else if (out_dtype == B200_BF16) FFA_SINK_FWD(__nv_bfloat16);
  else FFA_SINK_FWD(__half);
#undef FFA_SINK_FWD
  return cudaGetLastError();
}

cudaError_t launch_sink_bwd(float* dsink, const float* sink, const float* lse, const float* delta, int rows, int heads, int tq_ld,
                            int s_sink, int layout, int64_t lse_stride_t, int64_t lse_stride_h, cudaStream_t stream) {
  if (rows == 0 || heads == 0 || s_sink == 0) return cudaSuccess;
  const int threads = 256;
  int bx = (rows + threads * 8 - 1) / (threads * 8);  // ~8 rows per thread before the block reduction
  if (bx < 1) bx = 1;
  dim3 grid(bx, heads);
  sink_bwd_kernel<<<grid, threads, 0, stream>>>(dsink, sink, lse, delta, rows, heads, tq_ld, s_sink, layout, lse_stride_t, lse_stride_h);
  return cudaGetLastError();
}

}  // namespace ffa

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

int b200_sink_merge(void* out, float* lse, const float* sink, int64_t rows, int32_t heads, int32_t dim, int32_t seqlen_sink,
                    int32_t sink_layout, int32_t out_dtype, int64_t o_stride_t, int64_t o_stride_h, int64_t lse_stride_t,
                    int64_t lse_stride_h, b200_stream_t stream_) {
  if (rows == 0 || heads == 0 || seqlen_sink == 0) return 0;
  if (!out || !lse || !sink || rows < 0 || heads < 0 || dim <= 0) return B200_ERR_INVALID;
  if (seqlen_sink < 0 || seqlen_sink > 8 || (sink_layout != 0 && sink_layout != 1)) return B200_ERR_INVALID;
  if (out_dtype != B200_F16 && out_dtype != B200_BF16 && out_dtype != B200_F32) return B200_ERR_INVALID;
  cudaError_t e = ffa::launch_sink_fwd(out_dtype, out, lse, sink, rows, heads, dim, seqlen_sink, sink_layout, o_stride_t, o_stride_h,
                                       lse_stride_t, lse_stride_h, static_cast<cudaStream_t>(stream_));
  return e == cudaSuccess ? int(B200_OK) : int(B200_ERR_CUDA);
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
    const bool vec = (row_elems % 4 == 0) && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) % 16 == 0);
    if (vec)
      range_reduce_kernel<TD, TS, true><<<grid, 256, 0, stream>>>(static_cast<TD*>(dst), static_cast<const TS*>(src),
                                                                 reinterpret_cast<const int2*>(dst_ranges),
                                                                 reinterpret_cast<const int2*>(src_ranges), num_ranges, row_elems);
    else
      range_reduce_kernel<TD, TS, false><<<grid, 256, 0, stream>>>(static_cast<TD*>(dst), static_cast<const TS*>(src),
                                                                  reinterpret_cast<const int2*>(dst_ranges),
                                                                  reinterpret_cast<const int2*>(src_ranges), num_ranges, row_elems);
    return cudaGetLastError() == cudaSuccess ? int(B200_OK) : int(B200_ERR_CUDA);
  });
}

}  // extern "C"
