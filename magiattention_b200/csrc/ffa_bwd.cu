// ffa_bwd.cu — flex-flash-attention BACKWARD for sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces the reference's Hopper-only FFA backward (csrc/flexible_flash_attention/
// flash_bwd_preprocess_kernel.h, flash_bwd_kernel_sm90.h, mainloop_bwd_sm90_tma_gmma_ws.hpp,
// epilogue_bwd.hpp, flash_bwd_postprocess_kernel.h) and the FA4 CuTe-DSL backward it uses on Blackwell
// (flash_attn/cute/flash_bwd_sm100.py).  Math as functional/sdpa.py:270-327 with the GLOBAL lse / out:
//     P = exp2(S*scale*log2e - LSE*log2e)   dV += P^T dO    dP = dO V^T    Delta = rowsum(O o dO)
//     dS = P o (dP - Delta)                 dK += scale * dS^T Q           dQ += scale * dS K
//
//  * preprocess kernel: Delta and LSE*log2e into padded [heads, tq_ld] fp32 rows (dead rows: +inf so P = 0).
//  * main kernel: one CTA owns a 128-row kv tile of one kv head, keeps dK and dV in TMEM for the whole
//    loop over (q heads of the GQA group) x (slices touching the tile, kv-tile -> slice CSR) x (128-row q
//    blocks): no atomics and no zero-fill pass for dK/dV.  Everything is computed TRANSPOSED (kv on the
//    TMEM lanes) so that P^T and dS^T are directly the TMEM A operands of the dV and dK MMAs:
//        S^T = K Q^T (SS)   dP^T = V dO^T (SS)   dV += P^T dO (TS)   dK += dS^T Q (TS)   dQ = dS K (SS)
//    dS^T also goes to shared memory (it is the MN-major A operand of the dQ MMA).  dQ tiles are drained
//    from TMEM by a dedicated warpgroup and reduced into the fp32 global dq with red.global.add.v4.f32.
//  * TMEM (512 cols): S^T/P^T [0,128)  dP^T/dS^T/dQ [128,256)  dV [256,384)  dK [384,512).
//  * warps: 0-7 compute (two warpgroups, each owns 64 q columns of every kv row), 8-11 dQ drain,
//    12 MMA issuer, 13 TMA producer.  K,V resident; Q (+LSE2, Delta) and dO double-buffered.
#include <cuda.h>
#include <cmath>
#include <type_traits>
#include "../../include/b200_ffa.h"
#include "ffa_args.h"
#include "ffa_common.cuh"
#include "ffa_kernels.h"
#include "sm100.cuh"

// Timing-only ablation switches for kernel analysis (results become WRONG; never set in the product build):
//   1 drain skips staging + TMA reduce   2 drain skips its TMEM loads too   4 no dQ MMA   8 exp2 -> fma
//   16 pass 2 skips the dS^T st.shared   32 lse / delta read once instead of per element group
//   64 drain stages into smem but issues no TMA reduce   128 drain issues the TMA reduce without writing the staging
#ifndef FFA_BWD_ABL
#define FFA_BWD_ABL 0
#endif
// Register budget per role (threads x regs must fit 64K: 256 x COMPUTE + 128 x DRAIN + 128 x OTHER).  The dQ drain warps hold
// a whole 128 x 128 fp32 tile in registers so that the shared TMEM region is handed back to the MMA warp right after the
// loads instead of after the first two staged chunks (the measured hold time of the region fell from ~990 to ~250 clocks).
#ifndef FFA_BWD_REGS_COMPUTE
#define FFA_BWD_REGS_COMPUTE 136
#endif
#ifndef FFA_BWD_REGS_DRAIN
#define FFA_BWD_REGS_DRAIN 168
#endif
#ifndef FFA_BWD_REGS_OTHER
#define FFA_BWD_REGS_OTHER 72
#endif
// dK operand source: 0 (default) = dS^T from TMEM (TS MMA), dK before dQ;  1 = dS^T read from its shared-memory copy (SS MMA)
// with dQ issued BEFORE dK so that its drain starts a whole MMA earlier.  Measured equal-to-slightly-slower (round 2, causal
// 32K x 8 heads: 4.42 ms vs 4.38 ms): once the TMEM region is released before staging, the loop is bound by the dQ
// reduce path itself (st.shared staging + TMA reduce-add ~ 770 clocks per 16 KB chunk), not by the MMA order.
#ifndef FFA_BWD_DK_SS
#define FFA_BWD_DK_SS 0
#endif
static_assert(256 * FFA_BWD_REGS_COMPUTE + 128 * FFA_BWD_REGS_DRAIN + 128 * FFA_BWD_REGS_OTHER <= 65536, "register budget");
// (measured and rejected, round 2: draining dQ straight from registers with sector-coalesced red.global.add.v2.f32 —
//  tcgen05.ld.16x256b fragments, four lanes per 32-byte sector, no shared-memory staging — is 29 % SLOWER than the
//  smem-staged TMA reduce-add: SM-issued reductions are far more expensive at L2 than UTMAREDG; tools/probe/probe_ld.cu
//  keeps the fragment-layout probe.)

// Developer timeline (-DFFA_TRACE): CTA 0 stamps %clock64 at the hand-off points of iterations [32, 96) — see ffa_fwd.cu.
#ifdef FFA_TRACE
__device__ unsigned long long g_trace_bwd[4][64][8];
#define FFA_TRB(role, it, ev)                                                                                        \
  do {                                                                                                                \
    if (blockIdx.x == 0 && (it) >= 32u && (it) < 96u && (threadIdx.x & 31) == 0) g_trace_bwd[role][(it) - 32u][ev] = clock64(); \
  } while (0)
extern "C" __attribute__((visibility("default"))) int b200_debug_trace_bwd(unsigned long long* dst) {
  return cudaMemcpyFromSymbol(dst, g_trace_bwd, sizeof(g_trace_bwd)) == cudaSuccess ? 0 : -1;
}
#else
#define FFA_TRB(role, it, ev) do {} while (0)
#endif

namespace ffa {
using namespace sm100;

namespace bwd {
constexpr int kThreads = 512;
constexpr uint32_t kBoxBytes = 128 * 64 * 2;  // one TMA box: 128 rows x 64 columns of 16-bit elements
// K, V resident; dS^T staging; Q double-buffered; dO single-buffered (its next load is issued as soon as dV
// retires, ~1.5k clocks before dP needs it); two 16 KB staging buffers for the TMA reduce-add of dQ.
constexpr uint32_t kDSBytes = 2 * kBoxBytes;      // dS^T staging: 128 kv rows x 128 q columns
constexpr uint32_t kDQChunkBytes = 128 * 32 * 4;  // one dQ staging chunk: 128 q rows x 32 fp32 columns
template <int kD> struct Cfg {
  static_assert(kD == 64 || kD == 128, "head_dim must be 64 or 128");
  static constexpr uint32_t kBoxes = kD / 64;
  static constexpr uint32_t kTileBytes = kBoxes * kBoxBytes;  // K, V, Q, dO tiles: 128 rows x kD
  static constexpr uint32_t kSmemK = 0, kSmemV = kTileBytes, kSmemDS = 2 * kTileBytes, kSmemQ = kSmemDS + kDSBytes,
                            kSmemDO = kSmemQ + 2 * kTileBytes, kSmemDQ = kSmemDO + kTileBytes,
                            kSmemLse = kSmemDQ + 2 * kDQChunkBytes, kSmemDelta = kSmemLse + 1024,
                            kSmemBar = kSmemDelta + 1024, kSmemBytes = kSmemBar + 256;
};
constexpr int kBarKV = 0, kBarQFull = 1, kBarQEmpty = 3, kBarDOFull = 5, kBarDOEmpty = 6, kBarDSFree = 7, kBarS = 9, kBarDP = 10,
              kBarP = 11, kBarDS = 12, kBarDQ = 13, kBarDrained = 14, kBarDKV = 15, kNumBars = 16;
constexpr uint32_t kTmemS = 0, kTmemDP = 128, kTmemDV = 256, kTmemDK = 384;
}  // namespace bwd

// -------------------------------------------------------------------------------------------------
// preprocess: delta[h][r] = sum_d O[r,h,d]*dO[r,h,d];  lse2[h][r] = lse[r,h]*log2(e)  (+inf when dead/pad)
// -------------------------------------------------------------------------------------------------
template <int kD, bool kBF16>
__global__ void ffa_bwd_preprocess_kernel(const void* __restrict__ out, const void* __restrict__ dout,
                                          const float* __restrict__ lse, float* __restrict__ delta,
                                          float* __restrict__ lse2, int total_q, int num_heads, int tq_ld,
                                          int64_t o_stride_t, int64_t o_stride_h, int64_t do_stride_t,
                                          int64_t do_stride_h, int64_t lse_stride_t, int64_t lse_stride_h) {
  const int warps_per_block = blockDim.x >> 5;
  const int64_t idx = int64_t(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t total = int64_t(num_heads) * tq_ld;
  if (idx >= total) return;
  const int h = int(idx / tq_ld), r = int(idx % tq_ld);
  if (r >= total_q) {
    if (lane == 0) {
      delta[idx] = 0.f;
      lse2[idx] = INFINITY;
    }
    return;
  }
  using T = typename std::conditional<kBF16, __nv_bfloat16, __half>::type;
  const T* o = reinterpret_cast<const T*>(out) + int64_t(r) * o_stride_t + int64_t(h) * o_stride_h;
  const T* d = reinterpret_cast<const T*>(dout) + int64_t(r) * do_stride_t + int64_t(h) * do_stride_h;
  constexpr int kWords = kD / 64;  // 32-bit words (element pairs) per lane
  uint32_t ow[kWords], dw[kWords];
  if constexpr (kD == 128) {
    const uint2 ov = *reinterpret_cast<const uint2*>(o + lane * 4), dv = *reinterpret_cast<const uint2*>(d + lane * 4);
    ow[0] = ov.x; ow[1] = ov.y; dw[0] = dv.x; dw[1] = dv.y;
  } else {
    ow[0] = *reinterpret_cast<const uint32_t*>(o + lane * 2);
    dw[0] = *reinterpret_cast<const uint32_t*>(d + lane * 2);
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < kWords; ++i) {
    float o0, o1, d0, d1;
    if constexpr (kBF16) {
      o0 = __uint_as_float(ow[i] << 16); o1 = __uint_as_float(ow[i] & 0xffff0000u);
      d0 = __uint_as_float(dw[i] << 16); d1 = __uint_as_float(dw[i] & 0xffff0000u);
    } else {
      __half2 oh = *reinterpret_cast<const __half2*>(&ow[i]), dh = *reinterpret_cast<const __half2*>(&dw[i]);
      o0 = __low2float(oh); o1 = __high2float(oh); d0 = __low2float(dh); d1 = __high2float(dh);
    }
    acc = fmaf(o0, d0, acc);
    acc = fmaf(o1, d1, acc);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) {
    const float l = lse[int64_t(r) * lse_stride_t + int64_t(h) * lse_stride_h];
    delta[idx] = acc;
    lse2[idx] = (l == -INFINITY) ? INFINITY : l * 1.4426950408889634f;
  }
}

// -------------------------------------------------------------------------------------------------
// main kernel
// -------------------------------------------------------------------------------------------------
// Per-slice geometry for one backward CTA tile (kv rows [C0, C0+128)).
struct BwdGeom {
  int c_lo, c_hi, r_min, r_max, r_start, nblk, full_lo, full_hi;
  bool cols_full;
  __device__ __forceinline__ bool init(const Slice& s, int C0) {
    c_lo = max(s.k0, C0);
    c_hi = min(s.k1, C0 + kBwdTileK);
    if (!row_hull(s, c_lo, c_hi, r_min, r_max)) return false;
    r_start = r_min & ~3;  // 16-byte aligned LSE/Delta bulk copies
    nblk = (r_max - r_start + kBlockM - 1) / kBlockM;
    cols_full = (c_lo == C0) && (c_hi == C0 + kBwdTileK);
    full_lo = s.rlo(C0 + kBwdTileK - 1);  // largest first-visible row over the tile's columns
    full_hi = s.rhi(C0);                  // smallest end row
    return true;
  }
  __device__ __forceinline__ bool full(int rs) const { return cols_full && rs >= full_lo && rs + kBlockM <= full_hi; }
};

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}

template <int kD, bool kBF16, bool kDkvF32>
__global__ void __launch_bounds__(bwd::kThreads, 1)
ffa_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
               const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
               const __grid_constant__ CUtensorMap tmDQ, const BwdArgs args) {
  using namespace bwd;
  using C = Cfg<kD>;
  constexpr uint32_t kTileBytes = C::kTileBytes, kBoxes = C::kBoxes;
  constexpr uint32_t kSmemK = C::kSmemK, kSmemV = C::kSmemV, kSmemDS = C::kSmemDS, kSmemQ = C::kSmemQ, kSmemDO = C::kSmemDO,
                     kSmemDQ = C::kSmemDQ, kSmemLse = C::kSmemLse, kSmemDelta = C::kSmemDelta, kSmemBar = C::kSmemBar;
  constexpr int kDh = kD / 2;  // head-dim columns each compute warpgroup writes in the dK/dV epilogue
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmemBar + kNumBars * 8);
  const uint32_t bar0 = smem_base + kSmemBar;
  auto bar = [&](int i) -> uint32_t { return bar0 + uint32_t(i) * 8u; };

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int hk = blockIdx.x / args.n_tiles;
  const int tile = blockIdx.x - hk * args.n_tiles;
  const int C0 = tile * kBwdTileK;
  const int G = args.num_heads_q / args.num_heads_kv;

  if (tid == 0) {
    if (smem_base & 1023u) {
      printf("ffa_bwd: dynamic smem not 1024-aligned\n");
      __trap();
    }
    for (int i = 0; i < kNumBars; ++i) {
      uint32_t cnt = 1;
      if (i == kBarP || i == kBarDS) cnt = 256;
      if (i == kBarDrained) cnt = 128;
      mbar_init(bar(i), cnt);
    }
    fence_barrier_init();
  }
  if (wid == 12) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  if (wid == 13 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmDO);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    prefetch_tmap(&tmDQ);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (wid < 8) {
    // ================================ compute warpgroups ================================
    setmaxnreg_inc<FFA_BWD_REGS_COMPUTE>();
    const int half = wid >> 2;                       // which 64 q columns of the block
    const int krow = (wid & 3) * 32 + lane;          // kv row inside the tile == TMEM lane
    const int c = C0 + krow;                         // global kv index
    const uint32_t lane_off = uint32_t((wid & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_off + kTmemS + half * 64;
    const uint32_t tDP = tmem + lane_off + kTmemDP + half * 64;
    const float sl2 = args.scale * 1.4426950408889634f;
    uint8_t* sds_row = smem + kSmemDS + half * kBoxBytes + krow * 128;
    const int swz = krow & 7;

    uint32_t it = 0;
    for (int g = 0; g < G; ++g) {
      SliceCursor cur;
      cur.init(args.meta, args.plan, tile);
      Slice sl;
      BwdGeom gm;
      while (cur.next(sl)) {
        if (!gm.init(sl, C0)) continue;
        const bool col_in = (c >= sl.k0) && (c < sl.k1);
        const int rlo_c = col_in ? sl.rlo(c) : 0;
        const int rhi_c = col_in ? sl.rhi(c) : 0;
        for (int b = 0; b < gm.nblk; ++b, ++it) {
          const int rs = gm.r_start + b * kBlockM;
          const uint32_t st = it & 1;
          const bool is_full = gm.full(rs);
          mbar_wait(bar(kBarQFull + st), (it >> 1) & 1);  // LSE2 / Delta of this block are in smem
          const float* s_lse = reinterpret_cast<const float*>(smem + kSmemLse + st * 512) + half * 64;
          const float* s_del = reinterpret_cast<const float*>(smem + kSmemDelta + st * 512) + half * 64;
          if (wid == 0) FFA_TRB(1, it, 0);
          mbar_wait(bar(kBarS), it & 1);
          tc_fence_after();
          if (wid == 0) FFA_TRB(1, it, 1);
          float p[64];
          // ---- pass 1: P^T = exp2(S^T * sl2 - lse2[q]) -> TMEM (A operand of the dV MMA)
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t sr[32];
            tmem_ld32(tS + ch * 32, sr);
            tc_wait_ld();
            uint32_t bits = 0xffffffffu;
            if (!is_full) {
              const int q_base = rs + half * 64 + ch * 32;
              const int lo_c = min(max(rlo_c - q_base, 0), 32), hi_c = min(max(rhi_c - q_base, 0), 32);
              const uint32_t m_hi = hi_c >= 32 ? 0xffffffffu : ((1u << hi_c) - 1u);
              const uint32_t m_lo = lo_c >= 32 ? 0xffffffffu : ((1u << lo_c) - 1u);
              bits = m_hi & ~m_lo;
            }
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 l4 = *reinterpret_cast<const float4*>(s_lse + ((FFA_BWD_ABL & 32) ? 0 : ch * 32 + j * 4));
              const float lv[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = j * 4 + e;
                float x = (FFA_BWD_ABL & 8) ? fmaf(__uint_as_float(sr[i]), sl2, -lv[e]) : ex2(fmaf(__uint_as_float(sr[i]), sl2, -lv[e]));
                x = ((bits >> i) & 1u) ? x : 0.f;
                p[ch * 32 + i] = x;
              }
              pk[j * 2] = pack2<kBF16>(p[ch * 32 + j * 4], p[ch * 32 + j * 4 + 1]);
              pk[j * 2 + 1] = pack2<kBF16>(p[ch * 32 + j * 4 + 2], p[ch * 32 + j * 4 + 3]);
            }
            tmem_st16(tS + ch * 16, pk);
          }
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(bar(kBarP));
          if (wid == 0) FFA_TRB(1, it, 2);
          // ---- pass 2: dS^T = P^T o (dP^T - delta[q]) -> TMEM (A of dK MMA) and smem (A of dQ MMA)
          mbar_wait(bar(kBarDP), it & 1);
          tc_fence_after();
          if (wid == 0) FFA_TRB(1, it, 3);
          // the MMAs that read the previous dS^T copy in shared memory have retired
          if (it > 0) mbar_wait(bar(FFA_BWD_DK_SS ? kBarDSFree : kBarDQ), (it - 1) & 1);
          if (wid == 0) FFA_TRB(1, it, 4);
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t dr[32];
            tmem_ld32(tDP + ch * 32, dr);
            tc_wait_ld();
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 d4 = *reinterpret_cast<const float4*>(s_del + ((FFA_BWD_ABL & 32) ? 0 : ch * 32 + j * 4));
              const float dl[4] = {d4.x, d4.y, d4.z, d4.w};
              float ds[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) ds[e] = p[ch * 32 + j * 4 + e] * (__uint_as_float(dr[j * 4 + e]) - dl[e]);
              pk[j * 2] = pack2<kBF16>(ds[0], ds[1]);
              pk[j * 2 + 1] = pack2<kBF16>(ds[2], ds[3]);
            }
            if (!FFA_BWD_DK_SS) tmem_st16(tDP + ch * 16, pk);
#pragma unroll
            for (int j = 0; j < ((FFA_BWD_ABL & 16) ? 0 : 4); ++j) {
              const int chunk = ch * 4 + j;  // 16-byte chunk (8 q values) inside this thread's 128-byte row
              *reinterpret_cast<uint4*>(sds_row + ((chunk ^ swz) << 4)) =
                  make_uint4(pk[j * 4], pk[j * 4 + 1], pk[j * 4 + 2], pk[j * 4 + 3]);
            }
          }
          if (!FFA_BWD_DK_SS) tc_wait_st();
          tc_fence_before();
          fence_proxy_async_smem();
          mbar_arrive(bar(kBarDS));
          if (wid == 0) FFA_TRB(1, it, 5);
        }
      }
    }
    // ------------------------------- dK / dV epilogue -------------------------------
    if (it > 0) {
      mbar_wait(bar(kBarDKV), 0);
      tc_fence_after();
    }
    const bool row_ok = c < args.total_k;
    using DT = typename std::conditional<kDkvF32, float, typename std::conditional<kBF16, __nv_bfloat16, __half>::type>::type;
#pragma unroll
    for (int which = 0; which < 2; ++which) {  // 0: dV, 1: dK
      DT* base = reinterpret_cast<DT*>(which == 0 ? args.dv : args.dk);
      const int64_t st_t = which == 0 ? args.dv_stride_t : args.dk_stride_t;
      const int64_t st_h = which == 0 ? args.dv_stride_h : args.dk_stride_h;
      DT* rowp = base + int64_t(c) * st_t + int64_t(hk) * st_h + half * kDh;
      const float mul = which == 0 ? 1.f : args.scale;
      const uint32_t tsrc = tmem + lane_off + (which == 0 ? kTmemDV : kTmemDK) + half * kDh;
#pragma unroll
      for (int ch = 0; ch < kDh / 32; ++ch) {
        float x[32];
        if (it > 0) {
          uint32_t r[32];
          tmem_ld32(tsrc + ch * 32, r);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(r[i]) * mul;
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = 0.f;
        }
        if (!row_ok) continue;
        if constexpr (kDkvF32) {
          float4* dst = reinterpret_cast<float4*>(rowp + ch * 32);
          if (args.accumulate_dkv) {
            if (it > 0) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float4 y = dst[i];
                dst[i] = make_float4(y.x + x[4 * i], y.y + x[4 * i + 1], y.z + x[4 * i + 2], y.w + x[4 * i + 3]);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
          }
        } else {
          uint4* dst = reinterpret_cast<uint4*>(rowp + ch * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[i] = make_uint4(pack2<kBF16>(x[8 * i], x[8 * i + 1]), pack2<kBF16>(x[8 * i + 2], x[8 * i + 3]),
                                pack2<kBF16>(x[8 * i + 4], x[8 * i + 5]), pack2<kBF16>(x[8 * i + 6], x[8 * i + 7]));
        }
      }
    }
  } else if (wid < 12) {
    // ================================ dQ drain warpgroup ================================
    // dQ tile (128 q rows x 128 d, fp32) : TMEM -> registers (x scale) -> 128B-swizzled smem chunk of 32 columns
    // -> cp.reduce.async.bulk.tensor (TMA reduce-add into the fp32 global dq).  The TMA engine does the L2
    // reductions in full lines; the SM only pays 8 st.shared.v4 per thread per chunk.
    setmaxnreg_inc<FFA_BWD_REGS_DRAIN>();
    const uint32_t lane_off = uint32_t((wid & 3) * 32) << 16;
    const uint32_t tDQ = tmem + lane_off + kTmemDP;
    const int qrow_in_blk = (wid & 3) * 32 + lane;
    const int dtid = tid - 256;  // 0..127 inside the drain warpgroup
    uint8_t* stage_row = smem + kSmemDQ + qrow_in_blk * 128;
    const int swz = qrow_in_blk & 7;
    uint32_t it = 0, chunk_ctr = 0;
    for (int g = 0; g < G; ++g) {
      const int hq = hk * G + g;
      SliceCursor cur;
      cur.init(args.meta, args.plan, tile);
      Slice sl;
      BwdGeom gm;
      while (cur.next(sl)) {
        if (!gm.init(sl, C0)) continue;
        for (int b = 0; b < gm.nblk; ++b, ++it) {
          const int rs = gm.r_start + b * kBlockM;
          if (wid == 8) FFA_TRB(2, it, 0);
          mbar_wait(bar(kBarDQ), it & 1);
          tc_fence_after();
          if (wid == 8) FFA_TRB(2, it, 1);
          // the whole tile goes to registers first, so the TMEM region is released (kBarDrained) before any staging
          uint32_t r[kD / 32][32];
          if (!(FFA_BWD_ABL & 2)) {
#pragma unroll
            for (int c = 0; c < kD / 32; ++c) tmem_ld32(tDQ + c * 32, r[c]);
            tc_wait_ld();
          } else {
#pragma unroll
            for (int c = 0; c < kD / 32; ++c)
#pragma unroll
              for (int i = 0; i < 32; ++i) r[c][i] = 0;
          }
          tc_fence_before();
          mbar_arrive(bar(kBarDrained));  // TMEM dQ region may be overwritten by the next dP
          if (wid == 8) FFA_TRB(2, it, 2);
#pragma unroll
          for (int c = 0; c < ((FFA_BWD_ABL & 1) ? 0 : kD / 32); ++c, ++chunk_ctr) {
            const uint32_t buf = chunk_ctr & 1;
            if (dtid == 0) tma_store_wait_read<1>();  // the reduce that last read this buffer has consumed it
            named_bar_sync(1, 128);
            uint8_t* dst = stage_row + buf * kDQChunkBytes;
#pragma unroll
            for (int j = 0; j < ((FFA_BWD_ABL & 128) ? 0 : 8); ++j) {
              *reinterpret_cast<float4*>(dst + ((j ^ swz) << 4)) = make_float4(
                  __uint_as_float(r[c][4 * j]) * args.scale, __uint_as_float(r[c][4 * j + 1]) * args.scale,
                  __uint_as_float(r[c][4 * j + 2]) * args.scale, __uint_as_float(r[c][4 * j + 3]) * args.scale);
            }
            fence_proxy_async_smem();
            named_bar_sync(1, 128);
            if (dtid == 0 && !(FFA_BWD_ABL & 64)) {
              tma_reduce_add_3d(&tmDQ, smem_base + kSmemDQ + buf * kDQChunkBytes, c * 32, hq, rs);
              tma_store_commit();
            }
          }
          if (wid == 8) FFA_TRB(2, it, 3);
        }
      }
    }
    if (dtid == 0) tma_store_wait<0>();
  } else {
    setmaxnreg_dec<FFA_BWD_REGS_OTHER>();
    if (wid == 12) {
      {
        // ================================ MMA issuer ================================
        // warp-uniform loop (descriptors stay in uniform registers); one elected lane issues the async ops
        constexpr uint32_t id_kk = make_idesc(kBF16, false, false, 128, 128);  // S^T, dP^T [128 kv x 128 q], contraction over kD
        constexpr uint32_t id_tm = make_idesc(kBF16, false, true, 128, kD);    // dV, dK [128 kv x kD]: A tmem, B MN-major
        constexpr uint32_t id_mm = make_idesc(kBF16, true, true, 128, kD);     // dQ [128 q x kD]: A, B MN-major
        constexpr uint64_t mn_hi = desc_mnmajor(kBoxBytes);
        const uint32_t sK = smem_base + kSmemK, sV = smem_base + kSmemV, sDS = smem_base + kSmemDS;
        const uint32_t tS = tmem + kTmemS, tDP = tmem + kTmemDP, tDV = tmem + kTmemDV, tDK = tmem + kTmemDK;

        // each helper issues its MMA batch and then commits to up to two mbarriers (0 = none)
        auto mma_kmajor = [&](uint32_t d, uint32_t a, uint32_t b, uint32_t commit_bar) {
          if (elect_one()) {
#pragma unroll
            for (uint32_t kk = 0; kk < kD / 16; ++kk) {  // contraction over the head dim
              const uint32_t off = (kk >> 2) * kBoxBytes + (kk & 3) * 32;
              mma_ss(d, smem_desc(DESC_KMAJOR, a + off), smem_desc(DESC_KMAJOR, b + off), id_kk, kk > 0 ? 1u : 0u);
            }
            tc_commit(commit_bar);
          }
          __syncwarp();
        };
        auto mma_tmemA = [&](uint32_t d, uint32_t a_tmem, uint32_t b, bool acc, uint32_t commit_bar) {
          if (elect_one()) {
#pragma unroll
            for (uint32_t kk = 0; kk < 8; ++kk)
              mma_ts(d, a_tmem + (kk >> 2) * 64 + (kk & 3) * 8, smem_desc(mn_hi, b + kk * 2048), id_tm, (acc || kk > 0) ? 1u : 0u);
            tc_commit(commit_bar);
          }
          __syncwarp();
        };

        // total number of iterations for this CTA (all q heads walk the same slice list)
        uint32_t per_head = 0;
        {
          SliceCursor cur;
          cur.init(args.meta, args.plan, tile);
          Slice sl;
          BwdGeom gm;
          while (cur.next(sl))
            if (gm.init(sl, C0)) per_head += gm.nblk;
        }
        const uint32_t n_it = per_head * uint32_t(G);
        if (n_it > 0) {
          mbar_wait(bar(kBarKV), 0);
          // prologue: S_0, dP_0
          mbar_wait(bar(kBarQFull + 0), 0);
          tc_fence_after();
          mma_kmajor(tS, sK, smem_base + kSmemQ, bar(kBarS));
          mbar_wait(bar(kBarDOFull), 0);
          tc_fence_after();
          mma_kmajor(tDP, sV, smem_base + kSmemDO, bar(kBarDP));
          for (uint32_t it = 0; it < n_it; ++it) {
            const uint32_t st = it & 1, nst = st ^ 1;
            const uint32_t sQ = smem_base + kSmemQ + st * kTileBytes, sDO = smem_base + kSmemDO;
            const bool has_next = it + 1 < n_it;
            // dV += P^T dO
            FFA_TRB(0, it, 0);
            mbar_wait(bar(kBarP), it & 1);
            tc_fence_after();
            FFA_TRB(0, it, 1);
            mma_tmemA(tDV, tS, sDO, it > 0, bar(kBarDOEmpty));  // dO_it no longer needed once dV retires
            // S_{it+1} (S region is free after dV in tensor-pipe order)
            if (has_next) {
              mbar_wait(bar(kBarQFull + nst), ((it + 1) >> 1) & 1);
              tc_fence_after();
              mma_kmajor(tS, sK, smem_base + kSmemQ + nst * kTileBytes, bar(kBarS));
            }
            // dK += dS^T Q ; dQ = dS K
            FFA_TRB(0, it, 2);
            mbar_wait(bar(kBarDS), it & 1);
            tc_fence_after();
            FFA_TRB(0, it, 3);
            if (!FFA_BWD_DK_SS) mma_tmemA(tDK, tDP, sQ, it > 0, bar(kBarQEmpty + st));
            if (elect_one()) {
#pragma unroll
              for (uint32_t kk = 0; kk < ((FFA_BWD_ABL & 4) ? 0 : 8); ++kk)
                mma_ss(tDP, smem_desc(mn_hi, sDS + kk * 2048), smem_desc(mn_hi, sK + kk * 2048), id_mm, kk > 0 ? 1u : 0u);
              tc_commit(bar(kBarDQ));
              if (FFA_BWD_DK_SS) {
                // dK += dS^T Q with dS^T [128 kv x 128 q] read K-major from the same shared-memory tile (contraction over q)
#pragma unroll
                for (uint32_t kk = 0; kk < 8; ++kk)
                  mma_ss(tDK, smem_desc(DESC_KMAJOR, sDS + (kk >> 2) * kBoxBytes + (kk & 3) * 32), smem_desc(mn_hi, sQ + kk * 2048),
                         id_tm, (it > 0 || kk > 0) ? 1u : 0u);
                tc_commit(bar(kBarQEmpty + st));
                tc_commit(bar(kBarDSFree));
              }
            }
            __syncwarp();
            // dP_{it+1} once the dQ tile has been drained out of the shared TMEM region
            if (has_next) {
              FFA_TRB(0, it, 4);
              mbar_wait(bar(kBarDrained), it & 1);
              FFA_TRB(0, it, 5);
              mbar_wait(bar(kBarDOFull), (it + 1) & 1);
              tc_fence_after();
              FFA_TRB(0, it, 6);
              mma_kmajor(tDP, sV, smem_base + kSmemDO, bar(kBarDP));
            }
          }
          if (elect_one()) tc_commit(bar(kBarDKV));
          __syncwarp();
        }
      }
    } else if (wid == 13) {
      {
        // ================================ TMA producer (warp-uniform, elected lane issues) ================
        bool kv_loaded = false;
        uint32_t it = 0;
        for (int g = 0; g < G; ++g) {
          const int hq = hk * G + g;
          const float* lse2 = args.lse2 + int64_t(hq) * args.tq_ld;
          const float* delta = args.delta + int64_t(hq) * args.tq_ld;
          SliceCursor cur;
          cur.init(args.meta, args.plan, tile);
          Slice sl;
          BwdGeom gm;
          while (cur.next(sl)) {
            if (!gm.init(sl, C0)) continue;
            if (!kv_loaded) {
              if (elect_one()) {
                mbar_arrive_expect_tx(bar(kBarKV), 2 * kTileBytes);
#pragma unroll
                for (uint32_t bx = 0; bx < kBoxes; ++bx) {
                  tma_load_3d(smem_base + kSmemK + bx * kBoxBytes, &tmK, bar(kBarKV), 64 * bx, hk, C0);
                  tma_load_3d(smem_base + kSmemV + bx * kBoxBytes, &tmV, bar(kBarKV), 64 * bx, hk, C0);
                }
              }
              __syncwarp();
              kv_loaded = true;
            }
            for (int b = 0; b < gm.nblk; ++b, ++it) {
              const int rs = gm.r_start + b * kBlockM;
              const uint32_t st = it & 1, ph = (it >> 1) & 1;
              FFA_TRB(3, it, 0);
              mbar_wait(bar(kBarQEmpty + st), ph ^ 1);
              FFA_TRB(3, it, 1);
              if (elect_one()) {
                mbar_arrive_expect_tx(bar(kBarQFull + st), kTileBytes + 1024);
                const uint32_t sQ = smem_base + kSmemQ + st * kTileBytes;
#pragma unroll
                for (uint32_t bx = 0; bx < kBoxes; ++bx)
                  tma_load_3d(sQ + bx * kBoxBytes, &tmQ, bar(kBarQFull + st), 64 * bx, hq, rs);
                bulk_load(smem_base + kSmemLse + st * 512, lse2 + rs, 512, bar(kBarQFull + st));
                bulk_load(smem_base + kSmemDelta + st * 512, delta + rs, 512, bar(kBarQFull + st));
              }
              __syncwarp();
              mbar_wait(bar(kBarDOEmpty), (it & 1) ^ 1);
              FFA_TRB(3, it, 2);
              if (elect_one()) {
                mbar_arrive_expect_tx(bar(kBarDOFull), kTileBytes);
                const uint32_t sDO = smem_base + kSmemDO;
#pragma unroll
                for (uint32_t bx = 0; bx < kBoxes; ++bx)
                  tma_load_3d(sDO + bx * kBoxBytes, &tmDO, bar(kBarDOFull), 64 * bx, hq, rs);
              }
              __syncwarp();
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 12) tmem_dealloc(tmem, 512);
}

// -------------------------------------------------------------------------------------------------
// host launchers
// -------------------------------------------------------------------------------------------------
cudaError_t launch_bwd_preprocess(int head_dim, bool bf16, const void* out, const void* dout, const float* lse, float* delta,
                                  float* lse2, int total_q, int num_heads, int tq_ld, int64_t o_stride_t,
                                  int64_t o_stride_h, int64_t do_stride_t, int64_t do_stride_h, int64_t lse_stride_t,
                                  int64_t lse_stride_h, cudaStream_t stream) {
  const int64_t total = int64_t(num_heads) * tq_ld;
  const int wpb = 8;
  const unsigned blocks = unsigned((total + wpb - 1) / wpb);
  if (blocks == 0) return cudaSuccess;
#define FFA_PRE(D_, B_)                                                                                                   \
  ffa_bwd_preprocess_kernel<D_, B_><<<blocks, wpb * 32, 0, stream>>>(out, dout, lse, delta, lse2, total_q, num_heads, tq_ld, \
                                                                     o_stride_t, o_stride_h, do_stride_t, do_stride_h,      \
                                                                     lse_stride_t, lse_stride_h)
  if (head_dim == 64) {
    if (bf16) FFA_PRE(64, true); else FFA_PRE(64, false);
  } else {
    if (bf16) FFA_PRE(128, true); else FFA_PRE(128, false);
  }
#undef FFA_PRE
  return cudaGetLastError();
}

template <int kD, bool kBF16, bool kDkvF32>
static cudaError_t launch_bwd_one(const CUtensorMap& tmQ, const CUtensorMap& tmDO, const CUtensorMap& tmK,
                                  const CUtensorMap& tmV, const CUtensorMap& tmDQ, const BwdArgs& a, int grid,
                                  cudaStream_t stream) {
  auto kern = ffa_bwd_kernel<kD, kBF16, kDkvF32>;
  constexpr uint32_t smem_bytes = bwd::Cfg<kD>::kSmemBytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  kern<<<grid, bwd::kThreads, smem_bytes, stream>>>(tmQ, tmDO, tmK, tmV, tmDQ, a);
  return cudaGetLastError();
}

template <int kD>
static cudaError_t launch_bwd_d(bool bf16, bool dkv_f32, const CUtensorMap& tmQ, const CUtensorMap& tmDO, const CUtensorMap& tmK,
                                const CUtensorMap& tmV, const CUtensorMap& tmDQ, const BwdArgs& a, int grid, cudaStream_t stream) {
  if (bf16) return dkv_f32 ? launch_bwd_one<kD, true, true>(tmQ, tmDO, tmK, tmV, tmDQ, a, grid, stream)
                           : launch_bwd_one<kD, true, false>(tmQ, tmDO, tmK, tmV, tmDQ, a, grid, stream);
  return dkv_f32 ? launch_bwd_one<kD, false, true>(tmQ, tmDO, tmK, tmV, tmDQ, a, grid, stream)
                 : launch_bwd_one<kD, false, false>(tmQ, tmDO, tmK, tmV, tmDQ, a, grid, stream);
}

cudaError_t launch_bwd(int head_dim, bool bf16, bool dkv_f32, const CUtensorMap& tmQ, const CUtensorMap& tmDO,
                       const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmDQ, const BwdArgs& a,
                       cudaStream_t stream) {
  const int grid = a.n_tiles * a.num_heads_kv;
  if (grid == 0) return cudaSuccess;
  return head_dim == 64 ? launch_bwd_d<64>(bf16, dkv_f32, tmQ, tmDO, tmK, tmV, tmDQ, a, grid, stream)
                        : launch_bwd_d<128>(bf16, dkv_f32, tmQ, tmDO, tmK, tmV, tmDQ, a, grid, stream);
}

}  // namespace ffa
