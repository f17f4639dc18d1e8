// ffa_fwd64.cu — forward variant with 64-column kv blocks and DOUBLE-BUFFERED S (see ffa_fwd.cu for the
// common design notes).  Differences from the 128-column kernel:
//   * TMEM (512 cols): S[tile][2] 4 x 64 cols at [0,256), O0/O1 at [256,512); P (16-bit) overwrites the first 32
//     columns of the S buffer it came from.
//   * the MMA warp issues S_t(j+1) into the other buffer BEFORE it waits for P_t(j): the softmax warps find their
//     next block already computed and run back to back (throughput-bound instead of round-trip-latency-bound);
//     per kv block the issue order is  S0(cur) PV0(prev)  S1(cur) PV1(prev).
//   * a per-tile "PV retired" barrier guards the (rare) O rescale, because S(cur) no longer implies PV(prev) done.
//   * K/V tiles are 64 x 128 (16 KB) in a 10-deep ring; block-level mask classification is twice as fine.
#include <cuda.h>
#include <cmath>
#include <type_traits>
#include "ffa_args.h"
#include "ffa_common.cuh"
#include "ffa_kernels.h"
#include "sm100.cuh"

namespace ffa {
using namespace sm100;

namespace fwd64 {
constexpr int kThreads = 384;
constexpr int kBN = 64;                          // kv columns per block
constexpr int kStages = 10;
constexpr uint32_t kQTileBytes = 128 * 128 * 2;  // one 128x128 Q tile (2 boxes of 16 KB)
constexpr uint32_t kQBoxBytes = kQTileBytes / 2;
constexpr uint32_t kTileBytes = kBN * 128 * 2;   // one 64x128 K or V tile (2 boxes of 8 KB)
constexpr uint32_t kBoxBytes = kTileBytes / 2;
constexpr uint32_t kSmemQ = 0;
constexpr uint32_t kSmemKV = 2 * kQTileBytes;
constexpr uint32_t kSmemBar = kSmemKV + kStages * kTileBytes;
constexpr uint32_t kSmemBytes = kSmemBar + 512;
constexpr int kBarQ = 0, kBarFull = 1, kBarEmpty = kBarFull + kStages, kBarS = kBarEmpty + kStages,  // S[t*2+b]
              kBarP = kBarS + 4, kBarPV = kBarP + 4, kNumBars = kBarPV + 2;
static_assert(kNumBars * 8 + 8 <= 512, "barrier area");
constexpr uint32_t kTmemS = 0, kTmemO = 256;
constexpr float kRescaleThreshold = 8.0f;
}  // namespace fwd64

namespace {
struct CtaGeom64 {
  int c_min, nblk;
  __device__ __forceinline__ bool init(const Slice& s, int R0) {
    int r_lo = max(s.q0, R0), r_hi = min(s.q1, R0 + kFwdTileQ);
    int c_max;
    if (!col_hull(s, r_lo, r_hi, c_min, c_max)) return false;
    nblk = (c_max - c_min + fwd64::kBN - 1) / fwd64::kBN;
    return true;
  }
};
struct TileGeom64 {
  int ct_min, ct_max, full_lo, full_hi;
  bool act, rows_full;
  __device__ __forceinline__ void init(const Slice& s, int t0) {
    int a = max(s.q0, t0), b = min(s.q1, t0 + kBlockM);
    ct_min = ct_max = full_lo = full_hi = 0;
    act = col_hull(s, a, b, ct_min, ct_max);
    rows_full = (a == t0) && (b == t0 + kBlockM);
    if (act) {
      full_lo = s.lo(b - 1);
      full_hi = min(s.hi(a), s.k1);
    }
  }
  __device__ __forceinline__ bool active(int bs) const { return act && (bs + fwd64::kBN > ct_min) && (bs < ct_max); }
  __device__ __forceinline__ bool full(int bs) const { return rows_full && full_lo <= bs && bs + fwd64::kBN <= full_hi; }
};
}  // namespace

template <bool kBF16, bool kOutF32>
__global__ void __launch_bounds__(fwd64::kThreads, 1)
ffa_fwd64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const FwdArgs args) {
  using namespace fwd64;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmemBar + kNumBars * 8);
  const uint32_t bar0 = smem_base + kSmemBar;
  auto bar = [&](int i) -> uint32_t { return bar0 + uint32_t(i) * 8u; };

  const int tid = threadIdx.x;
  const int wid = tid >> 5;
  const int lane = tid & 31;

  const int G = args.num_heads_q / args.num_heads_kv;
  const int per_kvh = args.n_tiles * G;
  const int hk = blockIdx.x / per_kvh;
  const int rem = blockIdx.x - hk * per_kvh;
  const int tile = args.n_tiles - 1 - rem / G;
  const int hq = hk * G + rem % G;
  const int R0 = tile * kFwdTileQ;

  if (tid == 0) {
    if (smem_base & 1023u) __trap();
    mbar_init(bar(kBarQ), 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar(kBarFull + i), 1);
      mbar_init(bar(kBarEmpty + i), 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar(kBarS + i), 1);
      mbar_init(bar(kBarP + i), 128);
    }
    mbar_init(bar(kBarPV + 0), 1);
    mbar_init(bar(kBarPV + 1), 1);
    fence_barrier_init();
  }
  if (wid == 8) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  if (wid == 9 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (wid < 8) {
    // =========================== softmax warpgroups ===========================
    setmaxnreg_inc<208>();
    const int t = wid >> 2;
    const int row_in_tile = (wid & 3) * 32 + lane;
    const int row = R0 + t * kBlockM + row_in_tile;
    const uint32_t lane_off = uint32_t((wid & 3) * 32) << 16;
    const uint32_t tS0 = tmem + lane_off + kTmemS + t * 128;  // buffer b at + b*64
    const uint32_t tO = tmem + lane_off + kTmemO + t * 128;
    const float sl2 = args.scale * 1.4426950408889634f;

    float m_ref = -INFINITY, l_sum = 0.f;
    uint32_t n_act = 0;
    SliceCursor cur;
    cur.init(args.meta, args.plan, tile);
    Slice sl;
    CtaGeom64 g;
    TileGeom64 tg;
    while (cur.next(sl)) {
      if (!g.init(sl, R0)) continue;
      tg.init(sl, R0 + t * kBlockM);
      const bool row_in = (row >= sl.q0) && (row < sl.q1);
      const int lo_r = row_in ? sl.lo(row) : 0;
      const int hi_r = row_in ? min(sl.hi(row), sl.k1) : 0;
      for (int b = 0; b < g.nblk; ++b) {
        const int bs = g.c_min + b * kBN;
        if (!tg.active(bs)) continue;
        const uint32_t sb = n_act & 1;
        const uint32_t tS = tS0 + sb * 64;
        mbar_wait(bar(kBarS + t * 2 + sb), (n_act >> 1) & 1);
        tc_fence_after();
        uint32_t sr[64];
        tmem_ld32(tS, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
        tmem_ld32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
        tc_wait_ld();
        if (!tg.full(bs)) {
          const int lo_rel = lo_r - bs, hi_rel = hi_r - bs;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            int lo_c = min(max(lo_rel - c * 32, 0), 32), hi_c = min(max(hi_rel - c * 32, 0), 32);
            uint32_t m_hi = hi_c >= 32 ? 0xffffffffu : ((1u << hi_c) - 1u);
            uint32_t m_lo = lo_c >= 32 ? 0xffffffffu : ((1u << lo_c) - 1u);
            uint32_t bits = m_hi & ~m_lo;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (!((bits >> i) & 1u)) sr[c * 32 + i] = 0xff800000u;  // -inf
          }
        }
        float mb;
        {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 64; i += 8) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              m4[j] = fmaxf(m4[j], fmaxf(__uint_as_float(sr[i + 2 * j]), __uint_as_float(sr[i + 2 * j + 1])));
          }
          mb = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        }
        const float m_new = fmaxf(m_ref, mb);
        const bool upd = (m_new - m_ref) * sl2 > kRescaleThreshold;  // NaN (-inf - -inf) compares false
        if (__any_sync(0xffffffffu, upd)) {
          float corr = 1.f;
          if (upd) {
            corr = (m_ref == -INFINITY) ? 0.f : ex2((m_ref - m_new) * sl2);
            m_ref = m_new;
            l_sum *= corr;
          }
          if (n_act > 0) {
            // O may only be touched once the previous PV of this tile has retired
            mbar_wait(bar(kBarPV + t), (n_act - 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t o[32];
              tmem_ld32(tO + c * 32, o);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
              tmem_st32(tO + c * 32, o);
            }
          }
        }
        const float mneg = (m_ref == -INFINITY) ? 0.f : -m_ref * sl2;
        const float2 sl2_2 = make_float2(sl2, sl2), mneg2 = make_float2(mneg, mneg);
        float2 acc4[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t pr[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(sr[32 * c + 2 * i]), __uint_as_float(sr[32 * c + 2 * i + 1])),
                                        sl2_2, mneg2);
            float2 pv;
            pv.x = ex2(x.x);
            pv.y = ex2(x.y);
            acc4[i & 3] = __fadd2_rn(acc4[i & 3], pv);
            pr[i] = pack2<kBF16>(pv.x, pv.y);
          }
          tmem_st16(tS + c * 16, pr);
        }
        {
          const float2 a01 = __fadd2_rn(acc4[0], acc4[1]), a23 = __fadd2_rn(acc4[2], acc4[3]);
          l_sum += (a01.x + a01.y) + (a23.x + a23.y);
        }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(bar(kBarP + t * 2 + sb));
        ++n_act;
      }
    }

    // ------------------------------- epilogue -------------------------------
    const bool row_ok = row < args.total_q;
    const bool dead = (n_act == 0) || (m_ref == -INFINITY) || !(l_sum > 0.f);
    float lse_new = dead ? -INFINITY : fmaf(m_ref, args.scale, __logf(l_sum));
    float w_new = dead ? 0.f : 1.f / l_sum;
    float w_old = 0.f;
    float* lse_ptr = args.lse + int64_t(row) * args.lse_stride_t + int64_t(hq) * args.lse_stride_h;
    bool need_old = false;
    if (args.accumulate && row_ok) {
      const float lse_old = *lse_ptr;
      if (lse_old != -INFINITY) {
        need_old = true;
        if (dead) {
          w_old = 1.f;
          lse_new = lse_old;
        } else {
          const float mx = fmaxf(lse_old, lse_new), mn = fminf(lse_old, lse_new);
          const float lse_m = mx + log1pf(__expf(mn - mx));
          w_old = __expf(lse_old - lse_m);
          w_new *= __expf(lse_new - lse_m);
          lse_new = lse_m;
        }
      }
    }
    if (n_act > 0) {
      mbar_wait(bar(kBarPV + t), (n_act - 1) & 1);  // the last PV of this tile has retired
      tc_fence_after();
    }
    using OutT = typename std::conditional<kOutF32, float, typename std::conditional<kBF16, __nv_bfloat16, __half>::type>::type;
    OutT* orow = reinterpret_cast<OutT*>(args.out) + int64_t(row) * args.o_stride_t + int64_t(hq) * args.o_stride_h;
    const bool skip_store = !row_ok || (args.accumulate && dead && need_old);  // nothing new: keep the old partial
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float o[32];
      if (n_act > 0) {
        uint32_t r[32];
        tmem_ld32(tO + c * 32, r);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(r[i]) * w_new;
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0.f;
      }
      if (dead) {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0.f;
      }
      if (!skip_store) {
        if constexpr (kOutF32) {
          float4* dst = reinterpret_cast<float4*>(orow + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 x = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            if (need_old) {
              float4 y = dst[i];
              x.x = fmaf(w_old, y.x, x.x); x.y = fmaf(w_old, y.y, x.y);
              x.z = fmaf(w_old, y.z, x.z); x.w = fmaf(w_old, y.w, x.w);
            }
            dst[i] = x;
          }
        } else {
          uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = o[8 * i + j];
            if (need_old) {
              uint4 y = dst[i];
              const uint32_t yw[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float lo, hi;
                if constexpr (kBF16) {
                  lo = __uint_as_float(yw[j] << 16);
                  hi = __uint_as_float(yw[j] & 0xffff0000u);
                } else {
                  __half2 h2 = *reinterpret_cast<const __half2*>(&yw[j]);
                  lo = __low2float(h2);
                  hi = __high2float(h2);
                }
                x[2 * j] = fmaf(w_old, lo, x[2 * j]);
                x[2 * j + 1] = fmaf(w_old, hi, x[2 * j + 1]);
              }
            }
            uint4 pk;
            pk.x = pack2<kBF16>(x[0], x[1]);
            pk.y = pack2<kBF16>(x[2], x[3]);
            pk.z = pack2<kBF16>(x[4], x[5]);
            pk.w = pack2<kBF16>(x[6], x[7]);
            dst[i] = pk;
          }
        }
      }
    }
    if (!skip_store) *lse_ptr = lse_new;
  } else {
    // register pool = 384 threads x 168 regs (launch bounds) = 64512 = 256 x 208 + 128 x 88
    setmaxnreg_dec<88>();
    if (wid == 8) {
      // =========================== MMA issuer (warp-uniform loop, elected lane issues) ===========================
      constexpr uint32_t idesc_s = make_idesc(kBF16, false, false, 128, 64);
      constexpr uint32_t idesc_pv = make_idesc(kBF16, false, true, 128, 128);
      constexpr uint64_t desc_v_hi = desc_mnmajor(kBoxBytes);
      const uint32_t sQ = smem_base + kSmemQ, sKV = smem_base + kSmemKV;
      uint32_t n_act0 = 0, n_act1 = 0;            // S issued so far per tile
      bool has_prev0 = false, has_prev1 = false;  // a PV of that tile is pending
      uint32_t prev_load0 = 0, prev_load1 = 0;    // ring load index of the V tile of the pending PV
      uint64_t pend = 0;                          // outstanding PV users per ring slot, 4 bits each
      uint32_t blk = 0;
      bool q_ready = false;

      auto issue_pv = [&](const int t, const uint32_t n_act_t, const uint32_t li) {
        // pending PV belongs to the (n_act_t-1)-th active block of tile t (0-based): buffer (n_act_t-1)&1
        const uint32_t j = n_act_t - 1, sb = j & 1;
        const uint32_t buf = li % kStages;
        mbar_wait(bar(kBarFull + buf), (li / kStages) & 1);
        mbar_wait(bar(kBarP + t * 2 + sb), (j >> 1) & 1);
        tc_fence_after();
        const uint32_t sV = sKV + buf * kTileBytes;
        const uint32_t tP = tmem + kTmemS + t * 128 + sb * 64, tOt = tmem + kTmemO + t * 128;
        pend -= uint64_t(1) << (4 * buf);
        const bool release = ((pend >> (4 * buf)) & 15u) == 0;
        if (elect_one()) {
#pragma unroll
          for (uint32_t kk = 0; kk < 4; ++kk)
            mma_ts(tOt, tP + kk * 8, smem_desc(desc_v_hi, sV + kk * 2048), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
          tc_commit(bar(kBarPV + t));
          if (release) tc_commit(bar(kBarEmpty + buf));
        }
        __syncwarp();
      };
      auto issue_s = [&](const int t, const uint32_t n_act_t, const uint32_t sK) {
        const uint32_t sb = n_act_t & 1;
        const uint32_t tSt = tmem + kTmemS + t * 128 + sb * 64;
        if (elect_one()) {
#pragma unroll
          for (uint32_t kk = 0; kk < 8; ++kk) {
            const uint32_t qoff = (kk >> 2) * kQBoxBytes + (kk & 3) * 32;
            const uint32_t koff = (kk >> 2) * kBoxBytes + (kk & 3) * 32;
            mma_ss(tSt, smem_desc(DESC_KMAJOR, sQ + t * kQTileBytes + qoff), smem_desc(DESC_KMAJOR, sK + koff), idesc_s,
                   kk > 0 ? 1u : 0u);
          }
          tc_commit(bar(kBarS + t * 2 + sb));
        }
        __syncwarp();
      };

      SliceCursor cur;
      cur.init(args.meta, args.plan, tile);
      Slice sl;
      CtaGeom64 g;
      TileGeom64 tg0, tg1;
      while (cur.next(sl)) {
        if (!g.init(sl, R0)) continue;
        tg0.init(sl, R0);
        tg1.init(sl, R0 + kBlockM);
        for (int b = 0; b < g.nblk; ++b, ++blk) {
          const int bs = g.c_min + b * kBN;
          const bool a0 = tg0.active(bs), a1 = tg1.active(bs);
          if (!q_ready) {
            mbar_wait(bar(kBarQ), 0);
            q_ready = true;
          }
          const uint32_t lk = 2 * blk, kbuf = lk % kStages, lv = lk + 1, vbuf = lv % kStages;
          mbar_wait(bar(kBarFull + kbuf), (lk / kStages) & 1);
          tc_fence_after();
          pend = (pend & ~(uint64_t(15) << (4 * vbuf))) | (uint64_t((a0 ? 1u : 0u) + (a1 ? 1u : 0u)) << (4 * vbuf));
          const uint32_t sK = sKV + kbuf * kTileBytes;
          // S(cur) first (other S buffer), then the PV of the previous block of the same tile
          if (a0) issue_s(0, n_act0, sK);
          if (has_prev0) { issue_pv(0, n_act0, prev_load0); has_prev0 = false; }
          if (a0) { has_prev0 = true; prev_load0 = lv; ++n_act0; }
          if (a1) issue_s(1, n_act1, sK);
          if (has_prev1) { issue_pv(1, n_act1, prev_load1); has_prev1 = false; }
          if (a1) { has_prev1 = true; prev_load1 = lv; ++n_act1; }
          if (elect_one()) {
            tc_commit(bar(kBarEmpty + kbuf));
            if (((pend >> (4 * vbuf)) & 15u) == 0) tc_commit(bar(kBarEmpty + vbuf));
          }
          __syncwarp();
        }
      }
      if (has_prev0) issue_pv(0, n_act0, prev_load0);
      if (has_prev1) issue_pv(1, n_act1, prev_load1);
    } else if (wid == 9) {
      // =========================== TMA producer ===========================
      const uint32_t sQ = smem_base + kSmemQ, sKV = smem_base + kSmemKV;
      uint32_t blk = 0;
      bool q_loaded = false;
      SliceCursor cur;
      cur.init(args.meta, args.plan, tile);
      Slice sl;
      CtaGeom64 g;
      while (cur.next(sl)) {
        if (!g.init(sl, R0)) continue;
        if (!q_loaded) {
          if (elect_one()) {
            mbar_arrive_expect_tx(bar(kBarQ), 2 * kQTileBytes);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              tma_load_3d(sQ + t * kQTileBytes, &tmQ, bar(kBarQ), 0, hq, R0 + t * kBlockM);
              tma_load_3d(sQ + t * kQTileBytes + kQBoxBytes, &tmQ, bar(kBarQ), 64, hq, R0 + t * kBlockM);
            }
          }
          __syncwarp();
          q_loaded = true;
        }
        for (int b = 0; b < g.nblk; ++b, ++blk) {
          const int bs = g.c_min + b * kBN;
#pragma unroll
          for (int which = 0; which < 2; ++which) {
            const uint32_t li = 2 * blk + which, buf = li % kStages;
            mbar_wait(bar(kBarEmpty + buf), ((li / kStages) & 1) ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(bar(kBarFull + buf), kTileBytes);
              const CUtensorMap* tm = which == 0 ? &tmK : &tmV;
              tma_load_3d(sKV + buf * kTileBytes, tm, bar(kBarFull + buf), 0, hk, bs);
              tma_load_3d(sKV + buf * kTileBytes + kBoxBytes, tm, bar(kBarFull + buf), 64, hk, bs);
            }
            __syncwarp();
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 8) tmem_dealloc(tmem, 512);
}

template <bool kBF16, bool kOutF32>
static cudaError_t launch_one64(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const FwdArgs& a,
                                int grid, cudaStream_t stream) {
  auto kern = ffa_fwd64_kernel<kBF16, kOutF32>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd64::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  kern<<<grid, fwd64::kThreads, fwd64::kSmemBytes, stream>>>(tmQ, tmK, tmV, a);
  return cudaGetLastError();
}

// tmK / tmV must be encoded with 64-row boxes for this variant
cudaError_t launch_fwd64(bool bf16, bool out_f32, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                         const FwdArgs& a, cudaStream_t stream) {
  const int grid = a.n_tiles * a.num_heads_q;
  if (grid == 0) return cudaSuccess;
  if (bf16) return out_f32 ? launch_one64<true, true>(tmQ, tmK, tmV, a, grid, stream)
                           : launch_one64<true, false>(tmQ, tmK, tmV, a, grid, stream);
  return out_f32 ? launch_one64<false, true>(tmQ, tmK, tmV, a, grid, stream)
                 : launch_one64<false, false>(tmQ, tmK, tmV, a, grid, stream);
}

}  // namespace ffa
