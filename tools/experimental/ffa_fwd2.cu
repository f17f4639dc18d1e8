// ffa_fwd2.cu — flex-flash-attention FORWARD, CTA-PAIR variant (tcgen05.mma.cta_group::2), head_dim 128.
//
// EXPERIMENTAL / OPT-IN (B200_FFA_FWD_VARIANT=2): written after the round-1 GPU budget was spent, NOT YET RUN ON HARDWARE.
// The CTA-pair UMMA forms it relies on (SS and TS, M = 256, B split by N halves, multicast commit) are validated by
// tools/probe/probe2cta.cu (profiles/r01_probe2cta.txt); the cross-CTA barrier protocol below is not.
//
// Same algorithm, warp roles, TMEM map and softmax / epilogue code as ffa_fwd.cu (this file is derived from it); what changes:
//  * a cluster of two CTAs owns 512 q rows of one head: CTA r holds rows [512p + 256r, +256) = its two 128-row MMA tiles;
//  * every MMA is a cta_group::2 MMA with M = 256 issued by the leader CTA (rank 0): S_t of BOTH CTAs' tile t in one
//    instruction stream, PV_t likewise.  The B operands are split across the pair — CTA r stages kv rows [64r, 64r+64) of the
//    K block (B of S = Q K^T, N = kv) and head-dim columns [64r, 64r+64) of the V block (B of O = P V, N = d) — so each CTA
//    fills and reads HALF of every K/V block: shared-memory traffic per 2048 tensor clocks drops from 256 KB to 160 KB
//    (DESIGN.md §10);
//  * barriers: TMA loads of both CTAs complete on the LEADER's full barriers (cp.async.bulk.tensor.cta_group::2), MMA
//    completions are multicast to both CTAs (S ready, ring slot empty, O ready), the P-ready barriers live in the leader
//    and collect 256 arrivals (128 local + 128 remote softmax threads);
//  * both CTAs walk the same kv-block sequence (slice list of the 512-row pair tile); a tile that has nothing in a block its
//    peer needs still runs the block with every column masked (P = 0), which only happens next to mask boundaries.
#include <cuda.h>
#include <cmath>
#include <type_traits>
#include "ffa_common.cuh"
#include "ffa_args.h"
#include "ffa_kernels.h"
#include "sm100.cuh"

namespace ffa {
using namespace sm100;

namespace fwd2 {
constexpr int kThreads = 384;
constexpr int kStages = 8;     // ring entries of kHalfBytes (K half, V half alternate)
constexpr int kPairTileQ = 512;  // q rows per CTA pair
constexpr uint32_t kHalfBytes = 16384;  // half a K block (64 rows x 128) or half a V block (128 rows x 64 columns)
constexpr uint32_t kBoxBytes = 128 * 64 * 2;  // one TMA box: 128 rows x 64 columns of 16-bit elements
// smem map: [Q0 Q1 (128 x 128 each)][ring of kStages half blocks][barriers]
template <int kD> struct Cfg {
  static_assert(kD == 128, "the CTA-pair forward exists for head_dim 128 only");
  static constexpr uint32_t kBoxes = kD / 64;
  static constexpr uint32_t kTileBytes = kBoxes * kBoxBytes;
  static constexpr uint32_t kSmemQ = 0;
  static constexpr uint32_t kSmemKV = 2 * kTileBytes;
  static constexpr uint32_t kSmemBar = kSmemKV + kStages * kHalfBytes;
  static constexpr uint32_t kSmemBytes = kSmemBar + 256;
};
// barrier indices
constexpr int kBarQ = 0, kBarFull = 1, kBarEmpty = kBarFull + kStages, kBarS = kBarEmpty + kStages, kBarP = kBarS + 2,
              kBarPhi = kBarP + 2, kBarO = kBarPhi + 2, kNumBars = kBarO + 2;  // kBarP: first half of P, kBarPhi: second
static_assert(kNumBars * 8 + 8 <= 256, "barrier area");
constexpr uint32_t kTmemS = 0, kTmemO = 256;
constexpr float kRescaleThreshold = 8.0f;
// exp2 on the FMA pipe for kEmuCount of every kEmuPeriod element pairs (rest on the SFU).  Measured on B200
// (causal 32K x 8 heads, after the issue-loop fixes): 0/8 -> 1288, 1/8 -> 1258, 2/8 -> 1231, 3/8 -> 1191, 4/8 -> 1105
// TFLOP/s: with packed FFMA2/FADD2 the kernel is issue/latency-bound, not MUFU-bound, so the default is 0.
#ifndef FFA_EMU_COUNT
#define FFA_EMU_COUNT 0
#endif
constexpr int kEmuPeriod = 8, kEmuCount = FFA_EMU_COUNT, kEmuDegree = 4;
}  // namespace fwd2


// Per-slice geometry for one CTA pair (rows [R0, R0+512)) and for one 128-row MMA tile.
struct PairGeom {
  int c_min, nblk;
  __device__ __forceinline__ bool init(const Slice& s, int R0) {
    int r_lo = max(s.q0, R0), r_hi = min(s.q1, R0 + fwd2::kPairTileQ);
    int c_max;
    if (!col_hull(s, r_lo, r_hi, c_min, c_max)) return false;
    nblk = (c_max - c_min + kBlockN - 1) / kBlockN;
    return true;
  }
};
struct TileGeom2 {
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
  __device__ __forceinline__ bool active(int bs) const { return act && (bs + kBlockN > ct_min) && (bs < ct_max); }
  __device__ __forceinline__ bool full(int bs) const { return rows_full && full_lo <= bs && bs + kBlockN <= full_hi; }
};

// exp2 of 64 scores of one row (stale-or-fresh reference max folded into `mneg`), packed 16-bit P written to
// 32 TMEM columns starting at `t_dst`; returns the two partial row sums.  kEmuCount of every kEmuPeriod pairs go
// through the FMA-pipe polynomial, the rest through MUFU.EX2.
template <bool kBF16>
__device__ __forceinline__ float2 softmax_half2(const uint32_t* sr, float sl2, float mneg, uint32_t t_dst) {
  using namespace fwd2;
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
      if ((i % kEmuPeriod) < kEmuCount) {
        pv = ex2_emu2<kEmuDegree>(x);
      } else {
        pv.x = ex2(x.x);
        pv.y = ex2(x.y);
      }
      acc4[i & 3] = __fadd2_rn(acc4[i & 3], pv);
      pr[i] = pack2<kBF16>(pv.x, pv.y);
    }
    tmem_st16(t_dst + c * 16, pr);
  }
  const float2 a01 = __fadd2_rn(acc4[0], acc4[1]), a23 = __fadd2_rn(acc4[2], acc4[3]);
  return __fadd2_rn(a01, a23);
}

template <int kD, bool kBF16, bool kOutF32, bool kMaxLogits>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(fwd2::kThreads, 1)
ffa_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const FwdArgs args) {
  using namespace fwd2;
  constexpr uint32_t kTileBytes = Cfg<kD>::kTileBytes, kBoxes = Cfg<kD>::kBoxes;
  constexpr uint32_t kSmemQ = Cfg<kD>::kSmemQ, kSmemKV = Cfg<kD>::kSmemKV, kSmemBar = Cfg<kD>::kSmemBar;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmemBar + kNumBars * 8);
  const uint32_t bar0 = smem_base + kSmemBar;
  auto bar = [&](int i) -> uint32_t { return bar0 + uint32_t(i) * 8u; };

  const int tid = threadIdx.x;
  const int wid = tid >> 5;
  const int lane = tid & 31;

  // ---- work decode: kv-head major, then tile (heaviest = highest index first), then q head in group
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int G = args.num_heads_q / args.num_heads_kv;
  const int n_pairs = (args.n_tiles + 1) / 2;
  const int per_kvh = n_pairs * G;
  const int cid = blockIdx.x >> 1;
  const int hk = cid / per_kvh;
  const int rem = cid - hk * per_kvh;
  const int pair = n_pairs - 1 - rem / G;  // plan tile (512 rows)
  const int hq = hk * G + rem % G;
  const int Rp = pair * kPairTileQ;               // first row of the pair
  const int R0 = Rp + int(rank) * kFwdTileQ;      // first row of this CTA (may lie beyond total_q: nothing stored then)
  const int R0_peer = Rp + int(rank ^ 1u) * kFwdTileQ;

  if (tid == 0) {
    if (smem_base & 1023u) {
      printf("ffa_fwd: dynamic smem not 1024-aligned\n");
      __trap();
    }
    mbar_init(bar(kBarQ), 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar(kBarFull + i), 1);
      mbar_init(bar(kBarEmpty + i), 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar(kBarS + t), 1);
      mbar_init(bar(kBarP + t), 256);    // 128 softmax threads of each CTA (only the leader's copy is used)
      mbar_init(bar(kBarPhi + t), 256);
      mbar_init(bar(kBarO + t), 1);
    }
    fence_barrier_init();
  }
  if (wid == 8) {
    tmem_alloc2(smem_u32(tmem_slot), 512);  // one warp of each CTA of the pair
    tmem_relinquish2();
  }
  if (wid == 9 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before any remote arrive / cross-CTA TMA completion
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (wid < 8) {
    // =========================== softmax warpgroups ===========================
    setmaxnreg_inc<208>();
    const int t = wid >> 2;                        // MMA tile 0/1
    const int row_in_tile = (wid & 3) * 32 + lane;  // TMEM lane
    const int row = R0 + t * kBlockM + row_in_tile;
    const uint32_t lane_off = uint32_t((wid & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_off + kTmemS + t * 128;
    const uint32_t tO = tmem + lane_off + kTmemO + t * 128;
    const float sl2 = args.scale * 1.4426950408889634f;

    float m_ref = -INFINITY, l_sum = 0.f;
    float m_true = -INFINITY;  // exact running max of the raw logits (only consumed by max_logits)
    uint32_t n_act = 0;
    const uint32_t barP_leader = mapa_shared(bar(kBarP + t), 0), barPhi_leader = mapa_shared(bar(kBarPhi + t), 0);
    SliceCursor cur;
    cur.init(args.meta, args.plan, pair);
    Slice sl;
    PairGeom g;
    TileGeom2 tg, tgp;  // own tile, the peer CTA's tile with the same index
    while (cur.next(sl)) {
      if (!g.init(sl, Rp)) continue;
      tg.init(sl, R0 + t * kBlockM);
      tgp.init(sl, R0_peer + t * kBlockM);
      const bool row_in = (row >= sl.q0) && (row < sl.q1);
      const int lo_r = row_in ? sl.lo(row) : 0;
      const int hi_r = row_in ? min(sl.hi(row), sl.k1) : 0;  // empty range when row not in slice
      for (int b = 0; b < g.nblk; ++b) {
        const int bs = g.c_min + b * kBlockN;
        if (!(tg.active(bs) || tgp.active(bs))) continue;  // the pair's MMA for tile index t covers both CTAs
        mbar_wait(bar(kBarS + t), n_act & 1);
        tc_fence_after();
        uint32_t sr[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[c * 32]));
        tc_wait_ld();
        if (!tg.full(bs)) {
          const int lo_rel = lo_r - bs, hi_rel = hi_r - bs;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            int lo_c = min(max(lo_rel - c * 32, 0), 32), hi_c = min(max(hi_rel - c * 32, 0), 32);
            uint32_t m_hi = hi_c >= 32 ? 0xffffffffu : ((1u << hi_c) - 1u);
            uint32_t m_lo = lo_c >= 32 ? 0xffffffffu : ((1u << lo_c) - 1u);
            uint32_t bits = m_hi & ~m_lo;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (!((bits >> i) & 1u)) sr[c * 32 + i] = 0xff800000u;  // -inf
          }
        }
        // Block max on 4 independent chains.  Nothing below depends on it until the validity check, so the
        // scheduler overlaps it with the first half of the exponentials, which use the STALE reference max:
        // with lazy rescaling the reference only has to be within 2^8 of the true max, which the check verifies.
        float mb;
        {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 128; i += 8) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              m4[j] = fmaxf(m4[j], fmaxf(__uint_as_float(sr[i + 2 * j]), __uint_as_float(sr[i + 2 * j + 1])));
          }
          mb = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        }
        if constexpr (kMaxLogits) m_true = fmaxf(m_true, mb);
        float mneg = (m_ref == -INFINITY) ? 0.f : -m_ref * sl2;
        float2 acc_lo = softmax_half2<kBF16>(&sr[0], sl2, mneg, tS);
        const float m_new = fmaxf(m_ref, mb);
        const bool upd = (m_new - m_ref) * sl2 > kRescaleThreshold;  // NaN (-inf - -inf) compares false
        if (__any_sync(0xffffffffu, upd)) {
          // slow path (first block of a tile, or the running max grew by more than 2^8): move the reference,
          // rescale O (PV of the previous block has completed: its commit preceded S(cur)'s), redo the half.
          float corr = 1.f;
          if (upd) {
            corr = (m_ref == -INFINITY) ? 0.f : ex2((m_ref - m_new) * sl2);
            m_ref = m_new;
            l_sum *= corr;
          }
          if (n_act > 0) {
#pragma unroll
            for (int c = 0; c < kD / 32; ++c) {
              uint32_t o[32];
              tmem_ld32(tO + c * 32, o);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
              tmem_st32(tO + c * 32, o);
            }
          }
          mneg = (m_ref == -INFINITY) ? 0.f : -m_ref * sl2;
          acc_lo = softmax_half2<kBF16>(&sr[0], sl2, mneg, tS);
        }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive_cluster(barP_leader);  // P[:, 0:64) (and any O rescale) visible: first 4 PV k-steps may start
        const float2 acc_hi = softmax_half2<kBF16>(&sr[64], sl2, mneg, tS + 32);
        l_sum += (acc_lo.x + acc_lo.y) + (acc_hi.x + acc_hi.y);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive_cluster(barPhi_leader);
        ++n_act;
      }
    }

    // ------------------------------- epilogue -------------------------------
    if constexpr (kMaxLogits) {
      // per-head max of the scaled logits over every unmasked (q, k) pair (reference: max_logits of
      // _flex_flash_attn_forward, flex_flash_attn.py:397-403): warp max, then an ordered-int atomic max
      float mt = m_true;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
      if (lane == 0 && mt != -INFINITY) {
        const float val = mt * args.scale;
        if (val >= 0.f) atomicMax(reinterpret_cast<int*>(args.max_logits + hq), __float_as_int(val));
        else atomicMin(reinterpret_cast<unsigned int*>(args.max_logits + hq), __float_as_uint(val));
      }
    }
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
      mbar_wait(bar(kBarO + t), 0);
      tc_fence_after();
    }
    using OutT = typename std::conditional<kOutF32, float, typename std::conditional<kBF16, __nv_bfloat16, __half>::type>::type;
    OutT* orow = reinterpret_cast<OutT*>(args.out) + int64_t(row) * args.o_stride_t + int64_t(hq) * args.o_stride_h;
    const bool skip_store = !row_ok || (args.accumulate && dead && need_old);  // nothing new: keep the old partial
#pragma unroll
    for (int c = 0; c < kD / 32; ++c) {
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
    if (wid == 8 && rank == 0) {
      // =========================== MMA issuer (leader CTA only) ===========================
      // The whole warp runs this loop with warp-uniform control flow (descriptors live in uniform registers);
      // only the tcgen05.mma / tcgen05.commit instructions themselves are issued by one elected lane.
      constexpr uint32_t idesc_s = make_idesc(kBF16, false, false, 256, 128);   // M = 256: tile t of both CTAs
      constexpr uint32_t idesc_pv = make_idesc(kBF16, false, true, 256, kD);    // N = kD, 64 columns from each CTA
      constexpr uint64_t desc_v_hi = desc_mnmajor(kBoxBytes);
      const uint32_t sQ = smem_base + kSmemQ, sKV = smem_base + kSmemKV;
      uint32_t n_act0 = 0, n_act1 = 0;            // active blocks issued so far per tile
      bool has_prev0 = false, has_prev1 = false;  // a PV of that tile is pending
      uint32_t prev_load0 = 0, prev_load1 = 0;    // ring load index of the V block of the pending PV
      uint32_t pend = 0;                          // outstanding PV users per ring buffer, 4 bits each
      uint32_t blk = 0;                           // kv block counter (ring: K = 2*blk, V = 2*blk+1)
      bool q_ready = false;

      auto issue_pv = [&](const int t, const uint32_t n_act_t, const uint32_t li) {
        const uint32_t buf = li % kStages;
        const uint32_t sV = sKV + buf * kHalfBytes;  // [128 kv rows][64 head-dim columns] in each CTA
        const uint32_t tP = tmem + kTmemS + t * 128, tOt = tmem + kTmemO + t * 128;
        mbar_wait(bar(kBarFull + buf), (li / kStages) & 1);
        mbar_wait(bar(kBarP + t), (n_act_t - 1) & 1);  // first 64 columns of P_t are in TMEM
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (uint32_t kk = 0; kk < 4; ++kk)
            mma2_ts(tOt, tP + kk * 8, smem_desc(desc_v_hi, sV + kk * 2048), idesc_pv, (n_act_t > 1 || kk > 0) ? 1u : 0u);
        }
        __syncwarp();
        mbar_wait(bar(kBarPhi + t), (n_act_t - 1) & 1);  // second half
        tc_fence_after();
        pend -= 1u << (4 * buf);
        const bool release = ((pend >> (4 * buf)) & 15u) == 0;
        if (elect_one()) {
#pragma unroll
          for (uint32_t kk = 4; kk < 8; ++kk)
            mma2_ts(tOt, tP + kk * 8, smem_desc(desc_v_hi, sV + kk * 2048), idesc_pv, 1u);
          if (release) tc_commit2(bar(kBarEmpty + buf), 3);
        }
        __syncwarp();
      };
      auto issue_s = [&](const int t, const uint32_t sK) {
        const uint32_t tSt = tmem + kTmemS + t * 128;
        if (elect_one()) {
#pragma unroll
          for (uint32_t kk = 0; kk < kD / 16; ++kk) {  // contraction over the head dim
            const uint32_t qoff = (kk >> 2) * kBoxBytes + (kk & 3) * 32;      // Q tile: boxes of [128 rows][64 cols]
            const uint32_t koff = (kk >> 2) * (kBoxBytes / 2) + (kk & 3) * 32;  // K half: boxes of [64 rows][64 cols]
            mma2_ss(tSt, smem_desc(DESC_KMAJOR, sQ + t * kTileBytes + qoff), smem_desc(DESC_KMAJOR, sK + koff), idesc_s,
                    kk > 0 ? 1u : 0u);
          }
          tc_commit2(bar(kBarS + t), 3);
        }
        __syncwarp();
      };

      SliceCursor cur;
      cur.init(args.meta, args.plan, pair);
      Slice sl;
      PairGeom g;
      TileGeom2 tg00, tg01, tg10, tg11;  // [cta rank][tile]
      while (cur.next(sl)) {
        if (!g.init(sl, Rp)) continue;
        tg00.init(sl, Rp);
        tg01.init(sl, Rp + kBlockM);
        tg10.init(sl, Rp + kFwdTileQ);
        tg11.init(sl, Rp + kFwdTileQ + kBlockM);
        for (int b = 0; b < g.nblk; ++b, ++blk) {
          const int bs = g.c_min + b * kBlockN;
          // tile index t is issued when either CTA's tile t needs the block (the softmax warps apply the same rule)
          const bool a0 = tg00.active(bs) || tg10.active(bs), a1 = tg01.active(bs) || tg11.active(bs);
          if (!q_ready) {
            mbar_wait(bar(kBarQ), 0);
            q_ready = true;
          }
          const uint32_t lk = 2 * blk, kbuf = lk % kStages, lv = lk + 1, vbuf = lv % kStages;
          mbar_wait(bar(kBarFull + kbuf), (lk / kStages) & 1);
          tc_fence_after();
          pend = (pend & ~(15u << (4 * vbuf))) | (((a0 ? 1u : 0u) + (a1 ? 1u : 0u)) << (4 * vbuf));
          const uint32_t sK = sKV + kbuf * kHalfBytes;  // [64 kv rows][128] in each CTA
          if (has_prev0) { issue_pv(0, n_act0, prev_load0); has_prev0 = false; }
          if (a0) { issue_s(0, sK); has_prev0 = true; prev_load0 = lv; ++n_act0; }
          if (has_prev1) { issue_pv(1, n_act1, prev_load1); has_prev1 = false; }
          if (a1) { issue_s(1, sK); has_prev1 = true; prev_load1 = lv; ++n_act1; }
          if (elect_one()) {
            tc_commit2(bar(kBarEmpty + kbuf), 3);  // K halves free (in both CTAs) once both S MMAs retire
            if (((pend >> (4 * vbuf)) & 15u) == 0) tc_commit2(bar(kBarEmpty + vbuf), 3);  // V block nobody needs
          }
          __syncwarp();
        }
      }
      if (has_prev0) issue_pv(0, n_act0, prev_load0);
      if (has_prev1) issue_pv(1, n_act1, prev_load1);
      if (elect_one()) {
        if (n_act0 > 0) tc_commit2(bar(kBarO + 0), 3);
        if (n_act1 > 0) tc_commit2(bar(kBarO + 1), 3);
      }
      __syncwarp();
    } else if (wid == 9) {
      // =========================== TMA producer (both CTAs; warp-uniform loop, one elected lane issues) =====
      // every load of the pair completes on the LEADER's barrier; only the leader posts the expected byte count
      const uint32_t sQ = smem_base + kSmemQ, sKV = smem_base + kSmemKV;
      const uint32_t barQ_leader = mapa_shared(bar(kBarQ), 0);
      uint32_t blk = 0;
      bool q_loaded = false;
      SliceCursor cur;
      cur.init(args.meta, args.plan, pair);
      Slice sl;
      PairGeom g;
      while (cur.next(sl)) {
        if (!g.init(sl, Rp)) continue;
        if (!q_loaded) {
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(bar(kBarQ), 4 * kTileBytes);  // two Q tiles of each CTA
#pragma unroll
            for (int t = 0; t < 2; ++t) {
#pragma unroll
              for (uint32_t bx = 0; bx < kBoxes; ++bx)
                tma_load_3d_pair(sQ + t * kTileBytes + bx * kBoxBytes, &tmQ, barQ_leader, 64 * bx, hq, R0 + t * kBlockM);
            }
          }
          __syncwarp();
          q_loaded = true;
        }
        for (int b = 0; b < g.nblk; ++b, ++blk) {
          const int bs = g.c_min + b * kBlockN;
#pragma unroll
          for (int which = 0; which < 2; ++which) {
            const uint32_t li = 2 * blk + which, buf = li % kStages;
            mbar_wait(bar(kBarEmpty + buf), ((li / kStages) & 1) ^ 1);  // local: the leader's commits are multicast
            if (elect_one()) {
              const uint32_t full_leader = mapa_shared(bar(kBarFull + buf), 0);
              if (rank == 0) mbar_arrive_expect_tx(bar(kBarFull + buf), 2 * kHalfBytes);  // this CTA's half + the peer's
              const uint32_t dst = sKV + buf * kHalfBytes;
              if (which == 0) {
                // K: kv rows [bs + 64 rank, +64), all 128 head-dim columns = two [64 rows][64 cols] boxes (tmK: 64-row boxes)
#pragma unroll
                for (uint32_t bx = 0; bx < kBoxes; ++bx)
                  tma_load_3d_pair(dst + bx * (kBoxBytes / 2), &tmK, full_leader, 64 * bx, hk, bs + 64 * int(rank));
              } else {
                // V: all 128 kv rows, head-dim columns [64 rank, +64) = one [128 rows][64 cols] box
                tma_load_3d_pair(dst, &tmV, full_leader, 64 * int(rank), hk, bs);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA frees TMEM or exits while the pair's MMAs / multicast commits may still touch it
  if (wid == 8) tmem_dealloc2(tmem, 512);
}

// ---------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------
template <bool kBF16, bool kOutF32>
static cudaError_t launch_pair(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const FwdArgs& a,
                               int grid, cudaStream_t stream) {
  auto kern = ffa_fwd2_kernel<128, kBF16, kOutF32, false>;
  constexpr uint32_t smem_bytes = fwd2::Cfg<128>::kSmemBytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  kern<<<grid, fwd2::kThreads, smem_bytes, stream>>>(tmQ, tmK, tmV, a);  // __cluster_dims__(2,1,1): grid is even
  return cudaGetLastError();
}

// head_dim 128, no max_logits; the plan must have been built over 512-row tiles ((n_tiles + 1) / 2 of them) and the K map
// must use 64-row boxes.
cudaError_t launch_fwd_pair(bool bf16, bool out_f32, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                            const FwdArgs& a, cudaStream_t stream) {
  const int grid = 2 * ((a.n_tiles + 1) / 2) * a.num_heads_q;
  if (grid == 0) return cudaSuccess;
  if (bf16) return out_f32 ? launch_pair<true, true>(tmQ, tmK, tmV, a, grid, stream) : launch_pair<true, false>(tmQ, tmK, tmV, a, grid, stream);
  return out_f32 ? launch_pair<false, true>(tmQ, tmK, tmV, a, grid, stream) : launch_pair<false, false>(tmQ, tmK, tmV, a, grid, stream);
}

}  // namespace ffa
