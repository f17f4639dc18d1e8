// EXPERIMENT, NOT BUILT (tools/experimental): the forward kernel with the opt-in "split-row" layout (template flag kSplit:
// four softmax warpgroups = (q tile, column half), two threads per q row exchanging block maxima through shared memory, so
// that the fixed latencies of one warp hide behind the MUFU work of its sibling on the same SM sub-partition).  State at the
// end of round 2: compiles; after fixing the register-pool over-subscription (setmaxnreg only redistributes the launch
// allocation: 640 x 96 = 61440 registers = 512 x 104 + 128 x 64) the first three ladder cases pass and a later one (GQA /
// ragged) still hangs, with no GPU budget left to debug it.  The product kernel is magiattention_b200/csrc/ffa_fwd.cu.
// ffa_fwd.cu — flex-flash-attention FORWARD for sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces the reference's Hopper-only FFA forward (csrc/flexible_flash_attention/
// flash_fwd_kernel_sm90.h:189-455, mainloop_fwd_sm90_tma_gmma_ws.hpp, epilogue_fwd.hpp) and the FA4
// CuTe-DSL kernel it falls back to on Blackwell (flash_attn/cute/flash_fwd_sm100.py).  Design:
//
//  * One CTA owns a 256-row q tile of one head (two 128-row MMA tiles that share every K/V block) and
//    walks ALL slices touching those rows (tile->slice CSR from ffa_plan.cu).  Rows are therefore
//    reduced across slices in registers/TMEM: no global locks, no fp32 atomics on O, no LSE-merge pass.
//  * Warp roles (384 threads): warps 0-3 softmax of tile 0, warps 4-7 softmax of tile 1, warp 8 issues every
//    tcgen05.mma and warp 9 every TMA load (warp-uniform loops, the async instruction under elect.sync).  Softmax thread i owns row i of its
//    tile (TMEM lane i): row max / row sum need no shuffles.
//  * TMEM (512 cols): S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); P (16-bit) overwrites the
//    first 64 columns of its S.  MMA order per kv block:  PV0(prev) S0(cur) PV1(prev) S1(cur)  so the
//    softmax of one tile always overlaps two MMAs of the other.
//  * K and V blocks (128 x 128, two SWIZZLE_128B TMA boxes each) stream through a 5-deep ring; Q tiles
//    are loaded once.
//  * Online softmax in base 2 with lazy rescaling: the reference max only moves when it would grow by
//    more than 2^8, so the O correction (TMEM load-scale-store) is rare and done by the softmax warps
//    themselves; it is safe because S(cur) completing implies PV(prev) completed (in-order tensor pipe).
//    The first half of each row's exponentials runs on the stale reference while the block max is computed;
//    P is published in two halves so the first four PV k-steps can start early.  Packed FFMA2/FADD2 math.
//  * Epilogue normalises, optionally LSE-merges into a previous partial (out, lse) — the CP stage merge
//    the reference does with a separate Triton pass (functional/utils.py:467-558) — and stores rows.
#include <cuda.h>
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include "ffa_common.cuh"
#include "ffa_args.h"
#include "ffa_kernels.h"
#include "sm100.cuh"

// Timing-only ablation switches for kernel analysis (results become WRONG; never set in the product build):
//   1 exp2 -> fma   2 no block max   4 P not stored to TMEM   8 S not loaded from TMEM   16 K/V ring filled only once
//   32 no PV MMAs   64 no S MMAs
#ifndef FFA_FWD_ABL
#define FFA_FWD_ABL 0
#endif

// Developer timeline (-DFFA_TRACE): CTA 0 stamps %clock64 at the hand-off points of iterations [64, 128) into a device
// array that b200_debug_trace_fwd() copies out (tools/dev/trace_fwd.py prints it).  Compiled out of the product build.
#ifdef FFA_TRACE
__device__ unsigned long long g_trace_fwd[4][64][8];
#define FFA_TR(role, it, ev)                                                                                         \
  do {                                                                                                                \
    if (blockIdx.x == 0 && (it) >= 64u && (it) < 128u && (threadIdx.x & 31) == 0) g_trace_fwd[role][(it) - 64u][ev] = clock64(); \
  } while (0)
extern "C" __attribute__((visibility("default"))) int b200_debug_trace_fwd(unsigned long long* dst) {
  return cudaMemcpyFromSymbol(dst, g_trace_fwd, sizeof(g_trace_fwd)) == cudaSuccess ? 0 : -1;
}
#else
#define FFA_TR(role, it, ev) do {} while (0)
#endif

namespace ffa {
using namespace sm100;

namespace fwd {
// Two forward layouts share this file (template flag kSplit):
//   kSplit = false  384 threads: one softmax thread per q row (two warpgroups, one per 128-row tile), 5-deep K/V ring;
//   kSplit = true   640 threads: TWO softmax threads per q row (four warpgroups: tile x column half), each owning 64 of a
//                   kv block's 128 score columns; the halves exchange their block maxima through shared memory.  Two warps
//                   of the same tile then share every SM sub-partition, so the fixed latencies of one (TMEM load / store
//                   waits, barrier hand-offs) hide behind the MUFU work of the other.  4-deep K/V ring (pays for the
//                   exchange buffers).
constexpr int kThreads = 384;
constexpr int kThreadsSplit = 640;  // 16 softmax warps + MMA warp + TMA warp + 2 idle (setmaxnreg works on whole warpgroups)
constexpr int kStages = 5;
constexpr int kStagesSplit = 4;
constexpr uint32_t kBoxBytes = 128 * 64 * 2;  // one TMA box: 128 rows x 64 columns of 16-bit elements
// operand tiles are 128 rows x kD columns = kD/64 boxes; smem map: [Q0 Q1][K/V ring of kStages tiles][barriers]
template <int kD, bool kSplit = false> struct Cfg {
  static_assert(kD == 64 || kD == 128, "head_dim must be 64 or 128");
  static constexpr int kNumStages = kSplit ? kStagesSplit : kStages;
  static constexpr uint32_t kBoxes = kD / 64;
  static constexpr uint32_t kTileBytes = kBoxes * kBoxBytes;
  static constexpr uint32_t kSmemQ = 0;
  static constexpr uint32_t kSmemKV = 2 * kTileBytes;
  static constexpr uint32_t kSmemBar = kSmemKV + kNumStages * kTileBytes;
  static constexpr uint32_t kSmemXch = kSmemBar + 256;  // kSplit: [parity][tile][half][128] fp32 block maxima / row sums
  static constexpr uint32_t kSmemBytes = kSmemXch + (kSplit ? 2 * 2 * 2 * 128 * 4 : 0);
};
// barrier indices
// (laid out for the deeper ring; the 4-stage layout simply leaves two slots unused)
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
#ifndef FFA_EMU_DEGREE
#define FFA_EMU_DEGREE 3
#endif
constexpr int kEmuPeriod = 8, kEmuCount = FFA_EMU_COUNT, kEmuDegree = FFA_EMU_DEGREE;
}  // namespace fwd


// Per-slice geometry for one forward CTA tile (rows [R0, R0+256)) and for one of its two MMA tiles.
struct CtaGeom {
  int c_min, nblk;
  __device__ __forceinline__ bool init(const Slice& s, int R0) {
    int r_lo = max(s.q0, R0), r_hi = min(s.q1, R0 + kFwdTileQ);
    int c_max;
    if (!col_hull(s, r_lo, r_hi, c_min, c_max)) return false;
    nblk = (c_max - c_min + kBlockN - 1) / kBlockN;
    return true;
  }
};
struct TileGeom {
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

// exp2 of 32 scores of one row (stale-or-fresh reference max folded into `mneg`) -> 16 packed 16-bit P registers; returns
// the partial row sums (two lanes of a float2).  kEmuCount of every kEmuPeriod pairs go through the FMA-pipe polynomial
// (never in the row's last quarter: the emulated path has the longer latency), the rest through MUFU.EX2.
template <bool kBF16, bool kAllowEmu>
__device__ __forceinline__ float2 softmax_quarter(const uint32_t* sr, float sl2, float mneg, uint32_t (&pr)[16]) {
  using namespace fwd;
  const float2 sl2_2 = make_float2(sl2, sl2), mneg2 = make_float2(mneg, mneg);
  float2 acc4[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 x = __ffma2_rn(make_float2(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1])), sl2_2, mneg2);
    float2 pv;
    if (kAllowEmu && (i % kEmuPeriod) >= kEmuPeriod - kEmuCount) {
      pv = ex2_emu2<kEmuDegree>(x);
    } else if (FFA_FWD_ABL & 1) {
      pv = __ffma2_rn(x, make_float2(0.001f, 0.001f), make_float2(1.f, 1.f));
    } else {
      pv.x = ex2(x.x);
      pv.y = ex2(x.y);
    }
    acc4[i & 3] = __fadd2_rn(acc4[i & 3], pv);
    pr[i] = pack2<kBF16>(pv.x, pv.y);
  }
  const float2 a01 = __fadd2_rn(acc4[0], acc4[1]), a23 = __fadd2_rn(acc4[2], acc4[3]);
  return __fadd2_rn(a01, a23);
}

template <int kD, bool kBF16, bool kOutF32, bool kMaxLogits, bool kSplit>
__global__ void __launch_bounds__(kSplit ? fwd::kThreadsSplit : fwd::kThreads, 1)
ffa_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const FwdArgs args) {
  using namespace fwd;
  using C = Cfg<kD, kSplit>;
  constexpr uint32_t kTileBytes = C::kTileBytes, kBoxes = C::kBoxes;
  constexpr uint32_t kSmemQ = C::kSmemQ, kSmemKV = C::kSmemKV, kSmemBar = C::kSmemBar;
  constexpr int kStages = C::kNumStages;              // shadows fwd::kStages: ring depth of this layout
  constexpr int kSmWarps = kSplit ? 16 : 8;           // softmax warps
  constexpr int kMmaWarp = kSmWarps, kTmaWarp = kSmWarps + 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmemBar + kNumBars * 8);
  const uint32_t bar0 = smem_base + kSmemBar;
  auto bar = [&](int i) -> uint32_t { return bar0 + uint32_t(i) * 8u; };

  const int tid = threadIdx.x;
  const int wid = tid >> 5;
  const int lane = tid & 31;

  // ---- work decode: kv-head major, then tile (heaviest = highest index first), then q head in group
  const int G = args.num_heads_q / args.num_heads_kv;
  const int per_kvh = args.n_tiles * G;
  const int hk = blockIdx.x / per_kvh;
  const int rem = blockIdx.x - hk * per_kvh;
  const int tile = args.n_tiles - 1 - rem / G;
  const int hq = hk * G + rem % G;
  const int R0 = tile * kFwdTileQ;

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
      mbar_init(bar(kBarP + t), 128);
      mbar_init(bar(kBarPhi + t), 128);
      mbar_init(bar(kBarO + t), 1);
    }
    fence_barrier_init();
  }
  if (wid == kMmaWarp) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  if (wid == kTmaWarp && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if constexpr (kSplit) {
   if (wid < kSmWarps) {
    // =========================== softmax: four warpgroups = (tile, column half) ===========================
    setmaxnreg_inc<104>();
    const int grp = wid >> 2, t = grp >> 1, hf = grp & 1;
    const int row_in_tile = (wid & 3) * 32 + lane;  // TMEM lane
    const int row = R0 + t * kBlockM + row_in_tile;
    const uint32_t lane_off = uint32_t((wid & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_off + kTmemS + t * 128 + hf * 64;   // my 64 score columns
    const uint32_t tP = tmem + lane_off + kTmemS + t * 128 + hf * 32;   // my 32 packed P columns
    const uint32_t tO = tmem + lane_off + kTmemO + t * 128 + hf * (kD / 2);
    const float sl2 = args.scale * 1.4426950408889634f;
    float* xch = reinterpret_cast<float*>(smem + C::kSmemXch);
    auto xslot = [&](int parity, int half) -> float* { return xch + ((parity * 2 + t) * 2 + half) * 128 + row_in_tile; };

    float m_ref = -INFINITY, l_sum = 0.f, m_true = -INFINITY;
    uint32_t n_act = 0;
    SliceCursor cur;
    cur.init(args.meta, args.plan, tile);
    Slice sl;
    CtaGeom g;
    TileGeom tg;
    while (cur.next(sl)) {
      if (!g.init(sl, R0)) continue;
      tg.init(sl, R0 + t * kBlockM);
      const bool row_in = (row >= sl.q0) && (row < sl.q1);
      const int lo_r = row_in ? sl.lo(row) : 0;
      const int hi_r = row_in ? min(sl.hi(row), sl.k1) : 0;  // empty range when row not in slice
      for (int b = 0; b < g.nblk; ++b) {
        const int bs = g.c_min + b * kBlockN;
        if (!tg.active(bs)) continue;
        mbar_wait(bar(kBarS + t), n_act & 1);
        tc_fence_after();
        uint32_t sr[64];
        tmem_ld32(tS, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
        tmem_ld32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
        tc_wait_ld();
        if (!tg.full(bs)) {
          const int lo_rel = lo_r - bs - hf * 64, hi_rel = hi_r - bs - hf * 64;
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
        *xslot(n_act & 1, hf) = mb;  // my half's block maximum of this row
        // first 32 exponentials on the stale reference while the other half's maximum travels
        float mneg = (m_ref == -INFINITY) ? 0.f : -m_ref * sl2;
        uint32_t pr[16];
        float2 a0 = softmax_quarter<kBF16, true>(&sr[0], sl2, mneg, pr);
        named_bar_sync(1 + t, 256);  // also: both halves hold their S columns in registers => P may overwrite the S region
        tmem_st16(tP, pr);
        mb = fmaxf(mb, *xslot(n_act & 1, hf ^ 1));
        if constexpr (kMaxLogits) m_true = fmaxf(m_true, mb);
        const float m_new = fmaxf(m_ref, mb);
        const bool upd = (m_new - m_ref) * sl2 > kRescaleThreshold;  // identical in both halves of the row
        if (__any_sync(0xffffffffu, upd)) {
          float corr = 1.f;
          if (upd) {
            corr = (m_ref == -INFINITY) ? 0.f : ex2((m_ref - m_new) * sl2);
            m_ref = m_new;
            l_sum *= corr;
          }
          if (n_act > 0) {  // rescale my half of the O columns (PV of the previous block has completed, see above)
#pragma unroll
            for (int c = 0; c < kD / 64; ++c) {
              uint32_t o[32];
              tmem_ld32(tO + c * 32, o);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
              tmem_st32(tO + c * 32, o);
            }
          }
          mneg = (m_ref == -INFINITY) ? 0.f : -m_ref * sl2;
          a0 = softmax_quarter<kBF16, true>(&sr[0], sl2, mneg, pr);
          tmem_st16(tP, pr);
        }
        const float2 a1 = softmax_quarter<kBF16, false>(&sr[32], sl2, mneg, pr);
        tmem_st16(tP + 16, pr);
        l_sum += (a0.x + a0.y) + (a1.x + a1.y);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(bar((hf == 0 ? kBarP : kBarPhi) + t));  // my half of P (= PV k-steps 4*hf .. 4*hf+3) is in TMEM
        ++n_act;
      }
    }
    // ------------------------------- epilogue (split) -------------------------------
    if constexpr (kMaxLogits) {
      float mt = m_true;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
      if (lane == 0 && mt != -INFINITY && hf == 0) {
        const float val = mt * args.scale;
        if (!signbit(val)) atomicMax(reinterpret_cast<int*>(args.max_logits + hq), __float_as_int(val));
        else atomicMin(reinterpret_cast<unsigned int*>(args.max_logits + hq), __float_as_uint(val));
      }
    }
    const bool row_ok = row < args.total_q;
    float* lse_ptr = args.lse + int64_t(row) * args.lse_stride_t + int64_t(hq) * args.lse_stride_h;
    const float lse_old = (args.accumulate && row_ok) ? *lse_ptr : -INFINITY;  // read before the other half may write it
    // the slot of parity n_act was last read two blocks ago: free for the row-sum exchange
    *xslot(n_act & 1, hf) = l_sum;
    named_bar_sync(1 + t, 256);
    l_sum += *xslot(n_act & 1, hf ^ 1);
    const bool dead = (n_act == 0) || (m_ref == -INFINITY) || !(l_sum > 0.f);
    float lse_new = dead ? -INFINITY : fmaf(m_ref, args.scale, __logf(l_sum));
    float w_new = dead ? 0.f : 1.f / l_sum;
    float w_old = 0.f;
    bool need_old = false;
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
    if (n_act > 0) {
      mbar_wait(bar(kBarO + t), 0);
      tc_fence_after();
    }
    using OutT = typename std::conditional<kOutF32, float, typename std::conditional<kBF16, __nv_bfloat16, __half>::type>::type;
    OutT* orow = reinterpret_cast<OutT*>(args.out) + int64_t(row) * args.o_stride_t + int64_t(hq) * args.o_stride_h + hf * (kD / 2);
    const bool skip_store = !row_ok || (args.accumulate && dead && need_old);  // nothing new: keep the old partial
#pragma unroll
    for (int c = 0; c < kD / 64; ++c) {
      float o[32];
      if (n_act > 0 && !dead) {
        uint32_t r[32];
        tmem_ld32(tO + c * 32, r);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(r[i]) * w_new;
      } else {
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
            dst[i] = make_uint4(pack2<kBF16>(x[0], x[1]), pack2<kBF16>(x[2], x[3]), pack2<kBF16>(x[4], x[5]),
                                pack2<kBF16>(x[6], x[7]));
          }
        }
      }
    }
    if (!skip_store && hf == 0) *lse_ptr = lse_new;
   }
  }
  if (!kSplit && wid < 8) {
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
    SliceCursor cur;
    cur.init(args.meta, args.plan, tile);
    Slice sl;
    CtaGeom g;
    TileGeom tg;
    while (cur.next(sl)) {
      if (!g.init(sl, R0)) continue;
      tg.init(sl, R0 + t * kBlockM);
      const bool row_in = (row >= sl.q0) && (row < sl.q1);
      const int lo_r = row_in ? sl.lo(row) : 0;
      const int hi_r = row_in ? min(sl.hi(row), sl.k1) : 0;  // empty range when row not in slice
      for (int b = 0; b < g.nblk; ++b) {
        const int bs = g.c_min + b * kBlockN;
        if (!tg.active(bs)) continue;
        if ((wid & 3) == 0) FFA_TR(1 + t, n_act, 0);
        mbar_wait(bar(kBarS + t), n_act & 1);
        tc_fence_after();
        if ((wid & 3) == 0) FFA_TR(1 + t, n_act, 1);
        uint32_t sr[128];
        if (FFA_FWD_ABL & 8) {
#pragma unroll
          for (int c = 0; c < 128; ++c) sr[c] = 0x3f800000u + c + n_act;
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[c * 32]));
          tc_wait_ld();
        }
        if ((wid & 3) == 0) FFA_TR(1 + t, n_act, 2);
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
          for (int i = 0; i < ((FFA_FWD_ABL & 2) ? 8 : 128); i += 8) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              m4[j] = fmaxf(m4[j], fmaxf(__uint_as_float(sr[i + 2 * j]), __uint_as_float(sr[i + 2 * j + 1])));
          }
          mb = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        }
        if constexpr (kMaxLogits) m_true = fmaxf(m_true, mb);
        float mneg = (m_ref == -INFINITY) ? 0.f : -m_ref * sl2;
        uint32_t pr[16];
        float2 a0 = softmax_quarter<kBF16, true>(&sr[0], sl2, mneg, pr);
        if (!(FFA_FWD_ABL & 4)) tmem_st16(tS, pr);
        float2 a1 = softmax_quarter<kBF16, true>(&sr[32], sl2, mneg, pr);
        if (!(FFA_FWD_ABL & 4)) tmem_st16(tS + 16, pr);
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
          a0 = softmax_quarter<kBF16, true>(&sr[0], sl2, mneg, pr);
          if (!(FFA_FWD_ABL & 4)) tmem_st16(tS, pr);
          a1 = softmax_quarter<kBF16, true>(&sr[32], sl2, mneg, pr);
          if (!(FFA_FWD_ABL & 4)) tmem_st16(tS + 16, pr);
        }
        // third quarter's exponentials run while the stores of the first half land: the wait below is then (nearly) free
        const float2 a2 = softmax_quarter<kBF16, true>(&sr[64], sl2, mneg, pr);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(bar(kBarP + t));  // P[:, 0:64) (and any O rescale) visible: first 4 PV k-steps may start
        if ((wid & 3) == 0) FFA_TR(1 + t, n_act, 3);
        if (!(FFA_FWD_ABL & 4)) tmem_st16(tS + 32, pr);
        const float2 a3 = softmax_quarter<kBF16, false>(&sr[96], sl2, mneg, pr);
        if (!(FFA_FWD_ABL & 4)) tmem_st16(tS + 48, pr);
        l_sum += ((a0.x + a0.y) + (a1.x + a1.y)) + ((a2.x + a2.y) + (a3.x + a3.y));
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(bar(kBarPhi + t));
        if ((wid & 3) == 0) FFA_TR(1 + t, n_act, 4);
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
        // ordered-int trick keyed on the SIGN BIT (not on val >= 0: -0.0 compares >= 0 but is INT_MIN as an int)
        if (!signbit(val)) atomicMax(reinterpret_cast<int*>(args.max_logits + hq), __float_as_int(val));
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
  }
  if (wid >= kSmWarps) {
    // register pool: 384 threads x 168 regs (launch bounds) = 64512 = 256 x 208 + 128 x 88;  split layout: 640 threads x 96
    // at launch = 61440 registers in the CTA pool (setmaxnreg only redistributes what the launch allocated), 512 x 104 +
    // 128 x 64 = 61440 after the moves
    if constexpr (kSplit) setmaxnreg_dec<64>(); else setmaxnreg_dec<88>();
    if (wid == kMmaWarp) {
      // =========================== MMA issuer ===========================
      // ONE warp issues the MMAs of both q tiles in the fixed order PV0 S0 PV1 S1: that order staggers the two tiles (the
      // softmax of one overlaps the MMAs of the other).  A variant with one issuing warp per tile was measured 18 % SLOWER
      // (round 2): the two chains fall into lock-step, both softmax phases then share the MUFU and the pipe idles.
      // The whole warp runs this loop with warp-uniform control flow (descriptors live in uniform registers);
      // only the tcgen05.mma / tcgen05.commit instructions themselves are issued by one elected lane.
      constexpr uint32_t idesc_s = make_idesc(kBF16, false, false, 128, 128);
      constexpr uint32_t idesc_pv = make_idesc(kBF16, false, true, 128, kD);  // O tile is 128 x kD
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
        const uint32_t sV = sKV + buf * kTileBytes;
        const uint32_t tP = tmem + kTmemS + t * 128, tOt = tmem + kTmemO + t * 128;
        FFA_TR(0, n_act_t - 1, 4 * t + 0);
        mbar_wait(bar(kBarFull + buf), (li / kStages) & 1);
        mbar_wait(bar(kBarP + t), (n_act_t - 1) & 1);  // first 64 columns of P_t are in TMEM
        tc_fence_after();
        FFA_TR(0, n_act_t - 1, 4 * t + 1);
        if (elect_one()) {
#pragma unroll
          for (uint32_t kk = 0; kk < ((FFA_FWD_ABL & 32) ? 0 : 4); ++kk)
            mma_ts(tOt, tP + kk * 8, smem_desc(desc_v_hi, sV + kk * 2048), idesc_pv, (n_act_t > 1 || kk > 0) ? 1u : 0u);
        }
        __syncwarp();
        mbar_wait(bar(kBarPhi + t), (n_act_t - 1) & 1);  // second half
        tc_fence_after();
        FFA_TR(0, n_act_t - 1, 4 * t + 2);
        pend -= 1u << (4 * buf);
        const bool release = ((pend >> (4 * buf)) & 15u) == 0;
        if (elect_one()) {
#pragma unroll
          for (uint32_t kk = 4; kk < ((FFA_FWD_ABL & 32) ? 0 : 8); ++kk)
            mma_ts(tOt, tP + kk * 8, smem_desc(desc_v_hi, sV + kk * 2048), idesc_pv, 1u);
          if (release) tc_commit(bar(kBarEmpty + buf));
        }
        __syncwarp();
      };
      auto issue_s = [&](const int t, const uint32_t sK) {
        const uint32_t tSt = tmem + kTmemS + t * 128;
        if (elect_one()) {
#pragma unroll
          for (uint32_t kk = 0; kk < ((FFA_FWD_ABL & 64) ? 0 : kD / 16); ++kk) {  // contraction over the head dim
            const uint32_t off = (kk >> 2) * kBoxBytes + (kk & 3) * 32;
            mma_ss(tSt, smem_desc(DESC_KMAJOR, sQ + t * kTileBytes + off), smem_desc(DESC_KMAJOR, sK + off), idesc_s,
                   kk > 0 ? 1u : 0u);
          }
          tc_commit(bar(kBarS + t));
        }
        __syncwarp();
        FFA_TR(0, (t == 0 ? n_act0 : n_act1), 4 * t + 3);  // S of block n_act (0-based) issued
      };

      SliceCursor cur;
      cur.init(args.meta, args.plan, tile);
      Slice sl;
      CtaGeom g;
      TileGeom tg0, tg1;
      while (cur.next(sl)) {
        if (!g.init(sl, R0)) continue;
        tg0.init(sl, R0);
        tg1.init(sl, R0 + kBlockM);
        for (int b = 0; b < g.nblk; ++b, ++blk) {
          const int bs = g.c_min + b * kBlockN;
          const bool a0 = tg0.active(bs), a1 = tg1.active(bs);
          if (!q_ready) {
            mbar_wait(bar(kBarQ), 0);
            q_ready = true;
          }
          const uint32_t lk = 2 * blk, kbuf = lk % kStages, lv = lk + 1, vbuf = lv % kStages;
          mbar_wait(bar(kBarFull + kbuf), (lk / kStages) & 1);
          tc_fence_after();
          pend = (pend & ~(15u << (4 * vbuf))) | (((a0 ? 1u : 0u) + (a1 ? 1u : 0u)) << (4 * vbuf));
          const uint32_t sK = sKV + kbuf * kTileBytes;
          if (has_prev0) { issue_pv(0, n_act0, prev_load0); has_prev0 = false; }
          if (a0) { issue_s(0, sK); has_prev0 = true; prev_load0 = lv; ++n_act0; }
          if (has_prev1) { issue_pv(1, n_act1, prev_load1); has_prev1 = false; }
          if (a1) { issue_s(1, sK); has_prev1 = true; prev_load1 = lv; ++n_act1; }
          if (elect_one()) {
            tc_commit(bar(kBarEmpty + kbuf));  // K block free once both S MMAs retire
            if (((pend >> (4 * vbuf)) & 15u) == 0) tc_commit(bar(kBarEmpty + vbuf));  // V block nobody needs
          }
          __syncwarp();
        }
      }
      if (has_prev0) issue_pv(0, n_act0, prev_load0);
      if (has_prev1) issue_pv(1, n_act1, prev_load1);
      if (elect_one()) {
        if (n_act0 > 0) tc_commit(bar(kBarO + 0));
        if (n_act1 > 0) tc_commit(bar(kBarO + 1));
      }
      __syncwarp();
    } else if (wid == kTmaWarp) {
      // =========================== TMA producer (warp-uniform loop, one elected lane issues) ===========
      const uint32_t sQ = smem_base + kSmemQ, sKV = smem_base + kSmemKV;
      uint32_t blk = 0;
      bool q_loaded = false;
      SliceCursor cur;
      cur.init(args.meta, args.plan, tile);
      Slice sl;
      CtaGeom g;
      while (cur.next(sl)) {
        if (!g.init(sl, R0)) continue;
        if (!q_loaded) {
          if (elect_one()) {
            mbar_arrive_expect_tx(bar(kBarQ), 2 * kTileBytes);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
#pragma unroll
              for (uint32_t bx = 0; bx < kBoxes; ++bx)
                tma_load_3d(sQ + t * kTileBytes + bx * kBoxBytes, &tmQ, bar(kBarQ), 64 * bx, hq, R0 + t * kBlockM);
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
            FFA_TR(3, uint32_t(blk), 2 * which);
            mbar_wait(bar(kBarEmpty + buf), ((li / kStages) & 1) ^ 1);
            FFA_TR(3, uint32_t(blk), 2 * which + 1);
            if (elect_one()) {
              if ((FFA_FWD_ABL & 16) && li >= uint32_t(kStages)) {
                mbar_arrive(bar(kBarFull + buf));
              } else {
                mbar_arrive_expect_tx(bar(kBarFull + buf), kTileBytes);
                const CUtensorMap* tm = which == 0 ? &tmK : &tmV;
#pragma unroll
                for (uint32_t bx = 0; bx < kBoxes; ++bx)
                  tma_load_3d(sKV + buf * kTileBytes + bx * kBoxBytes, tm, bar(kBarFull + buf), 64 * bx, hk, bs);
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
  if (wid == kMmaWarp) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------
// B200_FFA_FWD_SPLIT=1 selects the four-warpgroup layout (two softmax threads per row); default: see kDefaultSplit
#ifndef FFA_FWD_DEFAULT_SPLIT
#define FFA_FWD_DEFAULT_SPLIT 0
#endif
static bool use_split() {
  static const int v = [] {
    const char* e = getenv("B200_FFA_FWD_SPLIT");
    return e ? atoi(e) : FFA_FWD_DEFAULT_SPLIT;
  }();
  return v != 0;
}

template <int kD, bool kBF16, bool kOutF32, bool kMaxLogits, bool kSplit>
static cudaError_t launch_one_s(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                                const FwdArgs& a, int grid, cudaStream_t stream) {
  auto kern = ffa_fwd_kernel<kD, kBF16, kOutF32, kMaxLogits, kSplit>;
  constexpr uint32_t smem_bytes = fwd::Cfg<kD, kSplit>::kSmemBytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  kern<<<grid, kSplit ? fwd::kThreadsSplit : fwd::kThreads, smem_bytes, stream>>>(tmQ, tmK, tmV, a);
  return cudaGetLastError();
}

template <int kD, bool kBF16, bool kOutF32, bool kMaxLogits>
static cudaError_t launch_one(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                              const FwdArgs& a, int grid, cudaStream_t stream) {
  return use_split() ? launch_one_s<kD, kBF16, kOutF32, kMaxLogits, true>(tmQ, tmK, tmV, a, grid, stream)
                     : launch_one_s<kD, kBF16, kOutF32, kMaxLogits, false>(tmQ, tmK, tmV, a, grid, stream);
}

template <int kD, bool kMaxLogits>
static cudaError_t launch_dm(bool bf16, bool out_f32, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                             const FwdArgs& a, int grid, cudaStream_t stream) {
  if (bf16) return out_f32 ? launch_one<kD, true, true, kMaxLogits>(tmQ, tmK, tmV, a, grid, stream)
                           : launch_one<kD, true, false, kMaxLogits>(tmQ, tmK, tmV, a, grid, stream);
  return out_f32 ? launch_one<kD, false, true, kMaxLogits>(tmQ, tmK, tmV, a, grid, stream)
                 : launch_one<kD, false, false, kMaxLogits>(tmQ, tmK, tmV, a, grid, stream);
}

// the max_logits bookkeeping is a separate instantiation so that the common path keeps its register budget
template <int kD>
static cudaError_t launch_d(bool bf16, bool out_f32, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                            const FwdArgs& a, int grid, cudaStream_t stream) {
  return a.max_logits ? launch_dm<kD, true>(bf16, out_f32, tmQ, tmK, tmV, a, grid, stream)
                      : launch_dm<kD, false>(bf16, out_f32, tmQ, tmK, tmV, a, grid, stream);
}

cudaError_t launch_fwd(int head_dim, bool bf16, bool out_f32, const CUtensorMap& tmQ, const CUtensorMap& tmK,
                       const CUtensorMap& tmV, const FwdArgs& a, cudaStream_t stream) {
  const int grid = a.n_tiles * a.num_heads_q;
  if (grid == 0) return cudaSuccess;
  return head_dim == 64 ? launch_d<64>(bf16, out_f32, tmQ, tmK, tmV, a, grid, stream)
                        : launch_d<128>(bf16, out_f32, tmQ, tmK, tmV, a, grid, stream);
}

}  // namespace ffa
