// ffa_common.cuh — slice geometry shared by the plan builder and the attention kernels.
//
// A "slice" is one (q_range x k_range, mask type) rectangle of the reference's FFA metadata
// (magi_attention/functional/flex_flash_attn.py:1066, mask semantics utils/_utils.py:834-877).
// In GLOBAL token coordinates, row r in [q0,q1) attends column c iff lo(r) <= c < hi(r) with
//     lo(r) = k0 + (type & 2 ? r - q0 : 0)           (inv-causal / bi-causal: top-left aligned)
//     hi(r) = type & 1 ? k1 - (q1 - 1 - r) : k1      (causal / bi-causal: bottom-right aligned)
// Both bounds are non-decreasing in r, which is what makes per-tile column hulls cheap.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ffa {

constexpr int kBlockM = 128;       // rows per MMA tile
constexpr int kBlockN = 128;       // kv columns per block
constexpr int kFwdTileQ = 256;     // q rows per forward CTA (two MMA tiles)
constexpr int kBwdTileK = 128;     // kv rows per backward CTA
constexpr int kHeadDim = 128;

struct Slice {
  int q0, q1, k0, k1, type;
  __host__ __device__ __forceinline__ int lo(int r) const { return k0 + ((type & 2) ? (r - q0) : 0); }
  __host__ __device__ __forceinline__ int hi(int r) const { return (type & 1) ? (k1 - (q1 - 1 - r)) : k1; }
  // rows that can see column c:  rlo(c) <= r < rhi(c)
  __host__ __device__ __forceinline__ int rlo(int c) const {
    int r = (type & 1) ? (q1 - (k1 - c)) : q0;
    return r > q0 ? r : q0;
  }
  __host__ __device__ __forceinline__ int rhi(int c) const {
    int r = (type & 2) ? (q0 + (c - k0) + 1) : q1;
    return r < q1 ? r : q1;
  }
};

// Column hull of rows [r_lo, r_hi) of a slice (r range already clipped to the slice): [c_min, c_max)
__host__ __device__ __forceinline__ bool col_hull(const Slice& s, int r_lo, int r_hi, int& c_min, int& c_max) {
  if (r_lo >= r_hi || s.k1 <= s.k0) return false;
  c_min = s.lo(r_lo);
  c_max = s.hi(r_hi - 1);
  if (c_max > s.k1) c_max = s.k1;
  return c_max > c_min;
}
// Row hull of columns [c_lo, c_hi) of a slice: [r_min, r_max)
__host__ __device__ __forceinline__ bool row_hull(const Slice& s, int c_lo, int c_hi, int& r_min, int& r_max) {
  if (c_lo >= c_hi || s.q1 <= s.q0) return false;
  r_min = s.rlo(c_lo);
  r_max = s.rhi(c_hi - 1);
  return r_max > r_min;
}

// Does the slice contribute anything to the forward tile of `rows` q rows starting at R0 (256 per CTA, 512 per CTA pair)?
__host__ __device__ __forceinline__ bool fwd_tile_hit(const Slice& s, int R0, int rows = kFwdTileQ) {
  int r_lo = s.q0 > R0 ? s.q0 : R0;
  int r_hi = s.q1 < R0 + rows ? s.q1 : R0 + rows;
  int a, b;
  return col_hull(s, r_lo, r_hi, a, b);
}
// Does the slice contribute anything to backward CTA tile (kv rows [C0, C0+128))?
__host__ __device__ __forceinline__ bool bwd_tile_hit(const Slice& s, int C0) {
  int c_lo = s.k0 > C0 ? s.k0 : C0;
  int c_hi = s.k1 < C0 + kBwdTileK ? s.k1 : C0 + kBwdTileK;
  int a, b;
  return row_hull(s, c_lo, c_hi, a, b);
}

// ---- plan workspace layout (ints) -------------------------------------------------------
//  [0] total entries  [1] overflow flag (1 => CSR unusable, kernels scan all slices)
//  [2] n_tiles        [3] capacity
//  [16 .. 16+n_tiles]  CSR offsets      then entries (slice ids) at `entries_off` ints.
constexpr int kPlanHeaderInts = 16;
__host__ __device__ __forceinline__ size_t plan_entries_off(int n_tiles) {
  size_t o = kPlanHeaderInts + size_t(n_tiles) + 1;
  return (o + 63) & ~size_t(63);
}
// Worst case one entry per (tile, slice); capped at kPlanMaxEntries (128 MiB of ids): a mask that overflows the cap sets
// the overflow flag and the kernels fall back to scanning every slice per tile (correct, slower) instead of asking the
// caller for an unbounded workspace.
constexpr size_t kPlanMaxEntries = size_t(32) << 20;
__host__ __device__ __forceinline__ size_t plan_bytes(int n_tiles, int num_slices) {
  size_t cap = size_t(n_tiles) * size_t(num_slices > 0 ? num_slices : 1);
  if (cap > kPlanMaxEntries) cap = kPlanMaxEntries;
  return (plan_entries_off(n_tiles) + cap) * sizeof(int);
}

struct SliceMeta {
  const int2* q_ranges;
  const int2* k_ranges;
  const int* types;  // may be null
  int num_slices;
};

// Iterates the slices relevant to one tile: through the plan's CSR list, or (overflow fallback)
// over all slices.  One slice of metadata is prefetched ahead of use.
struct SliceCursor {
  const int* ids;
  int pos, end;
  SliceMeta m;
  int2 nq, nk;
  int nt;
  __device__ __forceinline__ void init(const SliceMeta& meta, const int* plan, int tile) {
    m = meta;
    if (plan != nullptr && plan[1] == 0) {
      const int* offs = plan + kPlanHeaderInts;
      pos = offs[tile];
      end = offs[tile + 1];
      ids = plan + plan_entries_off(plan[2]);
    } else {
      pos = 0;
      end = meta.num_slices;
      ids = nullptr;
    }
    fetch();
  }
  __device__ __forceinline__ void fetch() {
    if (pos < end) {
      int sid = ids ? __ldg(ids + pos) : pos;
      nq = __ldg(m.q_ranges + sid);
      nk = __ldg(m.k_ranges + sid);
      nt = m.types ? __ldg(m.types + sid) : 0;
    }
  }
  __device__ __forceinline__ bool next(Slice& s) {
    if (pos >= end) return false;
    s.q0 = nq.x; s.q1 = nq.y; s.k0 = nk.x; s.k1 = nk.y; s.type = nt;
    ++pos;
    fetch();
    return true;
  }
};

}  // namespace ffa
