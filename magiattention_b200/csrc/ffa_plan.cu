// ffa_plan.cu — builds the tile -> slice CSR list the attention kernels iterate.
//
// The reference's FFA schedules (slice, m_block, head) work items and merges overlapping q rows
// through global locks (csrc/flexible_flash_attention/fwd_tile_scheduler.hpp:143-345,
// epilogue_fwd.hpp:271-331); its FA4 path converts slices to per-token interval tables and CSR block
// lists on the host (flash-attention/csrc/utils/magi_to_hstu, create_block_mask).  Here one CTA owns a
// q tile (fwd) or a kv tile (bwd) and walks every slice that touches it, so the only metadata needed
// is this CSR of slice ids, built on-device with no host sync: count -> block scan -> fill.
#include "ffa_common.cuh"

namespace ffa {

template <bool kFwd>
__global__ void __launch_bounds__(1024, 1)
plan_kernel(SliceMeta meta, int n_tiles, int tile_rows, int capacity, int* __restrict__ plan) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  int* offs = plan + kPlanHeaderInts;
  int* ids = plan + plan_entries_off(n_tiles);
  const int tid = threadIdx.x;
  if (tid == 0) s_carry = 0;
  __syncthreads();

  // pass 1: per-tile counts + exclusive scan, 1024 tiles per round
  for (int base = 0; base < n_tiles; base += 1024) {
    int tile = base + tid;
    int cnt = 0;
    if (tile < n_tiles) {
      int T0 = tile * tile_rows;
      for (int s = 0; s < meta.num_slices; ++s) {
        int2 q = __ldg(meta.q_ranges + s), k = __ldg(meta.k_ranges + s);
        Slice sl{q.x, q.y, k.x, k.y, meta.types ? __ldg(meta.types + s) : 0};
        cnt += (kFwd ? fwd_tile_hit(sl, T0, tile_rows) : bwd_tile_hit(sl, T0)) ? 1 : 0;
      }
    }
    // block exclusive scan of cnt
    int v = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, v, o);
      if ((tid & 31) >= o) v += n;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
      int w = s_warp[tid];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, w, o);
        if (tid >= o) w += n;
      }
      s_warp[tid] = w;
    }
    __syncthreads();
    int excl = v - cnt + ((tid >> 5) ? s_warp[(tid >> 5) - 1] : 0) + s_carry;
    if (tile < n_tiles) offs[tile] = excl;
    __syncthreads();
    if (tid == 1023) s_carry = excl + cnt;
    __syncthreads();
  }
  const int total = s_carry;
  const bool overflow = total > capacity;
  if (tid == 0) {
    offs[n_tiles] = total;
    plan[0] = total;
    plan[1] = overflow ? 1 : 0;
    plan[2] = n_tiles;
    plan[3] = capacity;
  }
  if (overflow) return;
  // pass 2: fill
  for (int tile = tid; tile < n_tiles; tile += 1024) {
    int T0 = tile * tile_rows;
    int o = offs[tile];
    for (int s = 0; s < meta.num_slices; ++s) {
      int2 q = __ldg(meta.q_ranges + s), k = __ldg(meta.k_ranges + s);
      Slice sl{q.x, q.y, k.x, k.y, meta.types ? __ldg(meta.types + s) : 0};
      if (kFwd ? fwd_tile_hit(sl, T0, tile_rows) : bwd_tile_hit(sl, T0)) ids[o++] = s;
    }
  }
}

// forward plan over tiles of `tile_rows` q rows (512 for the CTA-pair forward)
cudaError_t launch_plan_fwd_rows(const SliceMeta& meta, int n_tiles, int tile_rows, int capacity, int* plan, cudaStream_t stream) {
  plan_kernel<true><<<1, 1024, 0, stream>>>(meta, n_tiles, tile_rows, capacity, plan);
  return cudaGetLastError();
}

cudaError_t launch_plan(bool fwd, const SliceMeta& meta, int n_tiles, int capacity, int* plan, cudaStream_t stream) {
  if (fwd)
    plan_kernel<true><<<1, 1024, 0, stream>>>(meta, n_tiles, kFwdTileQ, capacity, plan);
  else
    plan_kernel<false><<<1, 1024, 0, stream>>>(meta, n_tiles, kBwdTileK, capacity, plan);
  return cudaGetLastError();
}

}  // namespace ffa
