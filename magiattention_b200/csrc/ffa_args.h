// ffa_args.h — kernel argument blocks (plain structs passed by value to the kernels).
#pragma once
#include <cstdint>
#include "ffa_common.cuh"

namespace ffa {

struct FwdArgs {
  SliceMeta meta;
  const int* plan;
  void* out;
  float* lse;
  int total_q, total_k;
  int num_heads_q, num_heads_kv;
  int n_tiles;
  int64_t o_stride_t, o_stride_h, lse_stride_t, lse_stride_h;
  float scale;       // softmax scale
  int accumulate;
  float* max_logits; // [num_heads_q] running per-head max of the scaled logits, or nullptr
};

struct BwdArgs {
  SliceMeta meta;
  const int* plan;       // kv-tile -> slice CSR
  const float* lse2;     // [num_heads_q, tq_ld]  lse * log2(e), +inf for dead / pad rows
  const float* delta;    // [num_heads_q, tq_ld]  rowsum(O o dO)
  float* dq;             // fp32, accumulated with red.global.add
  void* dk;
  void* dv;
  int total_q, total_k;
  int num_heads_q, num_heads_kv;
  int n_tiles;           // kv tiles
  int tq_ld;
  int64_t dq_stride_t, dq_stride_h, dk_stride_t, dk_stride_h, dv_stride_t, dv_stride_h;
  float scale;
  int accumulate_dkv;
};

}  // namespace ffa
