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
};

}  // namespace ffa
