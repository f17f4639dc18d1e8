// ffa_api.cu — the C ABI declared in include/b200_ffa.h: argument checking, TMA tensor-map encoding,
// workspace carving and kernel launches.  Mirrors the role of the reference's host glue
// (csrc/flexible_flash_attention/flex_flash_common.cpp, flex_flash_fwd.hpp:190-383,
// flex_flash_bwd.hpp:200-425) minus everything torch: no allocation, no sync, no exceptions.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../include/b200_ffa.h"
#include "ffa_args.h"
#include "ffa_kernels.h"

namespace {

thread_local char g_err[512] = "";
thread_local int g_launches = 0;

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiled_t resolve_encode() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
  return (EncodeTiled_t)p;
}

EncodeTiled_t get_encode() {
  static EncodeTiled_t fn = resolve_encode();  // function-local static: thread-safe one-time initialisation
  return fn;
}

// [tokens, heads, head_dim] 16-bit tensor -> 3-D map (d, h, t), box (64, 1, box_rows), SWIZZLE_128B.
int make_map_thd(CUtensorMap* m, const void* ptr, int dtype, int head_dim, int tokens, int heads, int64_t stride_t, int64_t stride_h,
                 int box_rows, const char* what) {
  EncodeTiled_t enc = get_encode();
  if (!enc) return fail(B200_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (stride_t * 2) % 16 || (stride_h * 2) % 16)
    return fail(B200_ERR_INVALID, "%s: pointer/strides must be 16-byte aligned", what);
  cuuint64_t dims[3] = {cuuint64_t(head_dim), cuuint64_t(heads), cuuint64_t(tokens)};
  cuuint64_t strides[2] = {cuuint64_t(stride_h) * 2, cuuint64_t(stride_t) * 2};
  cuuint32_t box[3] = {64, 1, cuuint32_t(box_rows)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, dtype == B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%d)", what, int(r));
  return 0;
}

// fp32 [tokens, heads, head_dim] accumulator -> 3-D map (d, h, t), box (32, 1, 128), SWIZZLE_128B (TMA reduce-add target)
int make_map_f32_acc(CUtensorMap* m, void* ptr, int head_dim, int tokens, int heads, int64_t stride_t, int64_t stride_h, const char* what) {
  EncodeTiled_t enc = get_encode();
  if (!enc) return fail(B200_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {cuuint64_t(head_dim), cuuint64_t(heads), cuuint64_t(tokens)};
  cuuint64_t strides[2] = {cuuint64_t(stride_h) * 4, cuuint64_t(stride_t) * 4};
  cuuint32_t box[3] = {32, 1, 128};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%d)", what, int(r));
  return 0;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int check_common(int dtype, int hq, int hk, int d, float softcap, int sm_margin) {
  if (sm_margin != 0) return fail(B200_ERR_UNSUPPORTED, "sm_margin %d is not supported (must be 0)", sm_margin);
  if (dtype != B200_F16 && dtype != B200_BF16) return fail(B200_ERR_INVALID, "dtype must be f16 or bf16");
  if (d != 64 && d != 128) return fail(B200_ERR_UNSUPPORTED, "head_dim %d unsupported (only 64 and 128)", d);
  if (hq <= 0 || hk <= 0 || hq % hk) return fail(B200_ERR_INVALID, "num_heads_q %d must be a multiple of num_heads_kv %d", hq, hk);
  if (softcap != 0.f) return fail(B200_ERR_UNSUPPORTED, "softcap is not supported");
  return 0;
}

int check_sink(const float* sink, int s_sink, int layout) {
  if (!sink) return 0;
  if (s_sink < 0 || s_sink > 8) return fail(B200_ERR_INVALID, "seqlen_sink %d must be in [0, 8]", s_sink);
  if (layout != 0 && layout != 1) return fail(B200_ERR_UNSUPPORTED, "sink_layout %d unsupported (0 = sh, 1 = ssh)", layout);
  return 0;
}

}  // namespace

extern "C" {

const char* b200_ffa_version(void) { return "b200_ffa 0.2.0 sm_100a"; }
const char* b200_ffa_last_error(void) { return g_err; }
int b200_ffa_last_launch_count(void) { return g_launches; }

size_t b200_ffa_fwd_workspace_bytes(const b200_ffa_fwd_params* p) {
  if (!p) return 0;
  int n_tiles = (p->total_q + ffa::kFwdTileQ - 1) / ffa::kFwdTileQ;
  return align_up(ffa::plan_bytes(n_tiles, p->num_slices), 256);
}

int b200_ffa_fwd(const b200_ffa_fwd_params* p, b200_stream_t stream_) {
  g_err[0] = 0;
  g_launches = 0;
  if (!p) return fail(B200_ERR_INVALID, "null params");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = check_common(p->dtype, p->num_heads_q, p->num_heads_kv, p->head_dim, p->softcap, p->sm_margin)) return rc;
  if (p->out_dtype != p->dtype && p->out_dtype != B200_F32) return fail(B200_ERR_INVALID, "out_dtype must equal dtype or be f32");
  if (p->total_q < 0 || p->total_k < 0 || p->num_slices < 0) return fail(B200_ERR_INVALID, "negative size");
  if (p->total_q == 0) return 0;
  if (!p->q || !p->out || !p->lse) return fail(B200_ERR_INVALID, "null q/out/lse");
  if (!(p->softmax_scale > 0.f)) return fail(B200_ERR_INVALID, "softmax_scale must be positive");
  if (int rc = check_sink(p->sink, p->seqlen_sink, p->sink_layout)) return rc;
  if (p->num_slices > 0 && (!p->q_ranges || !p->k_ranges)) return fail(B200_ERR_INVALID, "null ranges");
  if (p->total_k > 0 && (!p->k || !p->v)) return fail(B200_ERR_INVALID, "null k/v");
  const size_t o_align = p->out_dtype == B200_F32 ? 16 : 16;
  if ((reinterpret_cast<uintptr_t>(p->out) % o_align) || (p->o_stride_t * (p->out_dtype == B200_F32 ? 4 : 2)) % 16 ||
      (p->o_stride_h * (p->out_dtype == B200_F32 ? 4 : 2)) % 16)
    return fail(B200_ERR_INVALID, "out pointer/strides must be 16-byte aligned");
  const int n_tiles = (p->total_q + ffa::kFwdTileQ - 1) / ffa::kFwdTileQ;
  const size_t need = b200_ffa_fwd_workspace_bytes(p);
  if (!p->workspace || p->workspace_bytes < ffa::plan_entries_off(n_tiles) * sizeof(int) + sizeof(int) ||
      (reinterpret_cast<uintptr_t>(p->workspace) & 255))
    return fail(B200_ERR_WORKSPACE, "workspace too small (%zu < %zu) or not 256-byte aligned", p->workspace_bytes, need);

  ffa::SliceMeta meta{reinterpret_cast<const int2*>(p->q_ranges), reinterpret_cast<const int2*>(p->k_ranges),
                      p->attn_type_map, p->num_slices};
  int* plan = static_cast<int*>(p->workspace);
  const size_t cap_bytes = p->workspace_bytes - ffa::plan_entries_off(n_tiles) * sizeof(int);
  const int capacity = int(cap_bytes / sizeof(int) > 0x7fffffff ? 0x7fffffff : cap_bytes / sizeof(int));
  cudaError_t e = ffa::launch_plan(true, meta, n_tiles, capacity, plan, stream);
  if (e != cudaSuccess) return fail(B200_ERR_CUDA, "plan launch: %s", cudaGetErrorString(e));
  ++g_launches;

  CUtensorMap tmQ, tmK, tmV;
  if (int rc = make_map_thd(&tmQ, p->q, p->dtype, p->head_dim, p->total_q, p->num_heads_q, p->q_stride_t, p->q_stride_h, 128, "q")) return rc;
  // k/v may legally be empty (total_k == 0): point the maps at q so encoding succeeds; no block is ever loaded.
  const bool kv_empty = p->total_k == 0;
  const int k_box_rows = 128, v_box_rows = 128;
  if (int rc = make_map_thd(&tmK, kv_empty ? p->q : p->k, p->dtype, p->head_dim, kv_empty ? 1 : p->total_k, kv_empty ? p->num_heads_q : p->num_heads_kv,
                            kv_empty ? p->q_stride_t : p->k_stride_t, kv_empty ? p->q_stride_h : p->k_stride_h, k_box_rows, "k")) return rc;
  if (int rc = make_map_thd(&tmV, kv_empty ? p->q : p->v, p->dtype, p->head_dim, kv_empty ? 1 : p->total_k, kv_empty ? p->num_heads_q : p->num_heads_kv,
                            kv_empty ? p->q_stride_t : p->v_stride_t, kv_empty ? p->q_stride_h : p->v_stride_h, v_box_rows, "v")) return rc;

  ffa::FwdArgs a{};
  a.meta = meta;
  a.plan = plan;
  a.out = p->out;
  a.lse = p->lse;
  a.total_q = p->total_q;
  a.total_k = p->total_k;
  a.num_heads_q = p->num_heads_q;
  a.num_heads_kv = p->num_heads_kv;
  a.n_tiles = n_tiles;
  a.o_stride_t = p->o_stride_t;
  a.o_stride_h = p->o_stride_h;
  a.lse_stride_t = p->lse_stride_t;
  a.lse_stride_h = p->lse_stride_h;
  a.scale = p->softmax_scale;
  a.accumulate = p->accumulate;
  a.max_logits = p->max_logits;
  e = ffa::launch_fwd(p->head_dim, p->dtype == B200_BF16, p->out_dtype == B200_F32, tmQ, tmK, tmV, a, stream);
  if (e != cudaSuccess) return fail(B200_ERR_CUDA, "fwd launch: %s", cudaGetErrorString(e));
  ++g_launches;
  if (p->sink && p->seqlen_sink > 0) {
    e = ffa::launch_sink_fwd(p->out_dtype, p->out, p->lse, p->sink, p->total_q, p->num_heads_q, p->head_dim, p->seqlen_sink,
                             p->sink_layout, p->o_stride_t, p->o_stride_h, p->lse_stride_t, p->lse_stride_h, stream);
    if (e != cudaSuccess) return fail(B200_ERR_CUDA, "sink fwd launch: %s", cudaGetErrorString(e));
    ++g_launches;
  }
  return 0;
}


static inline int bwd_tq_ld(int total_q) { return (total_q + 127) / 128 * 128 + 128; }

size_t b200_ffa_bwd_workspace_bytes(const b200_ffa_bwd_params* p) {
  if (!p) return 0;
  const int n_tiles = (p->total_k + ffa::kBwdTileK - 1) / ffa::kBwdTileK;
  const size_t plan = align_up(ffa::plan_bytes(n_tiles, p->num_slices), 256);
  const size_t vec = align_up(size_t(p->num_heads_q) * bwd_tq_ld(p->total_q) * sizeof(float), 256);
  return plan + 2 * vec;
}

size_t b200_ffa_bwd_row_stats_bytes(int32_t total_q, int32_t num_heads_q) {
  if (total_q < 0 || num_heads_q <= 0) return 0;
  return 2 * align_up(size_t(num_heads_q) * bwd_tq_ld(total_q) * sizeof(float), 256);
}

int b200_ffa_bwd(const b200_ffa_bwd_params* p, b200_stream_t stream_) {
  g_err[0] = 0;
  g_launches = 0;
  if (!p) return fail(B200_ERR_INVALID, "null params");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = check_common(p->dtype, p->num_heads_q, p->num_heads_kv, p->head_dim, p->softcap, p->sm_margin)) return rc;
  if (p->dkv_dtype != p->dtype && p->dkv_dtype != B200_F32) return fail(B200_ERR_INVALID, "dkv_dtype must equal dtype or be f32");
  if (p->accumulate_dkv && p->dkv_dtype != B200_F32) return fail(B200_ERR_INVALID, "accumulate_dkv requires fp32 dk/dv");
  if (p->total_q < 0 || p->total_k < 0 || p->num_slices < 0) return fail(B200_ERR_INVALID, "negative size");
  if (p->total_q == 0 && p->total_k == 0) return 0;
  if (p->total_q > 0 && (!p->q || !p->dout || !p->out || !p->lse || !p->dq)) return fail(B200_ERR_INVALID, "null q/dout/out/lse/dq");
  if (p->total_k > 0 && (!p->k || !p->v || !p->dk || !p->dv)) return fail(B200_ERR_INVALID, "null k/v/dk/dv");
  if (p->num_slices > 0 && (!p->q_ranges || !p->k_ranges)) return fail(B200_ERR_INVALID, "null ranges");
  if (int rc = check_sink(p->sink, p->seqlen_sink, p->sink_layout)) return rc;
  if (p->sink && p->seqlen_sink > 0 && !p->dsink) return fail(B200_ERR_INVALID, "sink given without dsink");
  if (!(p->softmax_scale > 0.f)) return fail(B200_ERR_INVALID, "softmax_scale must be positive");
  // no keys: dq gets no contribution, nothing to write (rows that only see the sink have out = 0 => dsink += 0)
  if (p->total_k == 0) return 0;
  const int esz = p->dkv_dtype == B200_F32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(p->dk) & 15) || (reinterpret_cast<uintptr_t>(p->dv) & 15) || (p->dk_stride_t * esz) % 16 ||
      (p->dk_stride_h * esz) % 16 || (p->dv_stride_t * esz) % 16 || (p->dv_stride_h * esz) % 16)
    return fail(B200_ERR_INVALID, "dk/dv pointers/strides must be 16-byte aligned");
  if (p->total_q > 0 && ((reinterpret_cast<uintptr_t>(p->dq) & 15) || p->dq_stride_t % 4 || p->dq_stride_h % 4))
    return fail(B200_ERR_INVALID, "dq pointer/strides must be 16-byte aligned");
  if (p->total_q > 0 && ((reinterpret_cast<uintptr_t>(p->out) & 7) || (reinterpret_cast<uintptr_t>(p->dout) & 7) || p->o_stride_t % 4 ||
                         p->o_stride_h % 4 || p->do_stride_t % 4 || p->do_stride_h % 4))
    return fail(B200_ERR_INVALID, "out/dout pointers/strides must be 8-byte aligned");

  const int n_tiles = (p->total_k + ffa::kBwdTileK - 1) / ffa::kBwdTileK;
  const int tq_ld = bwd_tq_ld(p->total_q);
  const size_t vec = align_up(size_t(p->num_heads_q) * tq_ld * sizeof(float), 256);
  const size_t need = b200_ffa_bwd_workspace_bytes(p);
  const size_t min_plan = align_up((ffa::plan_entries_off(n_tiles) + 1) * sizeof(int), 256);
  if (!p->workspace || (reinterpret_cast<uintptr_t>(p->workspace) & 255) || p->workspace_bytes < 2 * vec + min_plan)
    return fail(B200_ERR_WORKSPACE, "workspace too small (%zu < %zu) or not 256-byte aligned", p->workspace_bytes, need);
  if (p->row_stats && (reinterpret_cast<uintptr_t>(p->row_stats) & 255)) return fail(B200_ERR_INVALID, "row_stats must be 256-byte aligned");
  char* ws = static_cast<char*>(p->workspace);
  char* stats = p->row_stats ? static_cast<char*>(p->row_stats) : ws;
  float* delta = reinterpret_cast<float*>(stats);
  float* lse2 = reinterpret_cast<float*>(stats + vec);
  const bool stats_ready = p->row_stats && p->row_stats_ready;
  int* plan = reinterpret_cast<int*>(ws + 2 * vec);
  const size_t plan_bytes_avail = p->workspace_bytes - 2 * vec;
  const size_t cap = (plan_bytes_avail - ffa::plan_entries_off(n_tiles) * sizeof(int)) / sizeof(int);
  const int capacity = int(cap > 0x7fffffff ? 0x7fffffff : cap);

  ffa::SliceMeta meta{reinterpret_cast<const int2*>(p->q_ranges), reinterpret_cast<const int2*>(p->k_ranges),
                      p->attn_type_map, p->num_slices};
  cudaError_t e = ffa::launch_plan(false, meta, n_tiles, capacity, plan, stream);
  if (e != cudaSuccess) return fail(B200_ERR_CUDA, "plan launch: %s", cudaGetErrorString(e));
  ++g_launches;
  if (!stats_ready) {
    e = ffa::launch_bwd_preprocess(p->head_dim, p->dtype == B200_BF16, p->out, p->dout, p->lse, delta, lse2, p->total_q, p->num_heads_q,
                                   tq_ld, p->o_stride_t, p->o_stride_h, p->do_stride_t, p->do_stride_h, p->lse_stride_t,
                                   p->lse_stride_h, stream);
    if (e != cudaSuccess) return fail(B200_ERR_CUDA, "bwd preprocess launch: %s", cudaGetErrorString(e));
    ++g_launches;
  }
  if (p->sink && p->seqlen_sink > 0 && p->total_q > 0) {
    e = ffa::launch_sink_bwd(p->dsink, p->sink, p->lse, delta, p->total_q, p->num_heads_q, tq_ld, p->seqlen_sink, p->sink_layout,
                             p->lse_stride_t, p->lse_stride_h, stream);
    if (e != cudaSuccess) return fail(B200_ERR_CUDA, "sink bwd launch: %s", cudaGetErrorString(e));
    ++g_launches;
  }

  CUtensorMap tmQ, tmDO, tmK, tmV, tmDQ;
  const bool q_empty = p->total_q == 0;
  if (int rc = make_map_thd(&tmQ, q_empty ? p->k : p->q, p->dtype, p->head_dim, q_empty ? 1 : p->total_q, q_empty ? p->num_heads_kv : p->num_heads_q,
                            q_empty ? p->k_stride_t : p->q_stride_t, q_empty ? p->k_stride_h : p->q_stride_h, 128, "q")) return rc;
  if (int rc = make_map_thd(&tmDO, q_empty ? p->k : p->dout, p->dtype, p->head_dim, q_empty ? 1 : p->total_q, q_empty ? p->num_heads_kv : p->num_heads_q,
                            q_empty ? p->k_stride_t : p->do_stride_t, q_empty ? p->k_stride_h : p->do_stride_h, 128, "dout")) return rc;
  if (int rc = make_map_thd(&tmK, p->k, p->dtype, p->head_dim, p->total_k, p->num_heads_kv, p->k_stride_t, p->k_stride_h, 128, "k")) return rc;
  if (int rc = make_map_thd(&tmV, p->v, p->dtype, p->head_dim, p->total_k, p->num_heads_kv, p->v_stride_t, p->v_stride_h, 128, "v")) return rc;
  if (q_empty) {
    tmDQ = tmK;  // never used: no q block is ever visited
  } else if (int rc = make_map_f32_acc(&tmDQ, p->dq, p->head_dim, p->total_q, p->num_heads_q, p->dq_stride_t, p->dq_stride_h, "dq")) {
    return rc;
  }

  ffa::BwdArgs a{};
  a.meta = meta;
  a.plan = plan;
  a.lse2 = lse2;
  a.delta = delta;
  a.dq = p->dq;
  a.dk = p->dk;
  a.dv = p->dv;
  a.total_q = p->total_q;
  a.total_k = p->total_k;
  a.num_heads_q = p->num_heads_q;
  a.num_heads_kv = p->num_heads_kv;
  a.n_tiles = n_tiles;
  a.tq_ld = tq_ld;
  a.dq_stride_t = p->dq_stride_t;
  a.dq_stride_h = p->dq_stride_h;
  a.dk_stride_t = p->dk_stride_t;
  a.dk_stride_h = p->dk_stride_h;
  a.dv_stride_t = p->dv_stride_t;
  a.dv_stride_h = p->dv_stride_h;
  a.scale = p->softmax_scale;
  a.accumulate_dkv = p->accumulate_dkv;
  e = ffa::launch_bwd(p->head_dim, p->dtype == B200_BF16, p->dkv_dtype == B200_F32, tmQ, tmDO, tmK, tmV, tmDQ, a, stream);
  if (e != cudaSuccess) return fail(B200_ERR_CUDA, "bwd launch: %s", cudaGetErrorString(e));
  ++g_launches;
  return 0;
}

}  // extern "C"
