/* b200_ffa.h — C ABI of the B200-native flex-flash-attention library (libb200_ffa.so).
 *
 * Drop-in boundary for the reference's native FFA module.  Each entry point names the
 * reference interface it replaces (paths relative to the MagiAttention repo root):
 *
 *   b200_ffa_fwd        <- pybind `mod.fwd(q,k,v,sink?,out?,lse?,...,q_ranges,k_ranges,attn_type_map,
 *                          ...,softmax_scale,softcap,out_type,sink_layout,sm_margin)`
 *                          magi_attention/csrc/flexible_flash_attention/fwd_inst_template.jinja:67-138,
 *                          host glue flex_flash_fwd.hpp:190-383, params flash.h:32-143 (Flash_fwd_params);
 *                          on B200 it takes the slot of fa4_fwd (magi_attention/functional/fa4.py:128-221).
 *   b200_ffa_bwd        <- pybind `mod.bwd(dout,q,k,v,sink?,out,dq?,dk?,dv?,dsink?,lse,q_ranges,k_ranges,
 *                          attn_type_map,...,softmax_scale,softcap,dq_type,dk_type,dv_type,...)`
 *                          bwd_inst_template.jinja:58-107, flex_flash_bwd.hpp:200-425, flash.h:145-189;
 *                          B200 slot: fa4_bwd (functional/fa4.py:225-321).
 *   b200_ffa_*_workspace_bytes <- the per-call workspace allocations the reference makes itself
 *                          (range_locks / tile semaphore flex_flash_fwd.hpp:299-318; softmax_d / lse_log2
 *                          flex_flash_bwd.hpp:351-352).  Here the CALLER owns the workspace.
 *   b200_lse_merge      <- correct_attn_out_lse (magi_attention/functional/utils.py:467-558, Triton).
 *   b200_sink_merge     <- correct_attn_out_lse_with_sink (magi_attention/functional/utils.py:634-672).
 *   b200_range_reduce_sum <- range_reduce(sum) post-process of group_reduce
 *                          (magi_attention/common/range_op/_range_reduce.py:360).
 *
 * Conventions (differences from the pybind module are deliberate and listed in INTEGRATION.md):
 *   - plain pointers and sizes only; no torch types; all tensors are DEVICE pointers;
 *   - q/k/v/out/dout/dq/dk/dv are [tokens, heads, head_dim] with head_dim contiguous; strides in ELEMENTS;
 *   - q_ranges / k_ranges are int32 [num_slices, 2] ([start, end) token spans), attn_type_map int32
 *     [num_slices] or NULL (= all FULL); 0 full, 1 causal (bottom-right aligned), 2 inv-causal
 *     (top-left aligned), 3 bi-causal  (magi_attention/common/enum.py:58-76);
 *   - never allocates, never synchronises, never throws: work is enqueued on `stream`; returns 0 or a
 *     negative code and b200_ffa_last_error() describes the failure (thread-local string);
 *   - slices are area-disjoint (reference API contract); q ranges may overlap.
 */
#ifndef B200_FFA_H_
#define B200_FFA_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define B200_API __attribute__((visibility("default")))
#else
#define B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b200_stream_t; /* cudaStream_t */

enum b200_dtype { B200_F16 = 0, B200_BF16 = 1, B200_F32 = 2 };

enum b200_status {
  B200_OK = 0,
  B200_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
  B200_ERR_WORKSPACE = -2,   /* workspace too small or misaligned */
  B200_ERR_CUDA = -3,        /* a CUDA runtime/driver call failed */
  B200_ERR_UNSUPPORTED = -4  /* valid in the reference, not implemented here (e.g. softcap) */
};

typedef struct {
  /* inputs */
  const void* q;               /* [total_q, num_heads_q, head_dim]  f16/bf16 */
  const void* k;               /* [total_k, num_heads_kv, head_dim] */
  const void* v;               /* [total_k, num_heads_kv, head_dim] */
  /* outputs (caller-allocated) */
  void* out;                   /* [total_q, num_heads_q, head_dim], out_dtype */
  float* lse;                  /* [total_q, num_heads_q] fp32, natural log */
  /* slice metadata (device) */
  const int32_t* q_ranges;     /* [num_slices, 2] */
  const int32_t* k_ranges;     /* [num_slices, 2] */
  const int32_t* attn_type_map;/* [num_slices] or NULL */
  int32_t num_slices;
  /* shapes */
  int32_t total_q, total_k, num_heads_q, num_heads_kv, head_dim;
  int64_t q_stride_t, q_stride_h;     /* element strides */
  int64_t k_stride_t, k_stride_h;
  int64_t v_stride_t, v_stride_h;
  int64_t o_stride_t, o_stride_h;
  int64_t lse_stride_t, lse_stride_h;
  float softmax_scale;
  float softcap;               /* must be 0 */
  int32_t dtype;               /* b200_dtype of q/k/v */
  int32_t out_dtype;           /* b200_dtype of out: same as dtype, or B200_F32 */
  /* accumulate != 0: (out, lse) hold a previous partial result over a disjoint key set and are merged
   * in place with this call's result (reference: passing out_/lse to _flex_flash_attn_forward,
   * functional/flex_flash_attn.py:334-456; the CP stage merge dist_attn.py:3126-3131).
   * accumulate == 0: out/lse are overwritten; uncovered rows get out = 0, lse = -inf. */
  int32_t accumulate;
  int32_t sm_margin;           /* must be 0: the kernels are not persistent, every SM takes one CTA; a non-zero value is
                                * rejected with B200_ERR_UNSUPPORTED instead of being ignored (reference: sm_margin,
                                * flex_flash_attn.py:1103; env/comm.py:44-69) */
  void* workspace;             /* >= b200_ffa_fwd_workspace_bytes(), 256-byte aligned */
  size_t workspace_bytes;
  /* attention sink (NULL = none): `seqlen_sink` (<= 8, like the reference's kMaxSeqlenSink,
   * flash_fwd_postprocess_kernel.h:88) extra fp32 logit columns with zero value vectors, merged into (out, lse)
   * AFTER the slices: lse' = logaddexp(lse, logsumexp(sink)), out' = out * exp(lse - lse').  A CP runtime passes
   * the sink on exactly one stage (reference: host stage only, dist_attn.py:1278-1280).
   * sink_layout 0 = "sh" [seqlen_sink, num_heads_q], 1 = "ssh" [total_q, seqlen_sink, num_heads_q]. */
  const float* sink;
  int32_t seqlen_sink;
  int32_t sink_layout;
  /* max_logits (NULL = not wanted): fp32 [num_heads_q]; max-reduced IN PLACE with the largest scaled logit
   * softmax_scale * q.k over every unmasked pair of this call (reference: return_max_logits,
   * flex_flash_attn.py:397-403; initialise to -inf for a fresh result). */
  float* max_logits;
} b200_ffa_fwd_params;

typedef struct {
  const void* dout;            /* [total_q, num_heads_q, head_dim] */
  const void* q;
  const void* k;
  const void* v;
  const void* out;             /* GLOBAL (fully merged) forward output, dtype == `dtype` */
  const float* lse;            /* GLOBAL lse [total_q, num_heads_q] */
  /* gradients (caller-allocated). dq is fp32 and is ALWAYS accumulated into (zero it first for a fresh
   * result; reference: dq passed to _flex_flash_attn_backward, flex_flash_attn.py:594-693).
   * dk/dv: dkv_dtype; accumulate_dkv != 0 adds into existing contents (fp32 only). */
  float* dq;                   /* [total_q, num_heads_q, head_dim] fp32 */
  void* dk;                    /* [total_k, num_heads_kv, head_dim] */
  void* dv;
  const int32_t* q_ranges;
  const int32_t* k_ranges;
  const int32_t* attn_type_map;
  int32_t num_slices;
  int32_t total_q, total_k, num_heads_q, num_heads_kv, head_dim;
  int64_t q_stride_t, q_stride_h;
  int64_t k_stride_t, k_stride_h;
  int64_t v_stride_t, v_stride_h;
  int64_t o_stride_t, o_stride_h;
  int64_t do_stride_t, do_stride_h;
  int64_t dq_stride_t, dq_stride_h;
  int64_t dk_stride_t, dk_stride_h;
  int64_t dv_stride_t, dv_stride_h;
  int64_t lse_stride_t, lse_stride_h;
  float softmax_scale;
  float softcap;               /* must be 0 */
  int32_t dtype;               /* of q/k/v/out/dout */
  int32_t dkv_dtype;           /* same as dtype, or B200_F32 */
  int32_t accumulate_dkv;
  int32_t sm_margin;
  void* workspace;
  size_t workspace_bytes;
  /* attention sink gradient (sink NULL = none): dsink (fp32, same layout as sink) is ACCUMULATED into:
   * dsink_j -= sum_rows exp(sink_j - lse_row) * rowsum(out o dout)_row, with `lse` the global lse that includes the
   * sink (reference: sink_bwd, functional/utils.py:117-206).  Zero it first for a fresh result. */
  const float* sink;
  float* dsink;
  int32_t seqlen_sink;
  int32_t sink_layout;
  /* Row statistics shared between calls (NULL = keep them in `workspace`, recomputed every call).  The backward first
   * reduces (out, dout, lse) to two fp32 vectors per (head, row): delta = rowsum(out o dout) and lse * log2(e).  Under
   * context parallelism one backward pass makes one call per stage with the SAME out / dout / lse and different K/V, so the
   * caller may pass a buffer of b200_ffa_bwd_row_stats_bytes(total_q, num_heads_q) bytes (256-byte aligned):
   * row_stats_ready == 0 computes the statistics into it, != 0 reuses what an earlier call of this pass left there
   * (stream-ordered).  The reference recomputes them in every stage call (flex_flash_bwd.hpp preprocess). */
  void* row_stats;
  int32_t row_stats_ready;
} b200_ffa_bwd_params;

/* Library / build identification: "b200_ffa <version> sm_100a". */
B200_API const char* b200_ffa_version(void);
/* Description of the last error on the calling thread ("" if none). */
B200_API const char* b200_ffa_last_error(void);

B200_API size_t b200_ffa_fwd_workspace_bytes(const b200_ffa_fwd_params* p);
B200_API size_t b200_ffa_bwd_workspace_bytes(const b200_ffa_bwd_params* p);
/* size of the optional `row_stats` buffer of b200_ffa_bwd_params */
B200_API size_t b200_ffa_bwd_row_stats_bytes(int32_t total_q, int32_t num_heads_q);
size_t b200_ffa_bwd_row_stats_bytes(int32_t total_q, int32_t num_heads_q);

B200_API int b200_ffa_fwd(const b200_ffa_fwd_params* p, b200_stream_t stream);
B200_API int b200_ffa_bwd(const b200_ffa_bwd_params* p, b200_stream_t stream);

/* In-place LSE merge of a second partial attention result into (out, lse):
 *   lse' = logaddexp(lse, lse2); out' = exp(lse-lse') out + exp(lse2-lse') out2; (-inf,-inf) stays (0,-inf).
 * out/out2: [rows, heads, dim] with the given dtypes (contiguous), lse/lse2: [rows, heads] fp32. */
B200_API int b200_lse_merge(void* out, float* lse, const void* out2, const float* lse2, int64_t rows, int32_t heads,
                   int32_t dim, int32_t out_dtype, int32_t out2_dtype, b200_stream_t stream);

/* In-place attention-sink merge of a finished (out, lse) pair — the sink pass of b200_ffa_fwd on its own:
 *   lse' = logaddexp(lse, logsumexp_j sink_j); out' = out * exp(lse - lse').
 * Replaces correct_attn_out_lse_with_sink (magi_attention/functional/utils.py:634-672).  out: [rows, heads, dim]
 * (element strides given), lse: [rows, heads] fp32, sink fp32 with sink_layout 0 = "sh" [seqlen_sink, heads],
 * 1 = "ssh" [rows, seqlen_sink, heads]; seqlen_sink <= 8. */
B200_API int b200_sink_merge(void* out, float* lse, const float* sink, int64_t rows, int32_t heads, int32_t dim,
                             int32_t seqlen_sink, int32_t sink_layout, int32_t out_dtype, int64_t o_stride_t, int64_t o_stride_h,
                             int64_t lse_stride_t, int64_t lse_stride_h, b200_stream_t stream);

/* dst[dst_ranges[i]] += src[src_ranges[i]] row-span by row-span (fp32 accumulate), row = `row_elems`
 * contiguous elements.  ranges are int32 [n,2] device arrays; spans i have equal length on both sides.
 * Distinct dst spans must not overlap; entries that share the SAME dst span (several sources reduced into one
 * destination) must be adjacent in the lists — they are folded in one race-free pass. */
B200_API int b200_range_reduce_sum(void* dst, int32_t dst_dtype, const void* src, int32_t src_dtype,
                          const int32_t* dst_ranges, const int32_t* src_ranges, int32_t num_ranges,
                          int64_t row_elems, b200_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * NVLink peer-memory exchange primitives — the engine under group_cast / group_reduce for ranks of one NVSwitch
 * domain.  They replace the reference's a2av group collective (magi_attention/comm/primitive/grpcoll/
 * _a2av_grpcoll_impl.py:69-179 group_cast, :228-340 group_reduce; all_to_all_v _all2all_v.py:52-111 and the
 * range_gather / range_reduce passes around it): rows move with copy engines between IPC-mapped arenas, ordering is
 * by 32-bit epoch flags in the arenas.  All calls enqueue on `stream` (except alloc/free/export/import, which are
 * synchronous set-up calls) and return 0 or a negative b200_status.
 * --------------------------------------------------------------------------------------------------------- */
/* cudaMalloc'd, zero-filled arena (IPC-exportable).  Synchronous. */
B200_API int b200_nvl_alloc(void** ptr, size_t bytes);
B200_API int b200_nvl_free(void* ptr);
/* 64-byte IPC handle of an arena / mapping of a peer's arena into this process (lazy peer access). */
B200_API int b200_nvl_export(const void* ptr, uint8_t handle[64]);
B200_API int b200_nvl_import(const uint8_t handle[64], void** ptr);
B200_API int b200_nvl_release(void* ptr);
/* n device-to-device copies dst[i] <- src[i] (bytes[i] each; local or peer-mapped pointers) on the copy engines. */
B200_API int b200_nvl_copy_spans(int32_t n, void* const* dst, const void* const* src, const size_t* bytes,
                                 b200_stream_t stream);
/* After all earlier work of `stream`: *flags[i] = value for i < n (n <= 64; flags may live in peers' arenas). */
B200_API int b200_nvl_signal(void* const* flags, int32_t n, uint32_t value, b200_stream_t stream);
/* Hold `stream` until every *flags[i] >= value (wrap-safe signed difference).  After timeout_ns the wait gives up and
 * sets *err = 1 (err: device-accessible, e.g. pinned host memory) so that a protocol error cannot hang the GPU. */
B200_API int b200_nvl_wait(void* const* flags, int32_t n, uint32_t value, uint32_t* err, uint64_t timeout_ns,
                           b200_stream_t stream);

/* Number of kernels the last b200_ffa_fwd / b200_ffa_bwd call on this thread launched (for bench accounting). */
B200_API int b200_ffa_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200_FFA_H_ */
