/*
 * inferix_b200 — C ABI of the B200-native (sm_100a) block-diffusion denoising hot path.
 *
 * Drop-in boundary for alibaba-damo-academy/Inferix's Self-Forcing / CausVid DiT block forward and the
 * KV-cache append/evict that feeds it.  The reference has no FFI of its own (it is pure Python calling
 * library kernels), so every entry point below names the reference Python call site it replaces
 * (paths relative to the reference root, file:line).  Signatures are plain C: raw device pointers, sizes,
 * a cudaStream_t passed as void*; every function returns an ifx_status and never throws across the ABI.
 *
 * Conventions
 *   - activations are bf16, row-major [tokens, channels]; "ld*" arguments are row strides in ELEMENTS
 *   - all kernels are enqueued on the caller's stream; nothing here synchronises the device
 *   - handles are not thread-safe; one host thread drives one GPU (process per GPU).  Library-internal device state
 *     (opt-in shared-memory attributes, the split-attention partial buffer, arrival counters) is kept PER DEVICE, so a
 *     process may drive several GPUs in turn; it is not kept per stream: launches that use the split-attention path
 *     (ifx_attention* with more work items than whole waves of SMs) must be ordered on ONE stream per device, or the
 *     caller provides the partial buffer itself (ifx_attention_partial + ifx_attention_combine)
 *   - the library is CUDA-only: there is no CPU fallback, calls fail with IFX_ERR_CUDA without a device
 */
#ifndef INFERIX_B200_H_
#define INFERIX_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IFX_ABI_VERSION 1

typedef enum ifx_status {
    IFX_OK = 0,
    IFX_ERR_INVALID = 1,     /* bad argument / shape (reference: Python assert / ValueError) */
    IFX_ERR_BOUNDS = 2,      /* index outside the cache (reference: IndexError on slice assignment) */
    IFX_ERR_HANDLE = 3,      /* unknown / destroyed handle (reference: KeyError) */
    IFX_ERR_OOM = 4,         /* host allocation failed */
    IFX_ERR_CUDA = 5,        /* CUDA runtime / driver error; text via ifx_last_error() */
    IFX_ERR_UNSUPPORTED = 6  /* valid in the reference but outside this build's envelope */
} ifx_status;

/* Human-readable description of the last failing call on this thread ("" if none). */
const char* ifx_last_error(void);
int ifx_abi_version(void);
/* Number of CUDA kernels this library has launched since process start / last reset (bench "gpu_launches"). */
uint64_t ifx_launch_count(void);
void ifx_reset_launch_count(void);

/* Optional per-kernel device timing for bench.py's roofline line: when enabled every kernel launch of this library
 * is bracketed by two cudaEvents on the launch stream.  ifx_prof_read synchronises those events and returns the
 * summed duration and launch count of all kernels whose label starts with `prefix` ("" = everything);
 * ifx_prof_labels writes the distinct labels seen, '\n'-separated.  Off by default (no events, no overhead). */
void ifx_prof_enable(int32_t on);
void ifx_prof_reset(void);
ifx_status ifx_prof_read(const char* prefix, double* total_ms, uint64_t* launches);
ifx_status ifx_prof_labels(char* buf, int32_t cap);

/* ------------------------------------------------------------------------------------------------
 * Paged KV cache with a frame-aligned block table.
 *
 * Replaces the tensor roll + slice-assign in CausalWanSelfAttention.forward
 *   inferix/models/self_forcing/causal_model.py:277-304,328-329   (evict / roll / append / end indices)
 * and the storage of inferix/kvcache_manager/kvcache_manager.py:222-244 (layout (2, nblk, blk, H, D)).
 *
 * The caller owns the memory: two bf16 buffers k_base / v_base of shape [num_pages * page_tokens, heads*head_dim].
 * The handle owns only the block table (logical page -> physical page), the free list and the two end indices
 * the reference keeps in kv_cache_meta["global_end_index"/"local_end_index"].  Eviction rotates the table
 * (0 bytes moved) where the reference copies up to 4*(L-S)*C*2 bytes.  The allocator recycles before it takes fresh
 * pages, so for the windows the reference produces the valid physical pages are the prefix [0, valid_pages) and the
 * attention streams them as one dense extent; any other table is read as runs of consecutive pages
 * (ifx_attention_kv); rows of unmapped pages are never attended, so the buffers may start uninitialised.
 * ------------------------------------------------------------------------------------------------ */
typedef struct ifx_kv ifx_kv;

#define IFX_KV_MAX_PLAN_PAGES 32

typedef struct ifx_kv_plan {
    int64_t local_start;   /* reference local_start_index  (causal_model.py:296,300) */
    int64_t local_end;     /* reference local_end_index    (causal_model.py:294,299) */
    int64_t global_end;    /* value the reference fill_()s into global_end_index (:328) */
    int64_t num_evicted;   /* reference num_evicted_tokens (:287), 0 when no roll happened */
    int32_t num_pages;     /* physical pages receiving the new tokens, in token order */
    int32_t pages[IFX_KV_MAX_PLAN_PAGES];
    int32_t first_offset;  /* token offset inside pages[0] where the new tokens start (0 when frame-aligned) */
} ifx_kv_plan;

/* Direct row access for caches that are only ever written in place (MAGI-1: kvcache_manager/model/
 * magi_kv_cache_manager.py:110-146 reads [0, start) and `set`s [start, start+clip); nothing is evicted).  Extends the
 * block table so that logical tokens [0, tokens) are mapped (identity: logical row i = physical row i) and returns the
 * base pointers of the K and V rows ([token, heads*head_dim] bf16).  A kernel may then write new tokens straight into
 * rows [start, start+n) and attend rows [0, start+n) without any get_range / cat copy.  IFX_ERR_UNSUPPORTED once
 * ifx_kv_plan_append has rotated the table (windowed Wan caches). */
ifx_status ifx_kv_map(ifx_kv* kv, int64_t tokens, void** k_rows, void** v_rows);

ifx_status ifx_kv_create(ifx_kv** out, void* k_base, void* v_base, int32_t num_pages, int32_t page_tokens,
                         int32_t heads, int32_t head_dim);
/* Point the handle at other buffers of the same geometry; table and indices are kept.  This is what the offload tier
 * (reference kvcache_manager.py:222-244: kv_offload -> pinned CPU tensors, copied to the GPU by get()) is built on: the
 * window of a layer lives in pinned host memory and is staged into one of a few device slots before its block runs. */
ifx_status ifx_kv_rebind(ifx_kv* kv, void* k_base, void* v_base);
ifx_status ifx_kv_destroy(ifx_kv* kv);
/* Reference: pipeline resets global_end_index/local_end_index to 0 (CausalInferencePipeline.py:193-199). */
ifx_status ifx_kv_reset(ifx_kv* kv);
/* Index arithmetic of causal_model.py:277-300 on host integers (no device sync), plus the table rotation.
 * `windowed` = (local_attn_size != -1).  Commits the new end indices (reference :328-329). */
ifx_status ifx_kv_plan_append(ifx_kv* kv, int64_t current_start, int64_t num_new_tokens, int64_t sink_tokens,
                              int32_t windowed, ifx_kv_plan* plan);
/* Snapshot of the state: end indices and the logical->physical table (table_out may be NULL). */
ifx_status ifx_kv_state(const ifx_kv* kv, int64_t* global_end, int64_t* local_end, int32_t* valid_pages,
                        int32_t* table_out, int32_t table_cap);
/* Gather tokens [start, start+length) in the reference's logical order into contiguous [length, H*D] buffers
 * (kvcache_manager.py:145-169 get / get_range; self_forcing_kv_cache_manager.py:112-127).  16-byte vector
 * loads through the block table; dst_k / dst_v may be NULL to skip one side. */
ifx_status ifx_kv_export(const ifx_kv* kv, void* dst_k, void* dst_v, int64_t start, int64_t length, void* stream);
/* Scatter contiguous tokens into logical positions [start, start+length)
 * (kvcache_manager.py:192-220 set; self_forcing_kv_cache_manager.py:129-145).  Extends local_end if needed. */
ifx_status ifx_kv_import(ifx_kv* kv, const void* src_k, const void* src_v, int64_t start, int64_t length,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-op kernels of the DiT block.
 * ------------------------------------------------------------------------------------------------ */

/* LayerNorm (fp32 statistics, eps) followed by AdaLN modulation, one rounding to bf16 per reference op:
 *   out = bf16(bf16(bf16(LN(x)) * bf16(1 + scale[f])) + shift[f]),  f = row / tokens_per_frame
 * Replaces  norm1/norm2 + modulation  causal_model.py:433,451-452  (WanLayerNorm components.py:129-142).
 * With ln_weight/ln_bias non-NULL and scale/shift NULL it is the affine norm3 of causal_model.py:448.
 * scale/shift point at bf16 vectors of `cols`, frame f at  ptr + f * mod_frame_stride  (elements). */
ifx_status ifx_ln_modulate(const void* x, void* out, const void* ln_weight, const void* ln_bias, const void* shift,
                           const void* scale, int64_t mod_frame_stride, int64_t rows, int32_t cols,
                           int64_t tokens_per_frame, float eps, void* stream);

typedef enum ifx_epilogue {
    IFX_EPI_BIAS = 0,          /* out = bf16(acc + bias)                                  nn.Linear            */
    IFX_EPI_BIAS_GELU = 1,     /* out = bf16(gelu_tanh(bf16(acc + bias)))                 ffn[0:2]  :377-379   */
    IFX_EPI_BIAS_GATE_RES = 2, /* out = bf16(res + bf16(bf16(acc + bias) * gate[f]))      :444, :455-456;      */
                               /* gate == NULL -> out = bf16(res + bf16(acc + bias))      cross-attn  :448     */
    IFX_EPI_BIAS_GELU_ERF = 3, /* out = bf16(gelu_erf(bf16(acc + bias)))   MAGI CustomMLP, dit_module.py:551   */
    IFX_EPI_BIAS_F32 = 4       /* out = fp32(acc + bias): `out` is float*, ldo counts floats (16-byte aligned     */
                               /* rows); MAGI linear_proj under autocast(float32), dit_module.py:1291-1293     */
} ifx_epilogue;

/* out[M,N] = epilogue(A[M,K] @ W[N,K]^T): tcgen05 BF16 MMA, FP32 accumulation in TMEM, TMA-fed pipeline.
 * Replaces the cuBLAS nn.Linear calls of causal_model.py:171-175,333,378-379 and wan_base/model.py:77,98
 * together with the eager bias / GELU / gate / residual ops around them.
 * K % 8 == 0, N % 8 == 0; all pointers 16-byte aligned. `residual` may alias `out`. */
ifx_status ifx_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* out,
                         int64_t ldo, int64_t M, int32_t N, int32_t K, int32_t epilogue, const void* residual,
                         int64_t ldr, const void* gate, int64_t gate_frame_stride, int64_t tokens_per_frame,
                         void* stream);

/* ---- FP8 (e4m3) linears, per-tensor scales.  In-tree specification: MAGI's PerTensorQuantizedFp8Linear +
 * div_clamp_to (models/magi/dit/dit_module.py:367-387,434-459):
 *     x_q = e4m3( bf16( clamp(float(x) / input_scale, -448, 448) ) )
 *     y   = bf16( (x_q @ W_q^T) * (input_scale * weight_scale) [+ bias] )            (flashinfer.bmm_fp8)
 * The Wan quantization examples call DAX (not vendored): "DAX parity unpinned". */

/* Activation quantisation x[rows, cols] bf16 -> e4m3 with a static per-tensor scale. */
ifx_status ifx_quantize_fp8(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int32_t cols,
                            float scale, void* stream);
/* Same with one scale per input channel: x_q[m, k] = e4m3( bf16( clamp(x[m, k] / col_scale[k], -448, 448) ) ).  MAGI's
 * PerTensorQuantizedFp8Linear keeps input_scale as an [in_features] vector and PerChannelQuantizedFp8Linear divides by
 * its smooth_scale [1, in_features] (dit_module.py:434-490); both then call bmm_fp8 with per-tensor scales
 * (ifx_gemm_fp8).  col_scale: fp32 [cols], 16-byte aligned. */
ifx_status ifx_quantize_fp8_cols(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int32_t cols,
                                 const float* col_scale, void* stream);
/* ifx_ln_modulate whose result is quantised on the way out (out: e4m3 [rows, cols]) — the quantisation of the
 * QKV / cross-q / FFN1 GEMM inputs fused into the kernel that produces them. */
ifx_status ifx_ln_modulate_fp8(const void* x, void* out, const void* ln_weight, const void* ln_bias, const void* shift,
                               const void* scale, int64_t mod_frame_stride, int64_t rows, int32_t cols,
                               int64_t tokens_per_frame, float eps, float out_scale, void* stream);
/* out[M,N] = epilogue((A_q[M,K] @ W_q[N,K]^T) * alpha + bias): tcgen05 kind::f8f6f4 (K = 32 per MMA, twice the
 * bf16 rate), fp32 accumulation; dequantisation (alpha = input_scale * weight_scale), bias and the same three
 * epilogues as ifx_gemm_bf16 are fused.  A_q / W_q are e4m3 bytes, K % 16 == 0; bias / residual / gate / out bf16. */
ifx_status ifx_gemm_fp8(const void* A, int64_t lda, const void* W, int64_t ldw, float alpha, const void* bias, void* out,
                        int64_t ldo, int64_t M, int32_t N, int32_t K, int32_t epilogue, const void* residual,
                        int64_t ldr, const void* gate, int64_t gate_frame_stride, int64_t tokens_per_frame,
                        void* stream);

/* ---- Dynamic 8-bit linears: per-token activation scales x per-channel weight scales, FP8 (e4m3) or INT8.
 * This is the qconfig the reference's quantisation examples request (example/quantization/run_causvid_quantized.py:32-37,
 * run_self_forcing_quantized.py: get_dynamic_fp8_per_token_act_per_channel_weight_qconfig + quantize_dynamic from DAX).
 * DAX is a dependency that is not vendored in the reference (git clone in its README): the published algorithm is
 * restated — "DAX parity unpinned" (oracle/wan_oracle.py: dynamic_q8_linear):
 *     s_a[m] = max(|x[m, :]|, 1e-12) / qmax        s_w[n] = max(|W[n, :]|, 1e-12) / qmax      qmax = 448 | 127
 *     x_q = rne_sat(x / s_a)   W_q = rne_sat(W / s_w)   y = bf16( (x_q @ W_q^T) * s_a[m] * s_w[n] + bias )
 * The activation scales never touch the host: they are produced by the kernel that produces the activation and consumed
 * by the GEMM epilogue. */
#define IFX_Q8_E4M3 0
#define IFX_Q8_INT8 1
/* x[rows, cols] bf16 -> out[rows, cols] 8-bit codes + scales[rows] (fp32). */
ifx_status ifx_quantize_rows(const void* x, int64_t ldx, void* out, int64_t ldo, float* scales, int64_t rows,
                             int32_t cols, int32_t kind, void* stream);
/* ifx_ln_modulate whose result is quantised per token on the way out (the QKV / cross-q / FFN1 GEMM inputs). */
ifx_status ifx_ln_modulate_quant(const void* x, void* out, float* scales, const void* ln_weight, const void* ln_bias,
                                 const void* shift, const void* scale, int64_t mod_frame_stride, int64_t rows,
                                 int32_t cols, int64_t tokens_per_frame, float eps, int32_t kind, void* stream);
/* out[M,N] = epilogue((A_q[M,K] @ W_q[N,K]^T) * row_scale[m] * col_scale[n] + bias): tcgen05 kind::f8f6f4 (e4m3) or
 * kind::i8 (S8 x S8 -> S32 accumulators in TMEM); dequantisation, bias and the ifx_gemm_bf16 epilogues fused. K % 16 == 0. */
ifx_status ifx_gemm_q8(const void* A, int64_t lda, const void* W, int64_t ldw, const float* row_scale,
                       const float* col_scale, int32_t kind, const void* bias, void* out, int64_t ldo, int64_t M,
                       int32_t N, int32_t K, int32_t epilogue, const void* residual, int64_t ldr, const void* gate,
                       int64_t gate_frame_stride, int64_t tokens_per_frame, void* stream);

/* 3-D RoPE table: complex128 [1024, head_dim/2] exactly as CausalWanModel.freqs (causal_model.py:634-641,
 * rope_params components.py:34-52), uploaded once by the host as interleaved (cos, sin) doubles. */
typedef struct ifx_rope_grid {
    int32_t frames;        /* F of grid_sizes */
    int32_t height;        /* H of grid_sizes (patch grid) */
    int32_t width;         /* W of grid_sizes */
    int32_t start_frame;   /* current_start // frame_seqlen   (causal_model.py:256) */
    int32_t hw_offset;     /* first h*w index owned by this rank (sequence parallel; causal_model.py:88) */
    int32_t hw_count;      /* h*w tokens per frame owned by this rank (= height*width when world_size == 1) */
} ifx_rope_grid;

/* Fused  QK RMSNorm (full-C, fp32, components.py:107-126) -> x weight -> 3-D RoPE in fp64
 * (causal_rope_apply causal_model.py:33-61 / _chunked :64-100) -> bf16, and the KV append of :303-304:
 * q goes to q_out[rows, C]; roped k and raw v go straight into the cache pages named by `plan`.
 * qkv is the [rows, 3C] output of the fused q|k|v projection.  k_dst/v_dst override the cache destination
 * (contiguous [rows, C] staging for the sequence-parallel all-gather) when kv == NULL. */
ifx_status ifx_qk_norm_rope_append(const void* qkv, int64_t ld_qkv, const void* norm_q_weight,
                                   const void* norm_k_weight, const double* freqs, const ifx_rope_grid* grid,
                                   void* q_out, int64_t ld_q, ifx_kv* kv, const ifx_kv_plan* plan, void* k_dst,
                                   void* v_dst, int64_t rows, int32_t heads, int32_t head_dim, float eps,
                                   void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Sequence parallel over peer memory (one NVSwitch box, <= 8 ranks).  Replaces the per-layer exchange of the
 * reference (ring P2P + LSE merges, models/attention/distributed.py:564-712) and this library's own NCCL variant
 * (staging -> all-gather -> ifx_kv_append_sp): the fused norm + RoPE kernel stores this rank's new K / V rows into the
 * same cache rows of EVERY rank's replicated cache with 16-byte P2P stores over NVLink and then publishes an epoch
 * number in every rank's flag array; ifx_peer_wait makes the attention that follows wait (on the stream) until all
 * ranks have published.  Write-after-read safety: a rank reaches the next write of a layer's rows only after passing a
 * later wait (next layer, or the head all-gather), which every rank signals after its attention of this layer.
 * ------------------------------------------------------------------------------------------------------------ */
#define IFX_MAX_PEERS 8
#define IFX_PEER_HANDLE_BYTES 64

typedef struct ifx_peer_dst {
    int32_t world, rank;
    void* k[IFX_MAX_PEERS];          /* this layer's cache K rows on every rank (entry `rank` = the local cache) */
    void* v[IFX_MAX_PEERS];
    int64_t* flags[IFX_MAX_PEERS];   /* every rank's flag array, int64[world]; this rank writes element `rank` */
    int64_t epoch;                   /* > 0, increases by one per call, identical on all ranks */
    int32_t local_only;              /* ifx_qk_norm_rope_append_peers: 1 = write this rank's cache only and publish  */
                                     /* nothing (the exchange then happens in ifx_peer_push or, in the shipped path,  */
                                     /* inside the attention kernel of ifx_wan_block_forward_sp / IFX_SP_OVERLAP)     */
} ifx_peer_dst;

/* CUDA IPC: handle of the allocation containing ptr (+ ptr's offset inside it), to be shipped to the other processes
 * of the box; ifx_peer_open maps a received handle and returns the allocation's base in this process. */
ifx_status ifx_peer_export(const void* ptr, void* handle /* IFX_PEER_HANDLE_BYTES */, int64_t* offset, uint64_t* base,
                           uint64_t* size);
ifx_status ifx_peer_open(const void* handle, void** base);
ifx_status ifx_peer_close(void* base);

/* ifx_qk_norm_rope_append for rank `peers->rank` of a sequence-parallel group: rows = frames * hw_count local tokens,
 * plan = the (identical on every rank) plan of the whole block (rows * world tokens). */
ifx_status ifx_qk_norm_rope_append_peers(const void* qkv, int64_t ld_qkv, const void* norm_q_weight,
                                         const void* norm_k_weight, const double* freqs, const ifx_rope_grid* grid,
                                         void* q_out, int64_t ld_q, ifx_kv* kv, const ifx_kv_plan* plan,
                                         const ifx_peer_dst* peers, int64_t rows, int32_t heads, int32_t head_dim,
                                         float eps, void* stream);
/* Exchange half of ifx_qk_norm_rope_append_peers as its own small grid (`ctas` CTAs of 1024 threads): copies this rank's
 * rows of the block's new pages from its cache to every other rank's cache and publishes `peers->epoch`.  Meant for a
 * side stream, next to the attention over the pages that were already cached (which needs no new K / V); the
 * attention over the new pages then follows an ifx_peer_wait.  A/B variant only: measured at 106 GB/s on the SMs the
 * attention leaves free (8 CTAs) and it delays the attention's tail, so the shipped path copies from inside the
 * attention kernel instead (IFX_SP_OVERLAP below). */
ifx_status ifx_peer_push(ifx_kv* kv, const ifx_kv_plan* plan, const ifx_peer_dst* peers, int32_t frames, int32_t chunk,
                         int32_t ctas, void* stream);
/* Stream-ordered wait until flags[s] >= epoch for every s < world; traps after timeout_ms instead of hanging. */
ifx_status ifx_peer_wait(const int64_t* flags, int32_t world, int64_t epoch, int32_t timeout_ms, void* stream);

/* Copy already-normalised K / V rows (e.g. all-gathered from peers) into the pages named by `plan`. */
ifx_status ifx_kv_append(ifx_kv* kv, const ifx_kv_plan* plan, const void* k_src, const void* v_src, int64_t ld_src,
                         int64_t rows, void* stream);

/* Sequence-parallel append: k_src / v_src hold the all-gathered new rows rank-major, [world, frames*chunk, H*D]
 * with `src_rank_stride` elements between consecutive ranks (so K and V may come out of ONE all-gather of a
 * [world, 2, rows, H*D] buffer).  Rank r owns hw indices [r*chunk, (r+1)*chunk) of every frame
 * (causal_model.py:939-942); rows are written in the single-process token order (frame, rank, hw)
 * == 'b (cp f hw) c -> b (f cp hw) c' of causal_model.py:1018. */
ifx_status ifx_kv_append_sp(ifx_kv* kv, const ifx_kv_plan* plan, const void* k_src, const void* v_src,
                            int64_t src_rank_stride, int32_t world, int32_t frames, int32_t chunk, void* stream);

/* WanRMSNorm on its own (cross-attention q / text k, wan_base/model.py:77,82): out = bf16(bf16(x*rsqrt(ms+eps))*w) */
ifx_status ifx_rmsnorm(const void* x, int64_t ldx, const void* weight, void* out, int64_t ldo, int64_t rows,
                       int32_t cols, float eps, void* stream);

/* Non-causal multi-head attention  out = softmax(q k^T * scale) v  with Lq = rows of q, Lk keys.
 * tcgen05 flash attention: S and O accumulate in TMEM, P is fed back from TMEM, K/V tiles arrive by TMA.
 * Replaces attention()/flash_attention()  inferix/models/attention/flash_attention.py:42-200 as called from
 * causal_model.py:307-315 (self-attention over cache[0:local_end]) and wan_base/model.py:94-95 (cross-attn).
 * q/k/v/out are [tokens, heads*head_dim] with row strides ld*; head_dim must be 128. */
ifx_status ifx_attention(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* out,
                         int64_t ldo, int64_t q_rows, int64_t kv_rows, int32_t heads, int32_t head_dim,
                         float softmax_scale, void* stream);
/* Grouped-query variant: q has `heads` heads, k/v have `kv_heads` (heads % kv_heads == 0); query head h reads
 * key/value head h / (heads / kv_heads).  MAGI-1's FullyParallelAttention (models/magi/dit/dit_module.py:833-1014:
 * 24 or 48 query heads over 8 KV groups); the per-range loop of :1000-1014 is one call per (q_range, k_range) row
 * with offset pointers. */
ifx_status ifx_attention_gqa(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* out,
                             int64_t ldo, int64_t q_rows, int64_t kv_rows, int32_t heads, int32_t kv_heads,
                             int32_t head_dim, float softmax_scale, void* stream);
/* Two-phase attention for overlapping the sequence-parallel K/V exchange with compute.  ifx_attention_partial attends
 * the queries to a LIST of key-row extents (up to 4 [row0, rows) pairs inside k/v[0:kv_rows_total)) and writes, for
 * every (head, 256-row pair) item, un-normalised partials (O, max, sum) into `workspace` at piece slots
 * [piece_first, piece_first + piece_count) of `pieces_per_item`; the extent list is cut into piece_count key ranges,
 * one CTA each.  ifx_attention_combine merges the pieces into out = O / l.  Usage: phase 1 = pages already in the
 * cache (runs while the all-gather of the new K/V is in flight), phase 2 = the block's own pages, then combine.
 * workspace bytes = items * pieces_per_item * 256 * 130 * 4, items = heads * ceil(q_rows / 256). */
ifx_status ifx_attention_partial(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                                 int64_t kv_rows_total, const int64_t* extents, int32_t n_ext, int64_t q_rows,
                                 int32_t heads, int32_t kv_heads, int32_t head_dim, float softmax_scale,
                                 void* workspace, int64_t workspace_bytes, int32_t pieces_per_item,
                                 int32_t piece_first, int32_t piece_count, void* stream);
ifx_status ifx_attention_combine(const void* workspace, int32_t pieces_per_item, void* out, int64_t ldo,
                                 int64_t q_rows, int32_t heads, int32_t head_dim, void* stream);
/* Same as ifx_attention_gqa, also returning the log-sum-exp of the scaled scores, lse[h * q_rows + r] (fp32, natural
 * log) — the (out, lse) pair the reference's attention backends return for LSE merging
 * (models/attention/backends.py:58-72 flash_attn_forward; distributed.py:27-46 update_out_and_lse). */
ifx_status ifx_attention_lse(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* out,
                             int64_t ldo, float* lse, int64_t q_rows, int64_t kv_rows, int32_t heads, int32_t kv_heads,
                             int32_t head_dim, float softmax_scale, void* stream);

/* Attention over a LIST of key-row extents of k/v[0:kv_rows_total): up to IFX_ATTN_MAX_EXTENTS [row0, rows) pairs
 * (runs of physically consecutive cache pages).  Rows that follow an extent in memory are never attended: their
 * scores are masked and their V rows are zeroed in shared memory before the P V product, so unmapped pages may hold
 * anything (NaN included).  This is how a paged cache whose valid pages are not a physical prefix is read. */
#define IFX_ATTN_MAX_EXTENTS 32
ifx_status ifx_attention_extents(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                                 int64_t kv_rows_total, const int64_t* extents, int32_t n_ext, void* out, int64_t ldo,
                                 int64_t q_rows, int32_t heads, int32_t kv_heads, int32_t head_dim,
                                 float softmax_scale, void* stream);
/* Keys / values taken through the block table of a paged cache: the valid logical pages [0, local_end / page_tokens)
 * are visited in PHYSICAL order (softmax attention does not depend on the key order) as runs of consecutive pages;
 * when they form the physical prefix — what the allocator produces for every window the reference can express — the
 * whole window is one dense TMA extent.  IFX_ERR_UNSUPPORTED beyond IFX_ATTN_MAX_EXTENTS runs. */
ifx_status ifx_attention_kv(const void* q, int64_t ldq, const ifx_kv* kv, void* out, int64_t ldo, int64_t q_rows,
                            float softmax_scale, void* stream);
/* Sequence-parallel form: the pages named by `fresh` (the plan of the block being appended) are still being stored
 * by the peer ranks when the kernel starts.  Every CTA attends its share of the other pages first; its TMA producer
 * thread then acquires the `world` epoch flags (int64[world], system scope, >= epoch; traps after timeout_ms) and only
 * then loads the fresh pages.  The exchange is hidden behind the attention over the cached window instead of sitting
 * in front of it (ifx_peer_wait).  Replaces the ring P2P + LSE merge of models/attention/distributed.py:564-712. */
ifx_status ifx_attention_kv_wait(const void* q, int64_t ldq, const ifx_kv* kv, const ifx_kv_plan* fresh,
                                 const int64_t* flags, int32_t world, int64_t epoch, int32_t timeout_ms, void* out,
                                 int64_t ldo, int64_t q_rows, float softmax_scale, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * MAGI-1 transformer layer row kernels (inferix/models/magi/dit/dit_module.py).  head_dim must be 128.
 * ------------------------------------------------------------------------------------------------------------ */

/* One pass over the fused projection qkvx [rows, ld] = q | k | v | qx (q, qx: q_heads*128 columns; k, v: kv_heads*128):
 *   q, k : .float() -> per-head LayerNorm with fp32 affine (FusedLayerNorm, :358-360) -> rotary on dims
 *          [0, 2*rotary_half) pairing j with j + rotary_half (flash_attn apply_rotary_emb, non-interleaved) -> bf16
 *          (get_q / get_k, :902-934); q -> q_out, k -> k_dst
 *   v    : raw copy -> v_dst (get_v, :936-938)
 *   qx   : per-head LayerNorm with bf16 affine -> qx_out (get_xqkv, :954-958)
 * rope [rows, ld_rope] fp32 = sin[rotary_half] | cos[rotary_half] per token (rotary_pos_emb.tensor_split, :1097).
 * k_dst / v_dst are row pointers into the layer's KV rows (ld_kv elements apart): the concatenation of key and value in
 * get_kv (:940-945) and the cache `set` of magi_kv_cache_manager.py:138-146 become this one write.
 * Head groups (Ulysses context parallel, distributed/parallelism/context_parallel.py:382-402): q head h goes to
 * q_out + (h / q_group_heads) * q_group_stride + row * ld_q + (h % q_group_heads) * 128 (k / v likewise), which is
 * the "(cp seq) hn hd" all-to-all send layout; pass q_group_heads = q_heads, kv_group_heads = kv_heads and zero
 * strides for the plain token-major layout. */
ifx_status ifx_magi_qkv_post(const void* qkvx, int64_t ld, int64_t rows, int32_t q_heads, int32_t kv_heads,
                             int32_t head_dim, const float* q_ln_w, const float* q_ln_b, const float* k_ln_w,
                             const float* k_ln_b, const void* qx_ln_w, const void* qx_ln_b, const float* rope,
                             int64_t ld_rope, int32_t rotary_half, float eps, void* q_out, int64_t ld_q,
                             int32_t q_group_heads, int64_t q_group_stride, void* k_dst, void* v_dst, int64_t ld_kv,
                             int32_t kv_group_heads, int64_t kv_group_stride, void* qx_out, int64_t ld_qx,
                             void* stream);

/* Per-head LayerNorm (bf16 in / affine / out, fp32 statistics) on [rows, heads, 128]: k_layernorm_xattn on the caption
 * keys (dit_module.py:968).  May run in place. */
ifx_status ifx_head_layernorm(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int32_t heads,
                              int32_t head_dim, const void* weight, const void* bias, float eps, void* stream);

/* bias_modulate_add (dit_module.py:295-313): out[r] = bf16( LN_fp32( x[r] * gate[row_map[r]] ) * norm_w + norm_b +
 * residual[r] ), i.e. range_mod_triton (:205-292) + fp32 FusedLayerNorm + residual add in one pass.
 * gate bf16 [num_gates, cols] (softcapped AdaModulateLayer output half), row_map int32 [rows] on the device
 * (condition_map), norm_w / norm_b fp32 [cols].  x is bf16 (the MLP branch) or, with x_is_f32, the fp32 result of the
 * output projection (IFX_EPI_BIAS_F32; dit_module.py:1291-1293 runs it under autocast(float32)); ldx counts elements
 * of x's type.  out (bf16) may alias residual, and x when x is bf16. */
ifx_status ifx_gate_norm_residual(const void* x, int64_t ldx, int32_t x_is_f32, const void* gate, int64_t gate_stride,
                                  int32_t num_gates, const int32_t* row_map, const float* norm_w, const float* norm_b,
                                  const void* residual, int64_t ldr, void* out, int64_t ldo, int64_t rows,
                                  int32_t cols, float eps, void* stream);

/* flashinfer.activation.silu_and_mul (dit_module.py:549): out[r, j] = bf16( silu(x[r, j]) * x[r, cols_out + j] ). */
ifx_status ifx_silu_mul(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int32_t cols_out,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Prologue / epilogue of one DiT forward: the ~45 eager ops between the latent tensor and the first
 * block, and between the last block and the denoised latent, as five small kernels.
 * ------------------------------------------------------------------------------------------------ */

/* Patch gather: x is the latent [c_in, F, H, W] addressed through element strides (any permuted view); out receives
 * this rank's token rows [frames * hw_count, c_in*pt*ph*pw] bf16, column k = ((c*pt + dt)*ph + dy)*pw + dx — the
 * flattening of Conv3d's weight — so that ifx_gemm_bf16(out, weight.view(dim, k), bias) is the patch embedding of
 * causal_model.py:916-921 with the sequence-parallel scatter of :939-942 folded in. */
ifx_status ifx_patchify(const void* x, int64_t stride_c, int64_t stride_f, int64_t stride_h, int64_t stride_w,
                        int32_t c_in, int32_t pt, int32_t ph, int32_t pw, int32_t frames, int32_t grid_h, int32_t grid_w,
                        int32_t hw_offset, int32_t hw_count, void* out, void* stream);
/* sinusoidal_embedding_1d (wan_base/components.py:11-31): out[n, dim] = bf16([cos | sin](pos * 10000^(-j/half))), fp64. */
ifx_status ifx_sinusoidal_embedding(const double* positions, int32_t n, int32_t dim, void* out, void* stream);
/* nn.Linear on M <= 8 rows (the time MLP of causal_model.py:922-936): out[m, :] = bf16(in[m, :] @ W^T + b), optional
 * SiLU (rounded to bf16) on the input.  With mod_table (bf16 [layers, N]) the result is not stored as such: out is
 * [layers, M, N] and receives bf16(mod_table[l, n] + bf16(acc + b[n])) — every layer's `modulation + e0` (:412). */
ifx_status ifx_linear_small(const void* x, int64_t ldx, const void* w, int64_t ldw, const void* bias, void* out,
                            int64_t ldo, int32_t M, int32_t N, int32_t K, int32_t silu_input, const void* mod_table,
                            int32_t layers, int64_t mod_layer_stride, int64_t out_layer_stride, void* stream);
/* Unpatchify (causal_model.py:1196-1219) + flow -> x0 (wrapper.py:259-283): head_tokens [F*grid_h*grid_w, ph*pw*C]
 * -> flow_out (optional) and x0_out, both contiguous [F, C, grid_h*ph, grid_w*pw] bf16;
 * x0 = bf16(x_t - sigma_t * flow) in fp64 with sigma_t = sigmas[argmin_i |timesteps[i] - t_f|].  xt is addressed
 * through element strides.  timestep: fp64 [frames] on the device (the reference's timestep tensor, `.double()`). */
ifx_status ifx_unpatchify_x0(const void* head_tokens, const void* xt, int64_t xt_stride_f, int64_t xt_stride_c,
                             int64_t xt_stride_h, int64_t xt_stride_w, const double* timestep,
                             const float* table_timesteps, const float* table_sigmas, int32_t table_len, int32_t frames,
                             int32_t channels, int32_t grid_h, int32_t grid_w, int32_t ph, int32_t pw, void* flow_out,
                             void* x0_out, void* stream);
/* FlowMatchScheduler.add_noise (schedulers/flow_match.py:159-176): out = bf16((1 - sigma_f) * x0 + sigma_f * noise),
 * fp32 arithmetic, contiguous [frames, per_frame] bf16 tensors. */
ifx_status ifx_add_noise(const void* x0, const void* noise, const double* timestep, const float* table_timesteps,
                         const float* table_sigmas, int32_t table_len, int32_t frames, int64_t per_frame, void* out,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * Whole DiT block: CausalWanAttentionBlock.forward  causal_model.py:384-484  in one call (13 launches).
 * ------------------------------------------------------------------------------------------------ */
typedef struct ifx_wan_block_weights {
    int32_t dim, ffn_dim, heads, head_dim;
    float eps;
    const void* qkv_w;      /* [3*dim, dim]  rows: q | k | v  (self_attn.q/k/v.weight concatenated) */
    const void* qkv_b;      /* [3*dim] */
    const void* norm_q_w;   /* [dim] */
    const void* norm_k_w;   /* [dim] */
    const void* o_w;        /* [dim, dim] */
    const void* o_b;
    const void* norm3_w;    /* [dim] affine LN before cross-attn (NULL: identity, cross_attn_norm=False) */
    const void* norm3_b;
    const void* cq_w;       /* cross_attn.q */
    const void* cq_b;
    const void* cnorm_q_w;  /* cross_attn.norm_q */
    const void* co_w;       /* cross_attn.o */
    const void* co_b;
    const void* ffn1_w;     /* [ffn_dim, dim] */
    const void* ffn1_b;
    const void* ffn2_w;     /* [dim, ffn_dim] */
    const void* ffn2_b;
} ifx_wan_block_weights;

typedef struct ifx_wan_block_io {
    void* x;                    /* [rows, dim] residual stream, updated in place */
    int64_t rows;
    int64_t tokens_per_frame;   /* frame_seqlen (per rank) */
    const void* mod;            /* bf16 [frames, 6, dim] = modulation + e   (causal_model.py:412) */
    const double* freqs;
    ifx_rope_grid grid;
    ifx_kv* kv;                 /* self-attention cache of this layer */
    int64_t current_start;      /* token offset of this block, as the reference passes it */
    int64_t sink_tokens;
    int32_t windowed;
    const void* cross_k;        /* [text_len, dim] cached text K (RMS-normed) */
    const void* cross_v;
    int64_t text_len;
    /* scratch, all bf16, caller-allocated: */
    void* ws_h;                 /* [rows, dim] */
    void* ws_qkv;               /* [rows, 3*dim] */
    void* ws_q;                 /* [rows, dim] */
    void* ws_attn;              /* [rows, dim] */
    void* ws_ffn;               /* [rows, ffn_dim] */
} ifx_wan_block_io;

ifx_status ifx_wan_block_forward(const ifx_wan_block_weights* w, const ifx_wan_block_io* io, ifx_kv_plan* plan_out,
                                 void* stream);

/* The same block for rank `peers->rank` of a sequence-parallel group (causal_model.py:939-942: io->rows local tokens =
 * frames * hw/P, io->tokens_per_frame = hw/P, io->grid.hw_offset / hw_count name the shard; the cache is replicated and
 * the plan covers rows * world tokens).  One call per layer, no collective:
 *   mode IFX_SP_STORE   : the norm + RoPE kernel stores K / V into every rank's cache, ifx_peer_wait, attention.
 *   mode IFX_SP_OVERLAP : the norm + RoPE kernel writes the local cache only; the ATTENTION KERNEL itself performs the
 *                         exchange: an otherwise idle warp of each of its first CTAs (at most `push_ctas`, 0 = default
 *                         32) copies a slice of this rank's new rows to the same cache rows of every peer
 *                         over NVLink and the last one publishes the epoch, while the MMA / softmax warps attend the
 *                         cached pages; every CTA acquires the peers' epoch flags only before its first fresh-page tile
 *                         (ifx_attention_kv_wait).  Compute and collective are one kernel; the exchange costs no SM and
 *                         no time on the critical path. */
#define IFX_SP_STORE 0
#define IFX_SP_OVERLAP 1
ifx_status ifx_wan_block_forward_sp(const ifx_wan_block_weights* w, const ifx_wan_block_io* io,
                                    const ifx_peer_dst* peers, int32_t mode, int32_t push_ctas, int32_t timeout_ms,
                                    ifx_kv_plan* plan_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* INFERIX_B200_H_ */
