// ABI plumbing: error text, launch counter, TMA tensor-map encoding, and the whole-block entry point that
// strings the 13 kernels of CausalWanAttentionBlock.forward (causal_model.py:384-484) together.
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <string>

#include "ifx_internal.h"

namespace ifx {

// ---------------------------------------------------------------- optional per-kernel timing
struct ProfRecord {
    std::string label;
    cudaEvent_t start, stop;
};
static bool g_prof_on = false;
static std::vector<ProfRecord> g_prof;

ProfScope::ProfScope(const char* label, cudaStream_t s) : slot(-1), stream(s) {
    if (!g_prof_on) return;
    ProfRecord r;
    r.label = label;
    if (cudaEventCreate(&r.start) != cudaSuccess) return;
    if (cudaEventCreate(&r.stop) != cudaSuccess) {
        cudaEventDestroy(r.start);
        return;
    }
    cudaEventRecord(r.start, s);
    g_prof.push_back(r);
    slot = static_cast<int>(g_prof.size()) - 1;
}
ProfScope::~ProfScope() {
    if (slot >= 0) cudaEventRecord(g_prof[slot].stop, stream);
}

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

ifx_status set_error(ifx_status code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}
void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("IFX_PDL");
        v = (e == nullptr || e[0] != '0') ? 1 : 0;
    }
    return v == 1;
}

int sm_count() {
    static int cached[64] = {0};       // per device: one process may drive several GPUs (tools/nvlink_probe.py)
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached[dev] = n;
        else
            cached[dev] = 148;
    }
    return cached[dev];
}

// Zero-initialised device counter of the current device, allocated on first use (arrival counters of kernels whose
// last CTA does something; launches on one device are stream-ordered by the callers).
ifx_status device_counter(unsigned int* (&slots)[64], unsigned int** out) {
    int dev = 0;
    IFX_CUDA_OK(cudaGetDevice(&dev));
    IFX_CHECK_ARG(dev >= 0 && dev < 64, "device index %d beyond 63", dev);
    if (!slots[dev]) {
        IFX_CUDA_OK(cudaMalloc(&slots[dev], sizeof(unsigned int)));
        IFX_CUDA_OK(cudaMemset(slots[dev], 0, sizeof(unsigned int)));
    }
    *out = slots[dev];
    return IFX_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static ifx_status make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dt, int elem_bytes, const void* base,
                               uint64_t inner_elems, uint64_t outer_rows, uint64_t row_stride_elems,
                               uint32_t box_inner, uint32_t box_rows);

ifx_status make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner_elems, uint64_t outer_rows,
                             uint64_t row_stride_elems, uint32_t box_inner, uint32_t box_rows) {
    return make_tmap_2d(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, inner_elems, outer_rows, row_stride_elems,
                        box_inner, box_rows);
}
ifx_status make_tmap_u8_2d(CUtensorMap* out, const void* base, uint64_t inner_elems, uint64_t outer_rows,
                           uint64_t row_stride_elems, uint32_t box_inner, uint32_t box_rows) {
    return make_tmap_2d(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, base, inner_elems, outer_rows, row_stride_elems,
                        box_inner, box_rows);
}

static ifx_status make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dt, int elem_bytes, const void* base,
                               uint64_t inner_elems, uint64_t outer_rows, uint64_t row_stride_elems,
                               uint32_t box_inner, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(IFX_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    cuuint64_t dims[2] = {inner_elems, outer_rows};
    cuuint64_t strides[1] = {row_stride_elems * static_cast<uint64_t>(elem_bytes)};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(IFX_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): inner=%llu rows=%llu stride=%llu", (int)r,
                         (unsigned long long)inner_elems, (unsigned long long)outer_rows,
                         (unsigned long long)row_stride_elems);
    return IFX_OK;
}

}  // namespace ifx

using namespace ifx;

extern "C" const char* ifx_last_error(void) { return g_error; }
extern "C" int ifx_abi_version(void) { return IFX_ABI_VERSION; }
extern "C" uint64_t ifx_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void ifx_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }

extern "C" void ifx_prof_enable(int32_t on) { g_prof_on = on != 0; }
extern "C" void ifx_prof_reset(void) {
    for (auto& r : g_prof) {
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    g_prof.clear();
}
extern "C" ifx_status ifx_prof_read(const char* prefix, double* total_ms, uint64_t* launches) {
    IFX_CHECK_ARG(prefix && total_ms && launches, "ifx_prof_read: null pointer");
    const size_t n = std::strlen(prefix);
    double tot = 0.0;
    uint64_t cnt = 0;
    for (auto& r : g_prof) {
        if (r.label.compare(0, n, prefix) != 0) continue;
        IFX_CUDA_OK(cudaEventSynchronize(r.stop));
        float ms = 0.f;
        IFX_CUDA_OK(cudaEventElapsedTime(&ms, r.start, r.stop));
        tot += ms;
        ++cnt;
    }
    *total_ms = tot;
    *launches = cnt;
    return IFX_OK;
}
extern "C" ifx_status ifx_prof_labels(char* buf, int32_t cap) {
    IFX_CHECK_ARG(buf && cap > 0, "ifx_prof_labels: bad buffer");
    std::string out;
    std::vector<std::string> seen;
    for (auto& r : g_prof) {
        bool dup = false;
        for (auto& s : seen) dup = dup || (s == r.label);
        if (!dup) {
            seen.push_back(r.label);
            out += r.label;
            out += '\n';
        }
    }
    IFX_CHECK_ARG(static_cast<int32_t>(out.size()) < cap, "ifx_prof_labels: buffer too small (%zu needed)", out.size() + 1);
    std::memcpy(buf, out.c_str(), out.size() + 1);
    return IFX_OK;
}

// Shared body of the single-GPU and the sequence-parallel block.  peers == nullptr: single GPU.
static ifx_status wan_block_launches(const ifx_wan_block_weights* w, const ifx_wan_block_io* io,
                                     const ifx_peer_dst* peers, int32_t sp_mode, int32_t push_ctas, int32_t timeout_ms,
                                     const ifx_kv_plan& plan, void* stream);

// Validates, plans the append (host integers only), launches the layer.  The plan advances the cache's block table and
// end indices before anything runs on the device; if a launch is then refused (bad pointer alignment, launch failure)
// the table and indices are put back, so a caught error leaves the native cache where kv_cache_meta still says it is.
static ifx_status wan_block_forward_impl(const ifx_wan_block_weights* w, const ifx_wan_block_io* io,
                                         const ifx_peer_dst* peers, int32_t sp_mode, int32_t push_ctas,
                                         int32_t timeout_ms, ifx_kv_plan* plan_out, void* stream) {
    IFX_CHECK_ARG(w && io, "ifx_wan_block_forward: null argument");
    IFX_CHECK_ARG(io->x && io->mod && io->freqs && io->kv && io->cross_k && io->cross_v,
                  "ifx_wan_block_forward: null tensor");
    IFX_CHECK_ARG(io->ws_h && io->ws_qkv && io->ws_q && io->ws_attn && io->ws_ffn,
                  "ifx_wan_block_forward: null workspace");
    IFX_CHECK_ARG(w->dim == w->heads * w->head_dim, "ifx_wan_block_forward: dim != heads*head_dim");
    if (!w->norm3_w || !w->norm3_b)
        return set_error(IFX_ERR_UNSUPPORTED, "ifx_wan_block_forward: cross_attn_norm=False is not built");
    const int64_t S = io->rows, fs = io->tokens_per_frame;
    IFX_CHECK_ARG(S > 0 && fs > 0 && S % fs == 0, "ifx_wan_block_forward: rows must be whole frames");
    const int world = peers ? peers->world : 1;
    if (peers) {
        IFX_CHECK_ARG(world >= 2 && world <= IFX_MAX_PEERS && peers->rank >= 0 && peers->rank < world && peers->epoch > 0,
                      "ifx_wan_block_forward_sp: bad world / rank / epoch");
        IFX_CHECK_ARG(sp_mode == IFX_SP_STORE || sp_mode == IFX_SP_OVERLAP, "ifx_wan_block_forward_sp: bad mode %d", sp_mode);
        IFX_CHECK_ARG(timeout_ms > 0 && push_ctas >= 0, "ifx_wan_block_forward_sp: timeout_ms must be positive");
        IFX_CHECK_ARG(peers->flags[peers->rank] != nullptr, "ifx_wan_block_forward_sp: null flag array");
    }
    KvImpl* kv = kv_cast(io->kv);
    if (!kv) return set_error(IFX_ERR_HANDLE, "ifx_wan_block_forward: bad kv handle");
    const KvImpl before = *kv;
    ifx_kv_plan plan;
    IFX_TRY(ifx_kv_plan_append(io->kv, io->current_start, S * world, io->sink_tokens, io->windowed, &plan));
    const ifx_status st = wan_block_launches(w, io, peers, sp_mode, push_ctas, timeout_ms, plan, stream);
    if (st != IFX_OK) {
        *kv = before;
        return st;
    }
    if (plan_out) *plan_out = plan;
    return IFX_OK;
}

static ifx_status wan_block_launches(const ifx_wan_block_weights* w, const ifx_wan_block_io* io,
                                     const ifx_peer_dst* peers, int32_t sp_mode, int32_t push_ctas, int32_t timeout_ms,
                                     const ifx_kv_plan& plan, void* stream) {
    const int C = w->dim, F = w->ffn_dim;
    const int64_t S = io->rows, fs = io->tokens_per_frame;
    const int world = peers ? peers->world : 1;
    const __nv_bfloat16* mod = static_cast<const __nv_bfloat16*>(io->mod);
    const int64_t mstride = 6ll * C;  // per-frame stride of [frames, 6, C]
    const float scale = 1.0f / std::sqrt(static_cast<float>(w->head_dim));
    cudaStream_t cs = static_cast<cudaStream_t>(stream);

    // --- self-attention (causal_model.py:431-444)
    IFX_TRY(ifx_ln_modulate(io->x, io->ws_h, nullptr, nullptr, mod + 0 * C, mod + 1 * C, mstride, S, C, fs, w->eps,
                            stream));
    IFX_TRY(ifx_gemm_bf16(io->ws_h, C, w->qkv_w, C, w->qkv_b, io->ws_qkv, 3 * C, S, 3 * C, C, IFX_EPI_BIAS, nullptr, 0,
                          nullptr, 0, 0, stream));
    if (!peers) {
        IFX_TRY(ifx_qk_norm_rope_append(io->ws_qkv, 3 * C, w->norm_q_w, w->norm_k_w, io->freqs, &io->grid, io->ws_q, C,
                                        io->kv, &plan, nullptr, nullptr, S, w->heads, w->head_dim, w->eps, stream));
        IFX_TRY(ifx_attention_kv(io->ws_q, C, io->kv, io->ws_attn, C, S, scale, stream));
    } else if (sp_mode == IFX_SP_STORE) {
        // exchange in front of the attention: K / V stored into every rank's cache by the producer kernel
        ifx_peer_dst pd = *peers;
        pd.local_only = 0;
        IFX_TRY(ifx_qk_norm_rope_append_peers(io->ws_qkv, 3 * C, w->norm_q_w, w->norm_k_w, io->freqs, &io->grid,
                                              io->ws_q, C, io->kv, &plan, &pd, S, w->heads, w->head_dim, w->eps, stream));
        IFX_TRY(ifx_peer_wait(peers->flags[peers->rank], world, peers->epoch, timeout_ms, stream));
        IFX_TRY(ifx_attention_kv(io->ws_q, C, io->kv, io->ws_attn, C, S, scale, stream));
    } else {
        // exchange fused into the attention kernel, behind the attention over the cached window (see the header)
        ifx_peer_dst pd = *peers;
        pd.local_only = 1;
        IFX_TRY(ifx_qk_norm_rope_append_peers(io->ws_qkv, 3 * C, w->norm_q_w, w->norm_k_w, io->freqs, &io->grid,
                                              io->ws_q, C, io->kv, &plan, &pd, S, w->heads, w->head_dim, w->eps, stream));
        PeerPushParams push;
        IFX_TRY(fill_peer_push(push, kv_cast(io->kv), &plan, peers, static_cast<int32_t>(S / fs),
                               static_cast<int32_t>(fs)));
        push.n_ctas = push_ctas;
        IFX_TRY(attention_kv_launch(io->ws_q, C, io->kv, io->ws_attn, C, S, scale, &plan, peers->flags[peers->rank], world,
                                    peers->epoch, timeout_ms, /*pdl=*/false, cs, &push));
    }
    IFX_TRY(ifx_gemm_bf16(io->ws_attn, C, w->o_w, C, w->o_b, io->x, C, S, C, C, IFX_EPI_BIAS_GATE_RES, io->x, C,
                          mod + 2 * C, mstride, fs, stream));
    // --- cross-attention (causal_model.py:448, wan_base/model.py:66-100)
    IFX_TRY(ifx_ln_modulate(io->x, io->ws_h, w->norm3_w, w->norm3_b, nullptr, nullptr, 0, S, C, fs, w->eps, stream));
    IFX_TRY(ifx_gemm_bf16(io->ws_h, C, w->cq_w, C, w->cq_b, io->ws_qkv, C, S, C, C, IFX_EPI_BIAS, nullptr, 0, nullptr,
                          0, 0, stream));
    IFX_TRY(ifx_rmsnorm(io->ws_qkv, C, w->cnorm_q_w, io->ws_q, C, S, C, w->eps, stream));
    IFX_TRY(ifx_attention(io->ws_q, C, io->cross_k, io->cross_v, C, io->ws_attn, C, S, io->text_len, w->heads,
                          w->head_dim, scale, stream));
    IFX_TRY(ifx_gemm_bf16(io->ws_attn, C, w->co_w, C, w->co_b, io->x, C, S, C, C, IFX_EPI_BIAS_GATE_RES, io->x, C,
                          nullptr, 0, 0, stream));
    // --- FFN (causal_model.py:450-456)
    IFX_TRY(ifx_ln_modulate(io->x, io->ws_h, nullptr, nullptr, mod + 3 * C, mod + 4 * C, mstride, S, C, fs, w->eps,
                            stream));
    IFX_TRY(ifx_gemm_bf16(io->ws_h, C, w->ffn1_w, C, w->ffn1_b, io->ws_ffn, F, S, F, C, IFX_EPI_BIAS_GELU, nullptr, 0,
                          nullptr, 0, 0, stream));
    IFX_TRY(ifx_gemm_bf16(io->ws_ffn, F, w->ffn2_w, F, w->ffn2_b, io->x, C, S, C, F, IFX_EPI_BIAS_GATE_RES, io->x, C,
                          mod + 5 * C, mstride, fs, stream));
    return IFX_OK;
}

extern "C" ifx_status ifx_wan_block_forward(const ifx_wan_block_weights* w, const ifx_wan_block_io* io,
                                            ifx_kv_plan* plan_out, void* stream) {
    return wan_block_forward_impl(w, io, nullptr, 0, 0, 0, plan_out, stream);
}

extern "C" ifx_status ifx_wan_block_forward_sp(const ifx_wan_block_weights* w, const ifx_wan_block_io* io,
                                               const ifx_peer_dst* peers, int32_t mode, int32_t push_ctas,
                                               int32_t timeout_ms, ifx_kv_plan* plan_out, void* stream) {
    IFX_CHECK_ARG(peers != nullptr, "ifx_wan_block_forward_sp: peers required");
    return wan_block_forward_impl(w, io, peers, mode, push_ctas, timeout_ms, plan_out, stream);
}
