// Internal host-side helpers shared by the .cu translation units (not part of the ABI).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <vector>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/inferix_b200.h"

namespace ifx {

// thread-local error text; set_error returns `code` so call sites can `return set_error(...)`.
ifx_status set_error(ifx_status code, const char* fmt, ...);
void count_launch(int n = 1);

#define IFX_CHECK_ARG(cond, ...)                                  \
    do {                                                          \
        if (!(cond)) return ::ifx::set_error(IFX_ERR_INVALID, __VA_ARGS__); \
    } while (0)

#define IFX_CUDA_OK(expr)                                                                             \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return ::ifx::set_error(IFX_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));    \
    } while (0)

#define IFX_TRY(expr)                     \
    do {                                  \
        ifx_status _s = (expr);           \
        if (_s != IFX_OK) return _s;      \
    } while (0)

// After a kernel launch: surfaces launch-configuration errors without synchronising.
#define IFX_LAUNCH_OK(name)                                                                          \
    do {                                                                                             \
        cudaError_t _e = cudaGetLastError();                                                         \
        if (_e != cudaSuccess)                                                                       \
            return ::ifx::set_error(IFX_ERR_CUDA, "launch %s failed: %s", name, cudaGetErrorString(_e)); \
        ::ifx::count_launch();                                                                       \
    } while (0)

// RAII bracket around a kernel launch; records two events on `stream` when profiling is enabled.
struct ProfScope {
    ProfScope(const char* label, cudaStream_t stream);
    ~ProfScope();
    int slot;
    cudaStream_t stream;
};

// 2-D bf16 tensor map, 128-byte swizzle, box = [box_rows, 64 elements].
ifx_status make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner_elems, uint64_t outer_rows,
                             uint64_t row_stride_elems, uint32_t box_inner, uint32_t box_rows);

// same for 1-byte elements (FP8 operands): box = [box_rows, 128 bytes]
ifx_status make_tmap_u8_2d(CUtensorMap* out, const void* base, uint64_t inner_elems, uint64_t outer_rows,
                           uint64_t row_stride_elems, uint32_t box_inner, uint32_t box_rows);

int sm_count();
ifx_status device_counter(unsigned int* (&slots)[64], unsigned int** out);

// Programmatic dependent launch (PDL).  Kernels of the DiT layer are launched with the programmatic-stream-
// serialization attribute: their CTAs may become resident while the previous kernel on the stream is still draining,
// run their prologue (barrier init, TMEM allocation, descriptor prefetch) and block in griddepcontrol.wait — which
// every such kernel executes before its first global-memory access — until the previous kernel has completed and
// flushed.  IFX_PDL=0 turns the attribute off (plain stream order; griddepcontrol.wait is then a no-op).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct KvImpl {
    uint32_t magic;
    void* k_base;
    void* v_base;
    int32_t num_pages;
    int32_t page_tokens;
    int32_t heads;
    int32_t head_dim;
    int64_t global_end;
    int64_t local_end;
    std::vector<int32_t> table;      // logical page -> physical page, size = ceil(local_end / page_tokens)
    std::vector<int32_t> free_pages; // recycled pages, reused LIFO-last (FIFO order) before fresh ones
    int32_t next_fresh;              // physical pages [0, next_fresh) have been handed out at least once
    bool rotated;                    // a plan_append has unlinked / relinked pages: logical order != physical order
};
// by-value page list handed to kernels (no device-side table, no host sync)
struct PageList {
    int32_t n;
    int32_t pages[IFX_KV_MAX_PLAN_PAGES];
};
// Exchange of the block's new K / V rows over peer memory: rank `rank` owns hw indices [rank*chunk, (rank+1)*chunk) of
// each of the `frames` new frames; its rows are copied from its own cache to the same rows of every other rank's
// cache, then `epoch` is published in every rank's flag array.  Run either as its own grid (peer_push_kernel) or as a
// side job of the attention kernel (one otherwise idle warp of each of the first n_ctas CTAs).
struct PeerPushParams {
    int32_t world, rank, frames, chunk, page_tokens, C;
    int32_t n_ctas;          // CTAs sharing the copy (in-attention mode); 0 = disabled
    PageList pl;
    __nv_bfloat16* peer_k[IFX_MAX_PEERS];
    __nv_bfloat16* peer_v[IFX_MAX_PEERS];
    long long* peer_flags[IFX_MAX_PEERS];
    long long epoch;
    unsigned int* done_counter;
};
// fills PeerPushParams from the ABI structs (validates geometry); done_counter is left to the caller
ifx_status fill_peer_push(PeerPushParams& p, const KvImpl* kv, const ifx_kv_plan* plan, const ifx_peer_dst* peers,
                          int32_t frames, int32_t chunk);

// mode 0: contiguous rows -> cache pages (append / import); mode 1: cache pages -> contiguous rows (export)
struct PagedCopyParams {
    __nv_bfloat16* cache_k;
    __nv_bfloat16* cache_v;
    __nv_bfloat16* lin_k;  // contiguous side (either may be null)
    __nv_bfloat16* lin_v;
    int64_t ld_lin;
    int64_t rows;           // rows in this launch
    int64_t first_logical;  // logical token index of row 0
    int32_t page_tokens;
    int32_t C;
    int32_t mode;
    PageList pl;            // pages[i] backs logical page (first_logical / page_tokens + i)
    int64_t linear_row0;    // pl.n == 0: the rows are physically contiguous starting at this cache row
};
ifx_status launch_paged_copy(const PagedCopyParams& p, cudaStream_t stream);

// attention over a paged cache; fresh != nullptr: the plan's pages sit behind the flag wait (ifx_attention_kv_wait);
// pdl: the kernel may run NEXT TO its predecessor on the stream (only for a predecessor that releases it explicitly);
// push != nullptr: the kernel also ships this rank's rows of the fresh pages to the peers (fused exchange)
ifx_status attention_kv_launch(const void* q, int64_t ldq, const ifx_kv* kv, void* out, int64_t ldo, int64_t q_rows,
                               float softmax_scale, const ifx_kv_plan* fresh, const int64_t* flags, int32_t world,
                               int64_t epoch, int32_t timeout_ms, bool pdl, cudaStream_t stream,
                               const PeerPushParams* push = nullptr);

constexpr uint32_t kKvMagic = 0x4B564958u;  // "XIVK"
KvImpl* kv_cast(ifx_kv* kv);
const KvImpl* kv_cast(const ifx_kv* kv);

}  // namespace ifx
