// Host side of the paged KV cache: block table, free list, and the reference's end-index arithmetic.
//
// The reference keeps one contiguous tensor per layer and, once the window is full, copies the surviving
// tokens left on every new block (inferix/models/self_forcing/causal_model.py:282-296).  Here the same logical
// sequence is a list of frame-sized pages; eviction unlinks pages from the list and re-links them at the tail.
// Index results (local_start, local_end, global_end, num_evicted) are bit-identical to the reference's.
#include <algorithm>
#include <new>

#include "ifx_internal.h"
#include "ifx_ptx.cuh"

namespace ifx {

KvImpl* kv_cast(ifx_kv* kv) {
    KvImpl* k = reinterpret_cast<KvImpl*>(kv);
    return (k && k->magic == kKvMagic) ? k : nullptr;
}
const KvImpl* kv_cast(const ifx_kv* kv) {
    const KvImpl* k = reinterpret_cast<const KvImpl*>(kv);
    return (k && k->magic == kKvMagic) ? k : nullptr;
}

static int32_t take_page(KvImpl* kv) {
    if (!kv->free_pages.empty()) {
        int32_t p = kv->free_pages.front();
        kv->free_pages.erase(kv->free_pages.begin());
        return p;
    }
    if (kv->next_fresh < kv->num_pages) return kv->next_fresh++;
    return -1;
}

}  // namespace ifx

using namespace ifx;

extern "C" ifx_status ifx_kv_create(ifx_kv** out, void* k_base, void* v_base, int32_t num_pages,
                                    int32_t page_tokens, int32_t heads, int32_t head_dim) {
    IFX_CHECK_ARG(out && k_base && v_base, "ifx_kv_create: null pointer");
    IFX_CHECK_ARG(num_pages > 0 && page_tokens > 0 && heads > 0 && head_dim > 0, "ifx_kv_create: bad geometry");
    IFX_CHECK_ARG((heads * head_dim) % 8 == 0, "ifx_kv_create: heads*head_dim must be a multiple of 8");
    IFX_CHECK_ARG((reinterpret_cast<uintptr_t>(k_base) & 15) == 0 && (reinterpret_cast<uintptr_t>(v_base) & 15) == 0,
                  "ifx_kv_create: buffers must be 16-byte aligned");
    KvImpl* kv = new (std::nothrow) KvImpl();
    if (!kv) return set_error(IFX_ERR_OOM, "ifx_kv_create: out of host memory");
    kv->magic = kKvMagic;
    kv->k_base = k_base;
    kv->v_base = v_base;
    kv->num_pages = num_pages;
    kv->page_tokens = page_tokens;
    kv->heads = heads;
    kv->head_dim = head_dim;
    kv->global_end = 0;
    kv->local_end = 0;
    kv->next_fresh = 0;
    kv->rotated = false;
    *out = reinterpret_cast<ifx_kv*>(kv);
    return IFX_OK;
}

extern "C" ifx_status ifx_kv_rebind(ifx_kv* kv_, void* k_base, void* v_base) {
    KvImpl* kv = kv_cast(kv_);
    if (!kv) return set_error(IFX_ERR_HANDLE, "ifx_kv_rebind: bad kv handle");
    IFX_CHECK_ARG(k_base && v_base, "ifx_kv_rebind: null pointer");
    IFX_CHECK_ARG((reinterpret_cast<uintptr_t>(k_base) & 15) == 0 && (reinterpret_cast<uintptr_t>(v_base) & 15) == 0,
                  "ifx_kv_rebind: buffers must be 16-byte aligned");
    kv->k_base = k_base;
    kv->v_base = v_base;
    return IFX_OK;
}

extern "C" ifx_status ifx_kv_destroy(ifx_kv* kv_) {
    KvImpl* kv = kv_cast(kv_);
    if (!kv) return set_error(IFX_ERR_HANDLE, "ifx_kv_destroy: bad kv handle");
    kv->magic = 0;
    delete kv;
    return IFX_OK;
}

extern "C" ifx_status ifx_kv_reset(ifx_kv* kv_) {
    KvImpl* kv = kv_cast(kv_);
    if (!kv) return set_error(IFX_ERR_HANDLE, "ifx_kv_reset: bad kv handle");
    kv->global_end = 0;
    kv->local_end = 0;
    kv->table.clear();
    kv->free_pages.clear();
    kv->next_fresh = 0;
    kv->rotated = false;
    return IFX_OK;
}

extern "C" ifx_status ifx_kv_plan_append(ifx_kv* kv_, int64_t current_start, int64_t num_new, int64_t sink_tokens,
                                         int32_t windowed, ifx_kv_plan* plan) {
    KvImpl* kv = kv_cast(kv_);
    if (!kv) return set_error(IFX_ERR_HANDLE, "ifx_kv_plan_append: bad kv handle");
    IFX_CHECK_ARG(plan != nullptr, "ifx_kv_plan_append: null plan");
    IFX_CHECK_ARG(current_start >= 0 && num_new > 0 && sink_tokens >= 0, "ifx_kv_plan_append: negative argument");
    const int64_t pt = kv->page_tokens;
    if (current_start % pt || num_new % pt || sink_tokens % pt)
        return set_error(IFX_ERR_UNSUPPORTED,
                         "ifx_kv_plan_append: current_start=%lld num_new=%lld sink=%lld must be multiples of the "
                         "page (frame) size %lld",
                         (long long)current_start, (long long)num_new, (long long)sink_tokens, (long long)pt);
    IFX_CHECK_ARG(num_new / pt <= IFX_KV_MAX_PLAN_PAGES, "ifx_kv_plan_append: more than %d frames per block",
                  IFX_KV_MAX_PLAN_PAGES);

    // ---- causal_model.py:277-300, on host integers
    const int64_t cache_size = static_cast<int64_t>(kv->num_pages) * pt;
    const int64_t current_end = current_start + num_new;
    int64_t evicted = 0;
    int64_t local_end_new;
    if (windowed && current_end > kv->global_end && num_new + kv->local_end > cache_size) {
        evicted = num_new + kv->local_end - cache_size;
        const int64_t rolled = kv->local_end - evicted - sink_tokens;
        if (rolled < 0)
            return set_error(IFX_ERR_BOUNDS, "ifx_kv_plan_append: window too small (evict %lld, sink %lld, have %lld)",
                             (long long)evicted, (long long)sink_tokens, (long long)kv->local_end);
        local_end_new = kv->local_end + current_end - kv->global_end - evicted;
    } else {
        local_end_new = kv->local_end + current_end - kv->global_end;
    }
    const int64_t local_start = local_end_new - num_new;
    if (local_start < 0 || local_end_new > cache_size)
        return set_error(IFX_ERR_BOUNDS, "ifx_kv_plan_append: tokens [%lld, %lld) fall outside the cache of %lld",
                         (long long)local_start, (long long)local_end_new, (long long)cache_size);

    // ---- table rotation instead of the byte roll of :289-292
    std::vector<int32_t> table = kv->table;
    std::vector<int32_t> free_pages = kv->free_pages;
    int32_t next_fresh = kv->next_fresh;
    if (evicted > 0) {
        const size_t s = static_cast<size_t>(sink_tokens / pt), e = static_cast<size_t>(evicted / pt);
        for (size_t i = 0; i < e; ++i) free_pages.push_back(table[s + i]);
        table.erase(table.begin() + s, table.begin() + s + e);
    }
    const size_t need_pages = static_cast<size_t>(local_end_new / pt);
    if (table.size() > need_pages) {
        // the write position moved backwards (re-generation): drop the tail
        for (size_t i = need_pages; i < table.size(); ++i) free_pages.push_back(table[i]);
        table.resize(need_pages);
    }
    while (table.size() < need_pages) {
        int32_t pg;
        if (!free_pages.empty()) {
            pg = free_pages.front();
            free_pages.erase(free_pages.begin());
        } else if (next_fresh < kv->num_pages) {
            pg = next_fresh++;
        } else {
            return set_error(IFX_ERR_BOUNDS, "ifx_kv_plan_append: out of pages");
        }
        table.push_back(pg);
    }
    // commit
    if (evicted > 0 || kv->table.size() > need_pages) kv->rotated = true;
    kv->table.swap(table);
    kv->free_pages.swap(free_pages);
    kv->next_fresh = next_fresh;
    kv->global_end = current_end;
    kv->local_end = local_end_new;

    plan->local_start = local_start;
    plan->local_end = local_end_new;
    plan->global_end = current_end;
    plan->num_evicted = evicted;
    plan->num_pages = static_cast<int32_t>(num_new / pt);
    plan->first_offset = 0;
    for (int i = 0; i < plan->num_pages; ++i) plan->pages[i] = kv->table[static_cast<size_t>(local_start / pt) + i];
    for (int i = plan->num_pages; i < IFX_KV_MAX_PLAN_PAGES; ++i) plan->pages[i] = -1;
    return IFX_OK;
}

extern "C" ifx_status ifx_kv_state(const ifx_kv* kv_, int64_t* global_end, int64_t* local_end, int32_t* valid_pages,
                                   int32_t* table_out, int32_t table_cap) {
    const KvImpl* kv = kv_cast(kv_);
    if (!kv) return set_error(IFX_ERR_HANDLE, "ifx_kv_state: bad kv handle");
    if (global_end) *global_end = kv->global_end;
    if (local_end) *local_end = kv->local_end;
    if (valid_pages) *valid_pages = static_cast<int32_t>(kv->table.size());
    if (table_out) {
        IFX_CHECK_ARG(table_cap >= static_cast<int32_t>(kv->table.size()), "ifx_kv_state: table_cap too small");
        std::copy(kv->table.begin(), kv->table.end(), table_out);
    }
    return IFX_OK;
}

extern "C" ifx_status ifx_kv_map(ifx_kv* kv_, int64_t tokens, void** k_rows, void** v_rows) {
    KvImpl* kv = kv_cast(kv_);
    if (!kv) return set_error(IFX_ERR_HANDLE, "ifx_kv_map: bad kv handle");
    IFX_CHECK_ARG(tokens >= 0, "ifx_kv_map: negative token count");
    const int64_t pt = kv->page_tokens;
    if (tokens > static_cast<int64_t>(kv->num_pages) * pt)
        return set_error(IFX_ERR_BOUNDS, "ifx_kv_map: %lld tokens beyond the cache of %lld", (long long)tokens,
                         (long long)(kv->num_pages * pt));
    if (kv->rotated)
        return set_error(IFX_ERR_UNSUPPORTED, "ifx_kv_map: the cache has been rotated by ifx_kv_plan_append; its "
                                              "logical order is no longer the physical order");
    const size_t need_pages = static_cast<size_t>((tokens + pt - 1) / pt);
    while (kv->table.size() < need_pages) {
        const int32_t pg = take_page(kv);
        if (pg != static_cast<int32_t>(kv->table.size())) {
            if (pg >= 0) kv->free_pages.insert(kv->free_pages.begin(), pg);
            return set_error(IFX_ERR_UNSUPPORTED, "ifx_kv_map: page allocator is not in identity order");
        }
        kv->table.push_back(pg);
    }
    if (k_rows) *k_rows = kv->k_base;
    if (v_rows) *v_rows = kv->v_base;
    return IFX_OK;
}

static ifx_status paged_io(const KvImpl* kv, void* lin_k, void* lin_v, int64_t start, int64_t length, int mode,
                           cudaStream_t stream) {
    const int64_t pt = kv->page_tokens;
    const int C = kv->heads * kv->head_dim;
    int64_t done = 0;
    const int64_t last_page = (start + length + pt - 1) / pt;  // exclusive
    while (done < length) {
        const int64_t lt = start + done;
        const int64_t lp0 = lt / pt;
        PagedCopyParams p;
        p.cache_k = static_cast<__nv_bfloat16*>(kv->k_base);
        p.cache_v = static_cast<__nv_bfloat16*>(kv->v_base);
        p.lin_k = lin_k ? static_cast<__nv_bfloat16*>(lin_k) + done * C : nullptr;
        p.lin_v = lin_v ? static_cast<__nv_bfloat16*>(lin_v) + done * C : nullptr;
        p.ld_lin = C;
        p.first_logical = lt;
        p.page_tokens = kv->page_tokens;
        p.C = C;
        p.mode = mode;
        p.linear_row0 = 0;
        // run of physically consecutive pages -> one dense extent (always the case for never-evicted caches)
        int64_t run = 1;
        while (lp0 + run < last_page &&
               kv->table[static_cast<size_t>(lp0 + run)] == kv->table[static_cast<size_t>(lp0)] + run)
            ++run;
        int64_t chunk_end_page;
        if (run >= 2 || last_page - lp0 == 1) {
            chunk_end_page = lp0 + run;
            p.pl.n = 0;
            p.linear_row0 = static_cast<int64_t>(kv->table[static_cast<size_t>(lp0)]) * pt + lt % pt;
        } else {
            chunk_end_page = std::min<int64_t>(lp0 + IFX_KV_MAX_PLAN_PAGES, last_page);
            p.pl.n = static_cast<int32_t>(chunk_end_page - lp0);
            for (int i = 0; i < p.pl.n; ++i) p.pl.pages[i] = kv->table[static_cast<size_t>(lp0 + i)];
        }
        p.rows = std::min<int64_t>(chunk_end_page * pt, start + length) - lt;
        ifx_status st = launch_paged_copy(p, stream);
        if (st != IFX_OK) return st;
        done += p.rows;
    }
    return IFX_OK;
}

extern "C" ifx_status ifx_kv_export(const ifx_kv* kv_, void* dst_k, void* dst_v, int64_t start, int64_t length,
                                    void* stream) {
    const KvImpl* kv = kv_cast(kv_);
    if (!kv) return set_error(IFX_ERR_HANDLE, "ifx_kv_export: bad kv handle");
    IFX_CHECK_ARG(dst_k || dst_v, "ifx_kv_export: nothing to export");
    IFX_CHECK_ARG(start >= 0 && length >= 0, "ifx_kv_export: negative range");
    if (length == 0) return IFX_OK;
    if (start + length > static_cast<int64_t>(kv->table.size()) * kv->page_tokens)
        return set_error(IFX_ERR_BOUNDS, "ifx_kv_export: [%lld, %lld) beyond the %lld mapped tokens",
                         (long long)start, (long long)(start + length),
                         (long long)(kv->table.size() * kv->page_tokens));
    return paged_io(kv, dst_k, dst_v, start, length, 1, static_cast<cudaStream_t>(stream));
}

extern "C" ifx_status ifx_kv_import(ifx_kv* kv_, const void* src_k, const void* src_v, int64_t start, int64_t length,
                                    void* stream) {
    KvImpl* kv = kv_cast(kv_);
    if (!kv) return set_error(IFX_ERR_HANDLE, "ifx_kv_import: bad kv handle");
    IFX_CHECK_ARG(src_k || src_v, "ifx_kv_import: nothing to import");
    IFX_CHECK_ARG(start >= 0 && length >= 0, "ifx_kv_import: negative range");
    if (length == 0) return IFX_OK;
    const int64_t pt = kv->page_tokens;
    if (start + length > static_cast<int64_t>(kv->num_pages) * pt)
        return set_error(IFX_ERR_BOUNDS, "ifx_kv_import: [%lld, %lld) beyond the cache of %lld tokens",
                         (long long)start, (long long)(start + length), (long long)(kv->num_pages * pt));
    const size_t need_pages = static_cast<size_t>((start + length + pt - 1) / pt);
    while (kv->table.size() < need_pages) {
        const int32_t pg = take_page(kv);
        if (pg < 0) return set_error(IFX_ERR_BOUNDS, "ifx_kv_import: out of pages");
        kv->table.push_back(pg);
    }
    return paged_io(kv, const_cast<void*>(src_k), const_cast<void*>(src_v), start, length, 0,
                    static_cast<cudaStream_t>(stream));
}
