// Peer-memory plumbing for the sequence-parallel path: CUDA IPC export / open of device allocations so that one
// rank's kernels can store straight into the other ranks' replicated KV caches over NVLink / NVSwitch
// (ifx_qk_norm_rope_append_peers) instead of staging + NCCL all-gather + re-interleave.
#include <cstring>

#include "ifx_internal.h"

namespace ifx {

typedef CUresult (*AddrRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);

static AddrRangeFn addr_range_fn() {
    static AddrRangeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<AddrRangeFn>(p);
    }
    return fn;
}

}  // namespace ifx

using namespace ifx;

static_assert(sizeof(cudaIpcMemHandle_t) == IFX_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");

extern "C" ifx_status ifx_peer_export(const void* ptr, void* handle, int64_t* offset, uint64_t* base, uint64_t* size) {
    IFX_CHECK_ARG(ptr && handle && offset, "ifx_peer_export: null pointer");
    AddrRangeFn fn = addr_range_fn();
    if (!fn) return set_error(IFX_ERR_CUDA, "ifx_peer_export: cuMemGetAddressRange not available");
    CUdeviceptr b = 0;
    size_t n = 0;
    CUresult r = fn(&b, &n, reinterpret_cast<CUdeviceptr>(ptr));
    if (r != CUDA_SUCCESS) return set_error(IFX_ERR_CUDA, "ifx_peer_export: cuMemGetAddressRange failed (%d)", (int)r);
    cudaIpcMemHandle_t h;
    IFX_CUDA_OK(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(b)));
    std::memcpy(handle, &h, sizeof(h));
    *offset = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(ptr) - b);
    if (base) *base = static_cast<uint64_t>(b);
    if (size) *size = static_cast<uint64_t>(n);
    return IFX_OK;
}

extern "C" ifx_status ifx_peer_open(const void* handle, void** base) {
    IFX_CHECK_ARG(handle && base, "ifx_peer_open: null pointer");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    IFX_CUDA_OK(cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess));
    return IFX_OK;
}

extern "C" ifx_status ifx_peer_close(void* base) {
    IFX_CHECK_ARG(base != nullptr, "ifx_peer_close: null pointer");
    IFX_CUDA_OK(cudaIpcCloseMemHandle(base));
    return IFX_OK;
}
