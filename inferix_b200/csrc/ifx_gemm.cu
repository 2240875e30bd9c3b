// out[M,N] = epilogue(A[M,K] @ W[N,K]^T)   bf16 in, fp32 accumulate in TMEM, bf16 out.
//
// Persistent warp-specialised tcgen05 GEMM for sm_100a:
//   warp 0      TMA producer  (cp.async.bulk.tensor 2D, SWIZZLE_128B, kStages-deep mbarrier ring)
//   warp 1      MMA issuer    (one thread, tcgen05.mma.cta_group::1.kind::f16, 128 x 256 x 16 per instruction)
//   warp 2      TMEM allocator (512 columns = two 128x256 fp32 accumulators, double-buffered)
//   warps 4..7  epilogue      (tcgen05.ld 32x32b -> registers -> bias / GELU / gate+residual -> 16-byte stores)
// The epilogue of tile i overlaps the main loop of tile i+1 through the second TMEM accumulator.
//
// Replaces the cuBLAS nn.Linear calls + eager elementwise ops of the reference block
// (inferix/models/self_forcing/causal_model.py:171-175,333,378-379,444,455-456; wan_base/model.py:77,98).
#include <cstdlib>

#include "ifx_gemm_common.cuh"

namespace ifx {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 bf16 = 128 bytes = one swizzle atom row
constexpr int kABytes = kBM * kBK * 2;  // 16 KiB
constexpr int kGemmThreads = 256;
// Tile width is a template parameter: 256 (4 stages of 48 KiB) is the throughput shape; 128 (6 stages of 32 KiB)
// is picked by the host when the 256-wide grid would leave most SMs idle (small M under sequence parallelism).
template <int kBN>
struct GemmCfg {
    static constexpr int kStages = kBN == 256 ? 4 : 6;
    static constexpr int kBBytes = kBN * kBK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kSmem = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int kEpi, int kBN, bool kFp8>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
    constexpr int kStages = GemmCfg<kBN>::kStages;
    constexpr int kBBytes = GemmCfg<kBN>::kBBytes;
    constexpr int kStageBytes = GemmCfg<kBN>::kStageBytes;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + kStages * kABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* full = bars;                    // [kStages]
    uint64_t* empty = bars + kStages;         // [kStages]
    uint64_t* tmem_full = bars + 2 * kStages; // [2]
    uint64_t* tmem_empty = tmem_full + 2;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    constexpr int kElemsPerKb = kFp8 ? 2 * kBK : kBK;  // one k-block is 128 bytes of K per row either way
    const int num_kb = (p.K + kElemsPerKb - 1) / kElemsPerKb;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // programmatic dependent launch: everything above overlapped the previous kernel's tail; no global memory has
    // been touched yet
    griddep_launch();
    griddep_wait();

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / p.num_n_tiles;
                const int n_blk = tile % p.num_n_tiles;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], kStageBytes);
                    tma_load_2d(sA + stage * kABytes, &tmA, &full[stage], kb * kElemsPerKb, m_blk * kBM);
                    tma_load_2d(sB + stage * kBBytes, &tmB, &full[stage], kb * kElemsPerKb, n_blk * kBN);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = kFp8 ? (p.int8 ? make_idesc_s8(kBM, kBN) : make_idesc_e4m3(kBM, kBN))
                                        : make_idesc_bf16(kBM, kBN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * kBN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc_sw128(smem_u32(sA + stage * kABytes), 16, 1024);
                    const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB + stage * kBBytes), 16, 1024);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
                        if (kFp8 && p.int8)
                            umma_ss_i8(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                        else if (kFp8)
                            umma_ss_f8(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                        else
                            umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tmem_full[as]);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m_blk = tile / p.num_n_tiles;
            const int n_blk = tile % p.num_n_tiles;
            const int64_t row = static_cast<int64_t>(m_blk) * kBM + q * 32 + lane;
            const bool row_ok = row < p.M;
            const __nv_bfloat16* gate_row = nullptr;
            if (kEpi == IFX_EPI_BIAS_GATE_RES && p.gate != nullptr && row_ok)
                gate_row = p.gate + (row / p.tokens_per_frame) * p.gate_frame_stride;

            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kBN;
#pragma unroll 1
            for (int c = 0; c < kBN / 32; ++c) {
                const int col0 = n_blk * kBN + c * 32;
                if (col0 >= p.N) break;  // warp-uniform
                uint32_t acc[32];
                tmem_ld32(t_row + c * 32, acc);
                tmem_wait_ld();
                if (row_ok) gemm_epilogue_chunk<kEpi, kFp8>(p, acc, row, col0, gate_row);
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <int kEpi, int kBN, bool kFp8>
static ifx_status launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                              cudaStream_t stream) {
    constexpr int kGemmSmem = GemmCfg<kBN>::kSmem;
    static uint64_t configured = 0;     // per-device bit: the opt-in shared-memory size is a per-device attribute
    int dev = 0;
    IFX_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 64 || !(configured & (1ull << dev))) {
        IFX_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<kEpi, kBN, kFp8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kGemmSmem));
        if (dev < 64) configured |= 1ull << dev;
    }
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    const int grid = tiles < sm_count() ? tiles : sm_count();
    {
        char label[96];
        snprintf(label, sizeof(label), "gemm_%s_tn_kernel<%d,%d>[M=%lld,N=%d,K=%d]", kFp8 ? "fp8" : "bf16", kEpi, kBN,
                 (long long)p.M, p.N, p.K);
        ProfScope prof(label, stream);
        IFX_CUDA_OK(launch_kernel(gemm_bf16_tn_kernel<kEpi, kBN, kFp8>, dim3(grid), dim3(kGemmThreads), kGemmSmem, stream,
                                  true, tmA, tmB, p));
    }
    IFX_LAUNCH_OK("gemm_bf16_tn_kernel");
    return IFX_OK;
}

}  // namespace ifx

using namespace ifx;

namespace ifx {
ifx_status gemm2_dispatch(bool fp8, int bn, const void* A, int64_t lda, const void* W, int64_t ldw, GemmParams p,
                          int epilogue, cudaStream_t stream);
// IFX_GEMM_2CTA=0 forces the 1-CTA kernel everywhere (A/B testing, bisecting)
static bool use_2cta() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("IFX_GEMM_2CTA");
        v = (e == nullptr || e[0] != '0') ? 1 : 0;
    }
    return v == 1;
}
}  // namespace ifx

// dynamic quantisation extras of the 8-bit family (ifx_gemm_q8)
struct Q8Scales {
    const float* row_scale = nullptr;
    const float* col_scale = nullptr;
    int int8 = 0;
};

template <bool kFp8>
static ifx_status gemm_entry(const void* A, int64_t lda, const void* W, int64_t ldw, float alpha, const void* bias,
                             void* out, int64_t ldo, int64_t M, int32_t N, int32_t K, int32_t epilogue,
                             const void* residual, int64_t ldr, const void* gate, int64_t gate_frame_stride,
                             int64_t tokens_per_frame, void* stream, Q8Scales q8 = Q8Scales()) {
    constexpr int kAlign = kFp8 ? 16 : 8;  // operand rows must be 16-byte multiples
    IFX_CHECK_ARG(A && W && out, "ifx_gemm_bf16: null pointer");
    IFX_CHECK_ARG(M > 0 && N > 0 && K > 0, "ifx_gemm_bf16: empty problem M=%lld N=%d K=%d", (long long)M, N, K);
    IFX_CHECK_ARG(N % 8 == 0 && K % kAlign == 0, "ifx_gemm: N %% 8 and K %% %d must be 0 (N=%d K=%d)", kAlign, N, K);
    IFX_CHECK_ARG(lda % kAlign == 0 && ldw % kAlign == 0 && ldo % 8 == 0, "ifx_gemm: strides must be 16-byte multiples");
    IFX_CHECK_ARG(lda >= K && ldw >= K && ldo >= N, "ifx_gemm_bf16: stride smaller than row");
    IFX_CHECK_ARG(epilogue >= IFX_EPI_BIAS && epilogue <= IFX_EPI_BIAS_F32, "ifx_gemm_bf16: bad epilogue %d",
                  epilogue);
    if (epilogue == IFX_EPI_BIAS_GATE_RES) {
        IFX_CHECK_ARG(residual != nullptr && ldr >= N && ldr % 8 == 0, "ifx_gemm_bf16: residual required");
        IFX_CHECK_ARG(gate == nullptr || tokens_per_frame > 0, "ifx_gemm_bf16: tokens_per_frame must be > 0");
    }
    auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    IFX_CHECK_ARG(aligned16(A) && aligned16(W) && aligned16(out) && aligned16(bias) && aligned16(residual) &&
                      aligned16(gate),
                  "ifx_gemm_bf16: pointers must be 16-byte aligned");

    // tile width: wave efficiency of the 256-wide grid vs the 128-wide one (which is ~15 % less efficient per tile
    // because A and B shared-memory reads per MMA are no longer amortised over 256 columns)
    const int sms = sm_count();
    const int64_t mt = (M + kBM - 1) / kBM;
    auto wave_eff = [&](int bn) {
        const int64_t tiles = mt * ((N + bn - 1) / bn);
        return static_cast<double>(tiles) / static_cast<double>(((tiles + sms - 1) / sms) * sms);
    };
    int bn = (0.85 * wave_eff(128) > wave_eff(256)) ? 128 : 256;
    // 2-CTA kernel (ifx_gemm2.cu): 256-row cluster tiles, 256 or 128 columns wide, picked by the same wave argument
    const bool pair = M >= 256 && use_2cta();
    if (pair) {
        const int clusters = sms / 2;
        const int64_t mt2 = (M + 255) / 256;
        auto eff2 = [&](int w) {
            const int64_t tiles = mt2 * ((N + w - 1) / w);
            return static_cast<double>(tiles) / static_cast<double>(((tiles + clusters - 1) / clusters) * clusters);
        };
        bn = (0.7 * eff2(128) > eff2(256)) ? 128 : 256;   // measured: a 128-wide cluster tile runs at ~0.7x the rate
    }

    CUtensorMap tmA, tmB;
    ifx_status st = kFp8 ? make_tmap_u8_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 2 * kBK, kBM)
                         : make_tmap_bf16_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, kBK, kBM);
    if (st != IFX_OK) return st;
    st = kFp8 ? make_tmap_u8_2d(&tmB, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, 2 * kBK, bn)
              : make_tmap_bf16_2d(&tmB, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, kBK, bn);
    if (st != IFX_OK) return st;

    GemmParams p;
    p.M = M;
    p.N = N;
    p.K = K;
    p.alpha = alpha;
    p.row_scale = q8.row_scale;
    p.col_scale = q8.col_scale;
    p.int8 = q8.int8;
    p.bias = static_cast<const __nv_bfloat16*>(bias);
    p.out = static_cast<__nv_bfloat16*>(out);
    p.ldo = ldo;
    p.residual = static_cast<const __nv_bfloat16*>(residual);
    p.ldr = ldr;
    p.gate = static_cast<const __nv_bfloat16*>(gate);
    p.gate_frame_stride = gate_frame_stride;
    p.tokens_per_frame = tokens_per_frame > 0 ? tokens_per_frame : 1;
    p.num_m_tiles = static_cast<int32_t>((M + kBM - 1) / kBM);
    p.num_n_tiles = (N + bn - 1) / bn;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (pair) return gemm2_dispatch(kFp8, bn, A, lda, W, ldw, p, epilogue, s);
    if (bn == 256) {
        switch (epilogue) {
            case IFX_EPI_BIAS: return launch_gemm<IFX_EPI_BIAS, 256, kFp8>(tmA, tmB, p, s);
            case IFX_EPI_BIAS_GELU: return launch_gemm<IFX_EPI_BIAS_GELU, 256, kFp8>(tmA, tmB, p, s);
            case IFX_EPI_BIAS_GELU_ERF: return launch_gemm<IFX_EPI_BIAS_GELU_ERF, 256, kFp8>(tmA, tmB, p, s);
            case IFX_EPI_BIAS_F32: return launch_gemm<IFX_EPI_BIAS_F32, 256, kFp8>(tmA, tmB, p, s);
            default: return launch_gemm<IFX_EPI_BIAS_GATE_RES, 256, kFp8>(tmA, tmB, p, s);
        }
    }
    switch (epilogue) {
        case IFX_EPI_BIAS: return launch_gemm<IFX_EPI_BIAS, 128, kFp8>(tmA, tmB, p, s);
        case IFX_EPI_BIAS_GELU: return launch_gemm<IFX_EPI_BIAS_GELU, 128, kFp8>(tmA, tmB, p, s);
        case IFX_EPI_BIAS_GELU_ERF: return launch_gemm<IFX_EPI_BIAS_GELU_ERF, 128, kFp8>(tmA, tmB, p, s);
        case IFX_EPI_BIAS_F32: return launch_gemm<IFX_EPI_BIAS_F32, 128, kFp8>(tmA, tmB, p, s);
        default: return launch_gemm<IFX_EPI_BIAS_GATE_RES, 128, kFp8>(tmA, tmB, p, s);
    }
}

extern "C" ifx_status ifx_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias,
                                    void* out, int64_t ldo, int64_t M, int32_t N, int32_t K, int32_t epilogue,
                                    const void* residual, int64_t ldr, const void* gate, int64_t gate_frame_stride,
                                    int64_t tokens_per_frame, void* stream) {
    return gemm_entry<false>(A, lda, W, ldw, 1.0f, bias, out, ldo, M, N, K, epilogue, residual, ldr, gate,
                             gate_frame_stride, tokens_per_frame, stream);
}

extern "C" ifx_status ifx_gemm_fp8(const void* A, int64_t lda, const void* W, int64_t ldw, float alpha,
                                   const void* bias, void* out, int64_t ldo, int64_t M, int32_t N, int32_t K,
                                   int32_t epilogue, const void* residual, int64_t ldr, const void* gate,
                                   int64_t gate_frame_stride, int64_t tokens_per_frame, void* stream) {
    IFX_CHECK_ARG(alpha > 0.f, "ifx_gemm_fp8: alpha (input_scale * weight_scale) must be positive");
    return gemm_entry<true>(A, lda, W, ldw, alpha, bias, out, ldo, M, N, K, epilogue, residual, ldr, gate,
                            gate_frame_stride, tokens_per_frame, stream);
}

extern "C" ifx_status ifx_gemm_q8(const void* A, int64_t lda, const void* W, int64_t ldw, const float* row_scale,
                                  const float* col_scale, int32_t kind, const void* bias, void* out, int64_t ldo,
                                  int64_t M, int32_t N, int32_t K, int32_t epilogue, const void* residual, int64_t ldr,
                                  const void* gate, int64_t gate_frame_stride, int64_t tokens_per_frame, void* stream) {
    IFX_CHECK_ARG(row_scale != nullptr && col_scale != nullptr, "ifx_gemm_q8: row_scale and col_scale are required");
    IFX_CHECK_ARG(kind == IFX_Q8_E4M3 || kind == IFX_Q8_INT8, "ifx_gemm_q8: kind must be IFX_Q8_E4M3 or IFX_Q8_INT8");
    Q8Scales q8;
    q8.row_scale = row_scale;
    q8.col_scale = col_scale;
    q8.int8 = kind == IFX_Q8_INT8 ? 1 : 0;
    return gemm_entry<true>(A, lda, W, ldw, 1.0f, bias, out, ldo, M, N, K, epilogue, residual, ldr, gate,
                            gate_frame_stride, tokens_per_frame, stream, q8);
}
