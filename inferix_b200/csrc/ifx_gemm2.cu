// 2-CTA (cta_group::2) variant of the tcgen05 GEMM: a cluster of two CTAs on one TPC owns a 256 x 256 output tile.
//
// Why: with one CTA per tile every 128x256x16 MMA reads 4 KiB of A and 8 KiB of B from shared memory while TMA writes
// the next stage into it — more than the SM's shared-memory port sustains, so the 1-CTA kernel tops out near half the
// tensor peak.  In pair mode each CTA stages its own 128 rows of A and HALF of B (128 of the 256 weight rows); the
// hardware feeds both halves to both tensor cores, so per-CTA operand traffic drops by a third and a stage shrinks
// from 48 to 32 KiB (6 stages instead of 4).
//
//   both CTAs : warp 0 TMA producer (cp.async.bulk.tensor ... cta_group::2, completion bytes land on the LEADER's
//               mbarrier), warp 2 TMEM allocator (cta_group::2), warps 4..11 epilogue of the CTA's own 128 rows
//               (two warps per TMEM lane quadrant, each taking half of the 256 columns)
//   leader    : warp 1 issues tcgen05.mma.cta_group::2 (M = 256) and commits with a cluster-multicast arrive, which
//               releases the smem stage in both CTAs and publishes the accumulator to both epilogues
//   peer      : its epilogue threads release the accumulator on the leader's tmem_empty barrier (mapa + remote arrive)
#include "ifx_gemm_common.cuh"

namespace ifx {

constexpr int k2BM = 128;        // rows per CTA (256 per cluster)
constexpr int k2BK = 64;
constexpr int k2ABytes = k2BM * k2BK * 2;      // 16 KiB
constexpr int k2Threads = 384;      // 4 control warps + 8 epilogue warps (two per SM sub-partition)
// Cluster tile width: 256 (throughput shape) or 128 (small-M grids: twice the tiles, e.g. the 1350-row shard of an
// 8-way sequence-parallel run).  Each CTA stages half of the weight rows.
template <int kBN>
struct Gemm2Cfg {
    static constexpr int kBNHalf = kBN / 2;
    static constexpr int kBBytes = kBNHalf * k2BK * 2;
    static constexpr int kStageBytes = k2ABytes + kBBytes;
    static constexpr int kStages = kBN == 256 ? 6 : 8;
    static constexpr int kSmem = kStages * kStageBytes + 1024 + 256;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load executed by either CTA of the pair; the completion bytes are credited to the leader CTA's mbarrier
// (same smem offset, CTA-rank bit cleared — CUTLASS SM100_TMA_2SM_LOAD_2D).
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                                                 int32_t c1) {
    const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%3, %4}], [%2], %5;"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar), "r"(c0), "r"(c1), "l"(kEvictNormal)
        : "memory");
}
__device__ __forceinline__ void umma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate, bool fp8, bool int8 = false) {
    if (int8) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
            :
            : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else if (fp8) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
            :
            : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
            :
            : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// arrive (once all MMAs issued so far by this thread are done) on the barrier at this smem offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 0x3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :
                 : "r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n"
        :
        : "r"(smem_u32(bar)), "r"(cta)
        : "memory");
}

template <int kEpi, bool kFp8, int k2BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1)
gemm2_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    constexpr int k2BNHalf = Gemm2Cfg<k2BN>::kBNHalf;
    constexpr int k2BBytes = Gemm2Cfg<k2BN>::kBBytes;
    constexpr int k2StageBytes = Gemm2Cfg<k2BN>::kStageBytes;
    constexpr int k2Stages = Gemm2Cfg<k2BN>::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + k2Stages * k2ABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + k2Stages * k2StageBytes);
    uint64_t* full = bars;                      // [k2Stages]  (leader's copy is the live one)
    uint64_t* empty = bars + k2Stages;          // [k2Stages]  (one per CTA, multicast-released)
    uint64_t* tmem_full = bars + 2 * k2Stages;  // [2]         (one per CTA, multicast-published)
    uint64_t* tmem_empty = tmem_full + 2;       // [2]         (leader's copy: 256 arrivals = both epilogues)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;  // tiles of 256 x 256
    constexpr int kElemsPerKb = kFp8 ? 2 * k2BK : k2BK;
    const int num_kb = (p.K + kElemsPerKb - 1) / kElemsPerKb;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < k2Stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 512);   // 8 epilogue warps x 32 lanes x 2 CTAs
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // programmatic dependent launch: the prologue above overlapped the previous kernel's tail; nothing has touched
    // global memory yet
    griddep_launch();
    griddep_wait();

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m_blk = tile / p.num_n_tiles;
                const int n_blk = tile % p.num_n_tiles;
                const int m0 = m_blk * (2 * k2BM) + static_cast<int>(rank) * k2BM;
                const int n0 = n_blk * k2BN + static_cast<int>(rank) * k2BNHalf;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (leader) mbar_expect_tx(&full[stage], 2 * k2StageBytes);   // both CTAs' bytes
                    tma_load_2d_pair(sA + stage * k2ABytes, &tmA, &full[stage], kb * kElemsPerKb, m0);
                    tma_load_2d_pair(sB + stage * k2BBytes, &tmB, &full[stage], kb * kElemsPerKb, n0);
                    if (++stage == k2Stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            const uint32_t idesc = kFp8 ? (p.int8 ? make_idesc_s8(2 * k2BM, k2BN) : make_idesc_e4m3(2 * k2BM, k2BN))
                                        : make_idesc_bf16(2 * k2BM, k2BN, 0, 0);
            const bool int8 = kFp8 && p.int8;
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * k2BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc_sw128(smem_u32(sA + stage * k2ABytes), 16, 1024);
                    const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB + stage * k2BBytes), 16, 1024);
#pragma unroll
                    for (int k = 0; k < k2BK / 16; ++k)
                        umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0, kFp8, int8);
                    umma_commit_pair(&empty[stage]);
                    if (++stage == k2Stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit_pair(&tmem_full[as]);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;              // TMEM lane quadrant
        const int chalf = (warp - 4) >> 2;   // warps 4..7 take the first half of the columns, warps 8..11 the second
        int it = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m_blk = tile / p.num_n_tiles;
            const int n_blk = tile % p.num_n_tiles;
            const int64_t row = static_cast<int64_t>(m_blk) * (2 * k2BM) + rank * k2BM + q * 32 + lane;
            const bool row_ok = row < p.M;
            const __nv_bfloat16* gate_row = nullptr;
            if (kEpi == IFX_EPI_BIAS_GATE_RES && p.gate != nullptr && row_ok)
                gate_row = p.gate + (row / p.tokens_per_frame) * p.gate_frame_stride;

            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * k2BN;
            constexpr int kChunksPerHalf = k2BN / 64;   // 32-column chunks each epilogue half owns
#pragma unroll 1
            for (int c = chalf * kChunksPerHalf; c < (chalf + 1) * kChunksPerHalf; ++c) {
                const int col0 = n_blk * k2BN + c * 32;
                if (col0 >= p.N) break;  // warp-uniform
                uint32_t acc[32];
                tmem_ld32(t_row + c * 32, acc);
                tmem_wait_ld();
                if (row_ok) gemm_epilogue_chunk<kEpi, kFp8>(p, acc, row, col0, gate_row);
            }
            tc_fence_before();
            mbar_arrive_cta(&tmem_empty[as], 0);   // leader's barrier, from either CTA
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

template <int kEpi, bool kFp8, int k2BN>
static ifx_status launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                               cudaStream_t stream) {
    constexpr int k2Smem = Gemm2Cfg<k2BN>::kSmem;
    static uint64_t configured = 0;     // per-device bit
    int dev = 0;
    IFX_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 64 || !(configured & (1ull << dev))) {
        IFX_CUDA_OK(cudaFuncSetAttribute(gemm2_tn_kernel<kEpi, kFp8, k2BN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, k2Smem));
        if (dev < 64) configured |= 1ull << dev;
    }
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    int clusters = sm_count() / 2;
    if (tiles < clusters) clusters = tiles;
    {
        char label[96];
        snprintf(label, sizeof(label), "gemm_%s_tn_kernel<%d,2cta,%d>[M=%lld,N=%d,K=%d]", kFp8 ? "fp8" : "bf16", kEpi,
                 k2BN, (long long)p.M, p.N, p.K);
        ProfScope prof(label, stream);
        IFX_CUDA_OK(launch_kernel(gemm2_tn_kernel<kEpi, kFp8, k2BN>, dim3(2 * clusters), dim3(k2Threads), k2Smem, stream,
                                  true, tmA, tmB, p));
    }
    IFX_LAUNCH_OK("gemm2_tn_kernel");
    return IFX_OK;
}

// Called by gemm_entry (ifx_gemm.cu) once arguments are validated.  bn = cluster tile width (256 or 128).
ifx_status gemm2_dispatch(bool fp8, int bn, const void* A, int64_t lda, const void* W, int64_t ldw, GemmParams p,
                          int epilogue, cudaStream_t stream) {
    CUtensorMap tmA, tmB;
    ifx_status st = fp8 ? make_tmap_u8_2d(&tmA, A, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)lda, 2 * k2BK, k2BM)
                        : make_tmap_bf16_2d(&tmA, A, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)lda, k2BK, k2BM);
    if (st != IFX_OK) return st;
    st = fp8 ? make_tmap_u8_2d(&tmB, W, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)ldw, 2 * k2BK, bn / 2)
             : make_tmap_bf16_2d(&tmB, W, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)ldw, k2BK, bn / 2);
    if (st != IFX_OK) return st;
    p.num_m_tiles = static_cast<int32_t>((p.M + 2 * k2BM - 1) / (2 * k2BM));
    p.num_n_tiles = (p.N + bn - 1) / bn;
#define IFX_G2(E)                                                                                         \
    do {                                                                                                  \
        if (bn == 256)                                                                                    \
            return fp8 ? launch_gemm2<E, true, 256>(tmA, tmB, p, stream) : launch_gemm2<E, false, 256>(tmA, tmB, p, stream); \
        return fp8 ? launch_gemm2<E, true, 128>(tmA, tmB, p, stream) : launch_gemm2<E, false, 128>(tmA, tmB, p, stream);     \
    } while (0)
    switch (epilogue) {
        case IFX_EPI_BIAS: IFX_G2(IFX_EPI_BIAS);
        case IFX_EPI_BIAS_GELU: IFX_G2(IFX_EPI_BIAS_GELU);
        case IFX_EPI_BIAS_GELU_ERF: IFX_G2(IFX_EPI_BIAS_GELU_ERF);
        case IFX_EPI_BIAS_F32: IFX_G2(IFX_EPI_BIAS_F32);
        default: IFX_G2(IFX_EPI_BIAS_GATE_RES);
    }
#undef IFX_G2
}

}  // namespace ifx
