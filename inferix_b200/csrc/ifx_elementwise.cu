// HBM-bound row kernels of the DiT block: LayerNorm+AdaLN modulation, WanRMSNorm, and the fused
// QK-RMSNorm -> 3-D RoPE (fp64) -> paged KV append.  One CTA per token row, 16-byte vector accesses,
// fp32 statistics; every reference rounding point (bf16 after each eager op) is reproduced explicitly.
#include "ifx_internal.h"
#include "ifx_ptx.cuh"

namespace ifx {

constexpr int kRowThreads = 256;
constexpr int kMaxVecPerThread = 8;  // cols <= 256 * 8 * 8; one warp per row up to 2048 columns

// threads per row CTA: one 16-byte vector per thread when the row fits, rounded up to whole warps
// Rows up to 2048 columns get ONE WARP (each lane keeps up to 8 x 16 B loads in flight and the row statistics need
// only shuffles); wider rows fall back to a multi-warp CTA with a shared-memory reduction.
// launch geometry of a row kernel (see RowGeom)
struct RowLaunch {
    unsigned grid, block;
    int warp_rows;
};
static inline int row_threads(int cols);
static inline RowLaunch row_launch(int64_t rows, int cols) {
    RowLaunch r;
    if ((cols >> 3) <= 32 * 8) {                 // one warp per row
        const int64_t ctas = (rows + 7) / 8;
        const int64_t cap = static_cast<int64_t>(sm_count()) * 4;   // 4 resident CTAs (32 warps) per SM
        r.grid = static_cast<unsigned>(ctas < cap ? ctas : cap);
        r.block = 256;
        r.warp_rows = 1;
    } else {
        r.grid = static_cast<unsigned>(rows);
        r.block = static_cast<unsigned>(row_threads(cols));
        r.warp_rows = 0;
    }
    return r;
}
static inline int row_threads(int cols) {
    const int nvec = cols >> 3;
    if (nvec <= 32 * kMaxVecPerThread) return 32;
    int t = ((nvec + kMaxVecPerThread - 1) / kMaxVecPerThread + 31) / 32 * 32;
    return t > kRowThreads ? kRowThreads : t;
}

// Row geometry shared by the row kernels.  Rows that fit one warp (<= 2048 columns: the Wan / MAGI widths) run in
// "warp rows" mode: a CTA is 8 independent warps and every warp walks rows with a grid-wide stride — a few hundred
// resident CTAs instead of one 32-thread CTA per token (whose launch rate, not HBM, bounded the round-1 kernels).
// Wider rows keep one multi-warp CTA per row.
struct RowGeom {
    int t;          // this thread's index inside its row group
    int tpr;        // threads per row
    int64_t row;    // first row of this group
    int64_t step;   // row stride
};
__device__ __forceinline__ RowGeom row_geom(int warp_rows) {
    RowGeom g;
    if (warp_rows) {
        g.t = threadIdx.x & 31;
        g.tpr = 32;
        g.row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
        g.step = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
    } else {
        g.t = threadIdx.x;
        g.tpr = blockDim.x;
        g.row = blockIdx.x;
        g.step = gridDim.x;
    }
    return g;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of up to two values; result broadcast to all threads
template <int kN>
__device__ __forceinline__ void block_sum(float (&v)[kN], float* scratch /* [kN][32] */, int tpr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = tpr >> 5;
#pragma unroll
    for (int i = 0; i < kN; ++i) v[i] = warp_sum(v[i]);
    if (nwarps == 1) {       // one warp per row: done (uniform branch)
        __syncwarp();
        return;
    }
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < kN; ++i) scratch[i * 32 + warp] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kN; ++i) {
        float t = lane < nwarps ? scratch[i * 32 + lane] : 0.f;
        v[i] = warp_sum(t);
    }
    __syncthreads();
}

__device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float2 t = __bfloat1622float2(h[e]);
        f[2 * e] = t.x;
        f[2 * e + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    return o;
}

// ------------------------------------------------------------------ LayerNorm + modulation
// e4m3( bf16( clamp(v / scale) ) ) for 8 values -> 8 bytes   (div_clamp_to, dit_module.py:367-387)
__device__ __forceinline__ uint2 quant8_e4m3(const float (&v)[8], float scale) {
    float q[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) q[e] = bf16_round(fminf(fmaxf(__fdiv_rn(v[e], scale), -448.f), 448.f));
    uint2 o;
    o.x = static_cast<uint32_t>(pack_e4m3x2(q[0], q[1])) | (static_cast<uint32_t>(pack_e4m3x2(q[2], q[3])) << 16);
    o.y = static_cast<uint32_t>(pack_e4m3x2(q[4], q[5])) | (static_cast<uint32_t>(pack_e4m3x2(q[6], q[7])) << 16);
    return o;
}

template <bool kOutFp8>
__global__ void __launch_bounds__(kRowThreads)
ln_modulate_kernel(const __nv_bfloat16* __restrict__ x, void* __restrict__ out_,
                   const __nv_bfloat16* __restrict__ ln_w, const __nv_bfloat16* __restrict__ ln_b,
                   const __nv_bfloat16* __restrict__ shift, const __nv_bfloat16* __restrict__ scale,
                   int64_t mod_frame_stride, int64_t rows, int cols, int64_t tokens_per_frame, float eps,
                   float out_scale, int warp_rows) {
    __shared__ float scratch[2 * 32];
    griddep_launch();
    griddep_wait();
    const RowGeom g = row_geom(warp_rows);
    const int nvec = cols >> 3;
    for (int64_t row = g.row; row < rows; row += g.step) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + row * cols);
        float v[kMaxVecPerThread][8];
        float acc[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < kMaxVecPerThread; ++i) {
            const int vi = g.t + i * g.tpr;
            if (vi < nvec) {
                unpack8(xr[vi], v[i]);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[0] += v[i][e];
            }
        }
        float s1[1] = {acc[0]};
        block_sum<1>(s1, scratch, g.tpr);
        const float mean = s1[0] / cols;
        // two-pass variance (matches at::native RowwiseMoments to fp32 rounding)
#pragma unroll
        for (int i = 0; i < kMaxVecPerThread; ++i) {
            const int vi = g.t + i * g.tpr;
            if (vi < nvec)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float d = v[i][e] - mean;
                    acc[1] += d * d;
                }
        }
        float s2[1] = {acc[1]};
        block_sum<1>(s2, scratch, g.tpr);
        const float rstd = rsqrtf(s2[0] / cols + eps);

        const int64_t frame = row / tokens_per_frame;
        const uint4* sh = shift ? reinterpret_cast<const uint4*>(shift + frame * mod_frame_stride) : nullptr;
        const uint4* sc = scale ? reinterpret_cast<const uint4*>(scale + frame * mod_frame_stride) : nullptr;
        uint4* orow = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(out_) + row * cols);
        uint2* orow8 = reinterpret_cast<uint2*>(static_cast<uint8_t*>(out_) + row * cols);
#pragma unroll
        for (int i = 0; i < kMaxVecPerThread; ++i) {
            const int vi = g.t + i * g.tpr;
            if (vi < nvec) {
                float y[8];
                if (ln_w) {
                    float w[8], b[8];
                    unpack8(__ldg(reinterpret_cast<const uint4*>(ln_w) + vi), w);
                    unpack8(__ldg(reinterpret_cast<const uint4*>(ln_b) + vi), b);
#pragma unroll
                    for (int e = 0; e < 8; ++e) y[e] = bf16_round((v[i][e] - mean) * rstd * w[e] + b[e]);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) y[e] = bf16_round((v[i][e] - mean) * rstd);
                }
                if (sc) {
                    float a[8], b[8];
                    unpack8(__ldg(sc + vi), a);
                    unpack8(__ldg(sh + vi), b);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float one_plus = bf16_round(1.0f + a[e]);
                        y[e] = bf16_round(y[e] * one_plus) + b[e];  // final rounding happens in pack8
                    }
                }
                if (kOutFp8) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) y[e] = bf16_round(y[e]);  // the bf16 tensor the reference would quantise
                    orow8[vi] = quant8_e4m3(y, out_scale);
                } else {
                    orow[vi] = pack8(y);
                }
            }
        }
    }
}

// ------------------------------------------------------------------ bf16 -> e4m3 quantisation
__global__ void __launch_bounds__(256)
quantize_fp8_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, uint8_t* __restrict__ out, int64_t ldo,
                    int64_t rows, int cols, float scale, const float* __restrict__ col_scale) {
    griddep_launch();
    griddep_wait();
    const int nvec = cols >> 3;
    const int64_t total = rows * nvec;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / nvec;
        const int vi = static_cast<int>(i % nvec);
        float v[8];
        unpack8(reinterpret_cast<const uint4*>(x + r * ldx)[vi], v);
        if (col_scale != nullptr) {
            // one scale per input channel (MAGI PerTensor input_scale vector / PerChannel smooth_scale)
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(col_scale) + 2 * vi);
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(col_scale) + 2 * vi + 1);
            const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            float q[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) q[e] = bf16_round(fminf(fmaxf(__fdiv_rn(v[e], sv[e]), -448.f), 448.f));
            uint2 o;
            o.x = static_cast<uint32_t>(pack_e4m3x2(q[0], q[1])) | (static_cast<uint32_t>(pack_e4m3x2(q[2], q[3])) << 16);
            o.y = static_cast<uint32_t>(pack_e4m3x2(q[4], q[5])) | (static_cast<uint32_t>(pack_e4m3x2(q[6], q[7])) << 16);
            reinterpret_cast<uint2*>(out + r * ldo)[vi] = o;
        } else {
            reinterpret_cast<uint2*>(out + r * ldo)[vi] = quant8_e4m3(v, scale);
        }
    }
}

// ------------------------------------------------------------------ WanRMSNorm
__global__ void __launch_bounds__(kRowThreads)
rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ w,
               __nv_bfloat16* __restrict__ out, int64_t ldo, int64_t rows, int cols, float eps, int warp_rows) {
    __shared__ float scratch[32];
    griddep_launch();
    griddep_wait();
    const RowGeom g = row_geom(warp_rows);
    const int nvec = cols >> 3;
    for (int64_t row = g.row; row < rows; row += g.step) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
        float v[kMaxVecPerThread][8];
        float ss[1] = {0.f};
#pragma unroll
        for (int i = 0; i < kMaxVecPerThread; ++i) {
            const int vi = g.t + i * g.tpr;
            if (vi < nvec) {
                unpack8(xr[vi], v[i]);
#pragma unroll
                for (int e = 0; e < 8; ++e) ss[0] += v[i][e] * v[i][e];
            }
        }
        block_sum<1>(ss, scratch, g.tpr);
        const float r = rsqrtf(ss[0] / cols + eps);
        uint4* orow = reinterpret_cast<uint4*>(out + row * ldo);
#pragma unroll
        for (int i = 0; i < kMaxVecPerThread; ++i) {
            const int vi = g.t + i * g.tpr;
            if (vi < nvec) {
                float wv[8], y[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(w) + vi), wv);
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] = bf16_round(v[i][e] * r) * wv[e];
                orow[vi] = pack8(y);
            }
        }
    }
}

// ------------------------------------------------------------------ QK RMSNorm + RoPE + KV append
struct NormRopeParams {
    const __nv_bfloat16* qkv;
    int64_t ld_qkv;
    const __nv_bfloat16* wq;
    const __nv_bfloat16* wk;
    const double2* freqs;  // [1024][hd/2] (cos, sin)
    ifx_rope_grid grid;
    __nv_bfloat16* q_out;
    int64_t ld_q;
    __nv_bfloat16* k_dst;  // cache base or contiguous staging
    __nv_bfloat16* v_dst;
    int32_t paged;         // 1: destination row = pages[t / page_tokens] * page_tokens + t % page_tokens
    int32_t page_tokens;
    PageList pl;
    int32_t C, heads, head_dim;
    float eps;
    // paged == 2 (sequence parallel over peer memory): this rank owns hw indices [hw_offset, hw_offset + hw_count) of
    // every frame of the block; its K / V rows are stored into the SAME cache row of every rank's replicated cache
    // (P2P stores over NVLink), then the last CTA publishes `epoch` in every rank's flag array.
    int32_t sp_world, sp_rank;
    __nv_bfloat16* peer_k[IFX_MAX_PEERS];
    __nv_bfloat16* peer_v[IFX_MAX_PEERS];
    long long* peer_flags[IFX_MAX_PEERS];
    long long epoch;
    unsigned int* done_counter;
    int32_t local_only;      // 1: store into this rank's cache only and publish nothing (ifx_peer_push does the exchange)
    int64_t rows;
    int32_t warp_rows;       // RowGeom mode
};

// pair index inside a head -> which RoPE axis it rotates with (causal_model.py:37: split [c-2(c/3), c/3, c/3])
__device__ __forceinline__ int rope_pos(int pair, int half, int t_pos, int h_pos, int w_pos) {
    const int third = half / 3;
    const int n_t = half - 2 * third;
    return pair < n_t ? t_pos : (pair < n_t + third ? h_pos : w_pos);
}

// kPeers: compile the peer-memory stores + epoch publication in (sequence parallel); the single-GPU instantiation
// keeps the register footprint and code of the plain append
template <bool kPeers>
__global__ void __launch_bounds__(kRowThreads)
qk_norm_rope_append_kernel(const NormRopeParams p) {
    __shared__ float scratch[2 * 32];
    // one 1-KiB table of rotation factors per row group (= per warp in warp-rows mode)
    __shared__ double2 cs_all[(kRowThreads / 32) * 128];
    griddep_launch();
    griddep_wait();
    const RowGeom g = row_geom(p.warp_rows);
    const int C = p.C;
    const int nvec = C >> 3;
    const int half = p.head_dim >> 1;
    double2* cs_s = cs_all + (p.warp_rows ? (threadIdx.x >> 5) * 128 : 0);

    for (int64_t t = g.row; t < p.rows; t += g.step) {   // t: token row
        const uint4* qr = reinterpret_cast<const uint4*>(p.qkv + t * p.ld_qkv);
        const uint4* kr = reinterpret_cast<const uint4*>(p.qkv + t * p.ld_qkv + C);
        const uint4* vr = reinterpret_cast<const uint4*>(p.qkv + t * p.ld_qkv + 2 * C);

        // destination row in the cache
        int64_t drow;
        if (p.paged == 1) {
            const int pg = static_cast<int>(t / p.page_tokens);
            drow = static_cast<int64_t>(p.pl.pages[pg]) * p.page_tokens + (t % p.page_tokens);
        } else if (kPeers && p.paged == 2) {
            // token index inside the block in single-process order: (frame, rank, hw)  (causal_model.py:1016-1021)
            const int64_t fs_full = static_cast<int64_t>(p.sp_world) * p.grid.hw_count;
            const int64_t tb = (t / p.grid.hw_count) * fs_full + p.grid.hw_offset + (t % p.grid.hw_count);
            drow = static_cast<int64_t>(p.pl.pages[tb / p.page_tokens]) * p.page_tokens + (tb % p.page_tokens);
        } else {
            drow = t;
        }
        // (frame, h, w) of this token; under sequence parallelism the rank owns hw indices [hw_offset, +hw_count)
        const int f = static_cast<int>(t / p.grid.hw_count);
        const int hw = p.grid.hw_offset + static_cast<int>(t % p.grid.hw_count);
        const int t_pos = p.grid.start_frame + f;
        const int h_pos = hw / p.grid.width;
        const int w_pos = hw % p.grid.width;

        // this token's rotation factors are shared by every head and by q and k: stage them once instead of
        // re-reading the table from L2 for each of the 2 x heads x 64 pairs
        if (p.warp_rows) __syncwarp(); else __syncthreads();      // previous row's readers are done
        for (int pr = g.t; pr < half; pr += g.tpr)
            cs_s[pr] = __ldg(&p.freqs[rope_pos(pr, half, t_pos, h_pos, w_pos) * half + pr]);
        if (p.warp_rows) __syncwarp(); else __syncthreads();

        float q[kMaxVecPerThread][8], k[kMaxVecPerThread][8];
        float ss[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < kMaxVecPerThread; ++i) {
            const int vi = g.t + i * g.tpr;
            if (vi < nvec) {
                unpack8(qr[vi], q[i]);
                unpack8(kr[vi], k[i]);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    ss[0] += q[i][e] * q[i][e];
                    ss[1] += k[i][e] * k[i][e];
                }
                // V is appended untouched
                if (kPeers && p.paged == 2) {
                    const uint4 vv = vr[vi];
                    if (p.local_only) {
                        reinterpret_cast<uint4*>(p.peer_v[p.sp_rank] + drow * C)[vi] = vv;
                    } else {
                        for (int d = 0; d < p.sp_world; ++d) {
                            const int dst = (p.sp_rank + 1 + d) % p.sp_world;  // start at the neighbour: spread the links
                            reinterpret_cast<uint4*>(p.peer_v[dst] + drow * C)[vi] = vv;
                        }
                    }
                } else {
                    reinterpret_cast<uint4*>(p.v_dst + drow * C)[vi] = vr[vi];
                }
            }
        }
        block_sum<2>(ss, scratch, g.tpr);
        const float rq = rsqrtf(ss[0] / C + p.eps);
        const float rk = rsqrtf(ss[1] / C + p.eps);
#pragma unroll
        for (int i = 0; i < kMaxVecPerThread; ++i) {
            const int vi = g.t + i * g.tpr;
            if (vi < nvec) {
                float wq[8], wk[8], qo[8], ko[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(p.wq) + vi), wq);
                unpack8(__ldg(reinterpret_cast<const uint4*>(p.wk) + vi), wk);
                const int col0 = vi * 8;
                const int pair0 = (col0 % p.head_dim) >> 1;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int pair = pair0 + e;
                    const double2 cs = cs_s[pair];
                    // RMSNorm: bf16(x * rsqrt) then bf16(.. * weight)  (components.py:118-126)
                    const double qa = bf16_round(bf16_round(q[i][2 * e] * rq) * wq[2 * e]);
                    const double qb = bf16_round(bf16_round(q[i][2 * e + 1] * rq) * wq[2 * e + 1]);
                    const double ka = bf16_round(bf16_round(k[i][2 * e] * rk) * wk[2 * e]);
                    const double kb = bf16_round(bf16_round(k[i][2 * e + 1] * rk) * wk[2 * e + 1]);
                    // complex multiply in fp64 (causal_model.py:46-56), rounded once to bf16
                    qo[2 * e] = static_cast<float>(qa * cs.x - qb * cs.y);
                    qo[2 * e + 1] = static_cast<float>(qa * cs.y + qb * cs.x);
                    ko[2 * e] = static_cast<float>(ka * cs.x - kb * cs.y);
                    ko[2 * e + 1] = static_cast<float>(ka * cs.y + kb * cs.x);
                }
                reinterpret_cast<uint4*>(p.q_out + t * p.ld_q)[vi] = pack8(qo);
                if (kPeers && p.paged == 2) {
                    const uint4 kk = pack8(ko);
                    if (p.local_only) {
                        reinterpret_cast<uint4*>(p.peer_k[p.sp_rank] + drow * C)[vi] = kk;
                    } else {
                        for (int d = 0; d < p.sp_world; ++d) {
                            const int dst = (p.sp_rank + 1 + d) % p.sp_world;
                            reinterpret_cast<uint4*>(p.peer_k[dst] + drow * C)[vi] = kk;
                        }
                    }
                } else {
                    reinterpret_cast<uint4*>(p.k_dst + drow * C)[vi] = pack8(ko);
                }
            }
        }
    }
    if (kPeers && p.paged == 2 && !p.local_only) {
        // every thread's peer stores are ordered before the CTA's arrival; the last CTA to arrive publishes the epoch
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int prev = atomicAdd(p.done_counter, 1u);
            if (prev == gridDim.x - 1) {
                *p.done_counter = 0;                       // next launch on this stream starts from zero
                __threadfence_system();
                for (int d = 0; d < p.sp_world; ++d) st_relaxed_sys(p.peer_flags[d] + p.sp_rank, p.epoch);
            }
        }
    }
}


// ================================================================== warp-per-row kernels (rows <= 2048 columns)
// The Wan / MAGI widths fit one warp per row.  A CTA is 8 independent warps walking rows with a grid-wide stride; the
// row stays PACKED in registers (kVec 16-byte vectors per lane) and is unpacked on use, which keeps the kernels at
// <= 64 registers = 4 CTAs (32 warps, ~100 KB of loads in flight) per SM.  Per-lane accumulation order and the shuffle
// tree are those of the generic kernels above, so results are bit-identical to them.
constexpr int kWarpRowCtaThreads = 256;

template <int kVec, bool kOutFp8>
__global__ void __launch_bounds__(kWarpRowCtaThreads, 3)
ln_modulate_warp_kernel(const __nv_bfloat16* __restrict__ x, void* __restrict__ out_,
                        const __nv_bfloat16* __restrict__ ln_w, const __nv_bfloat16* __restrict__ ln_b,
                        const __nv_bfloat16* __restrict__ shift, const __nv_bfloat16* __restrict__ scale,
                        int64_t mod_frame_stride, int64_t rows, int cols, int64_t tokens_per_frame, float eps,
                        float out_scale) {
    griddep_launch();
    griddep_wait();
    const int lane = threadIdx.x & 31;
    const int nvec = cols >> 3;
    const int64_t step = static_cast<int64_t>(gridDim.x) * (kWarpRowCtaThreads / 32);
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * (kWarpRowCtaThreads / 32) + (threadIdx.x >> 5); row < rows;
         row += step) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + row * cols);
        uint4 raw[kVec];
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
            const int vi = lane + i * 32;
            raw[i] = vi < nvec ? xr[vi] : make_uint4(0, 0, 0, 0);
        }
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; ++i)
            if (lane + i * 32 < nvec) {
                float v[8];
                unpack8(raw[i], v);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc += v[e];
            }
        const float mean = warp_sum(acc) / cols;
        acc = 0.f;   // two-pass variance (matches at::native RowwiseMoments to fp32 rounding)
#pragma unroll
        for (int i = 0; i < kVec; ++i)
            if (lane + i * 32 < nvec) {
                float v[8];
                unpack8(raw[i], v);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float d = v[e] - mean;
                    acc += d * d;
                }
            }
        const float rstd = rsqrtf(warp_sum(acc) / cols + eps);

        const int64_t frame = row / tokens_per_frame;
        const uint4* sh = shift ? reinterpret_cast<const uint4*>(shift + frame * mod_frame_stride) : nullptr;
        const uint4* sc = scale ? reinterpret_cast<const uint4*>(scale + frame * mod_frame_stride) : nullptr;
        uint4* orow = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(out_) + row * cols);
        uint2* orow8 = reinterpret_cast<uint2*>(static_cast<uint8_t*>(out_) + row * cols);
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
            const int vi = lane + i * 32;
            if (vi < nvec) {
                float v[8], y[8];
                unpack8(raw[i], v);
                if (ln_w) {
                    float w[8], b[8];
                    unpack8(__ldg(reinterpret_cast<const uint4*>(ln_w) + vi), w);
                    unpack8(__ldg(reinterpret_cast<const uint4*>(ln_b) + vi), b);
#pragma unroll
                    for (int e = 0; e < 8; ++e) y[e] = bf16_round((v[e] - mean) * rstd * w[e] + b[e]);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) y[e] = bf16_round((v[e] - mean) * rstd);
                }
                if (sc) {
                    float a[8], b[8];
                    unpack8(__ldg(sc + vi), a);
                    unpack8(__ldg(sh + vi), b);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float one_plus = bf16_round(1.0f + a[e]);
                        y[e] = bf16_round(y[e] * one_plus) + b[e];  // final rounding happens in pack8
                    }
                }
                if (kOutFp8) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) y[e] = bf16_round(y[e]);
                    orow8[vi] = quant8_e4m3(y, out_scale);
                } else {
                    orow[vi] = pack8(y);
                }
            }
        }
    }
}

template <int kVec>
__global__ void __launch_bounds__(kWarpRowCtaThreads, 4)
rmsnorm_warp_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ w,
                    __nv_bfloat16* __restrict__ out, int64_t ldo, int64_t rows, int cols, float eps) {
    griddep_launch();
    griddep_wait();
    const int lane = threadIdx.x & 31;
    const int nvec = cols >> 3;
    const int64_t step = static_cast<int64_t>(gridDim.x) * (kWarpRowCtaThreads / 32);
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * (kWarpRowCtaThreads / 32) + (threadIdx.x >> 5); row < rows;
         row += step) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
        uint4 raw[kVec];
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
            const int vi = lane + i * 32;
            raw[i] = vi < nvec ? xr[vi] : make_uint4(0, 0, 0, 0);
        }
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; ++i)
            if (lane + i * 32 < nvec) {
                float v[8];
                unpack8(raw[i], v);
#pragma unroll
                for (int e = 0; e < 8; ++e) ss += v[e] * v[e];
            }
        const float r = rsqrtf(warp_sum(ss) / cols + eps);
        uint4* orow = reinterpret_cast<uint4*>(out + row * ldo);
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
            const int vi = lane + i * 32;
            if (vi < nvec) {
                float v[8], wv[8], y[8];
                unpack8(raw[i], v);
                unpack8(__ldg(reinterpret_cast<const uint4*>(w) + vi), wv);
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] = bf16_round(v[e] * r) * wv[e];
                orow[vi] = pack8(y);
            }
        }
    }
}

template <int kVec, bool kPeers>
__global__ void __launch_bounds__(kWarpRowCtaThreads, 2)
qk_norm_rope_append_warp_kernel(const NormRopeParams p) {
    __shared__ double2 cs_all[(kWarpRowCtaThreads / 32) * 128];   // per warp: this token's rotation factors (<= 128 pairs)
    griddep_launch();
    griddep_wait();
    const int lane = threadIdx.x & 31;
    const int C = p.C;
    const int nvec = C >> 3;
    const int half = p.head_dim >> 1;
    double2* cs_s = cs_all + (threadIdx.x >> 5) * 128;
    const int64_t step = static_cast<int64_t>(gridDim.x) * (kWarpRowCtaThreads / 32);
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * (kWarpRowCtaThreads / 32) + (threadIdx.x >> 5); t < p.rows;
         t += step) {
        const uint4* qr = reinterpret_cast<const uint4*>(p.qkv + t * p.ld_qkv);
        const uint4* kr = reinterpret_cast<const uint4*>(p.qkv + t * p.ld_qkv + C);
        const uint4* vr = reinterpret_cast<const uint4*>(p.qkv + t * p.ld_qkv + 2 * C);
        uint4 qraw[kVec], kraw[kVec];
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
            const int vi = lane + i * 32;
            qraw[i] = vi < nvec ? qr[vi] : make_uint4(0, 0, 0, 0);
            kraw[i] = vi < nvec ? kr[vi] : make_uint4(0, 0, 0, 0);
        }
        // destination row in the cache
        int64_t drow;
        if (p.paged == 1) {
            const int pg = static_cast<int>(t / p.page_tokens);
            drow = static_cast<int64_t>(p.pl.pages[pg]) * p.page_tokens + (t % p.page_tokens);
        } else if (kPeers && p.paged == 2) {
            // token index inside the block in single-process order: (frame, rank, hw)  (causal_model.py:1016-1021)
            const int64_t fs_full = static_cast<int64_t>(p.sp_world) * p.grid.hw_count;
            const int64_t tb = (t / p.grid.hw_count) * fs_full + p.grid.hw_offset + (t % p.grid.hw_count);
            drow = static_cast<int64_t>(p.pl.pages[tb / p.page_tokens]) * p.page_tokens + (tb % p.page_tokens);
        } else {
            drow = t;
        }
        // V is appended untouched
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
            const int vi = lane + i * 32;
            if (vi < nvec) {
                const uint4 vv = vr[vi];
                if (kPeers && p.paged == 2) {
                    if (p.local_only) {
                        reinterpret_cast<uint4*>(p.peer_v[p.sp_rank] + drow * C)[vi] = vv;
                    } else {
                        for (int d = 0; d < p.sp_world; ++d) {
                            const int dst = (p.sp_rank + 1 + d) % p.sp_world;  // start at the neighbour: spread the links
                            reinterpret_cast<uint4*>(p.peer_v[dst] + drow * C)[vi] = vv;
                        }
                    }
                } else {
                    reinterpret_cast<uint4*>(p.v_dst + drow * C)[vi] = vv;
                }
            }
        }
        // (frame, h, w) of this token; under sequence parallelism the rank owns hw indices [hw_offset, +hw_count)
        const int f = static_cast<int>(t / p.grid.hw_count);
        const int hw = p.grid.hw_offset + static_cast<int>(t % p.grid.hw_count);
        const int t_pos = p.grid.start_frame + f;
        const int h_pos = hw / p.grid.width;
        const int w_pos = hw % p.grid.width;
        __syncwarp();                                  // previous row's readers are done with cs_s
        for (int pr = lane; pr < half; pr += 32)
            cs_s[pr] = __ldg(&p.freqs[rope_pos(pr, half, t_pos, h_pos, w_pos) * half + pr]);
        __syncwarp();

        float sq = 0.f, sk = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; ++i)
            if (lane + i * 32 < nvec) {
                float a[8], b[8];
                unpack8(qraw[i], a);
                unpack8(kraw[i], b);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    sq += a[e] * a[e];
                    sk += b[e] * b[e];
                }
            }
        const float rq = rsqrtf(warp_sum(sq) / C + p.eps);
        const float rk = rsqrtf(warp_sum(sk) / C + p.eps);
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
            const int vi = lane + i * 32;
            if (vi < nvec) {
                float q[8], k[8], wq[8], wk[8], qo[8], ko[8];
                unpack8(qraw[i], q);
                unpack8(kraw[i], k);
                unpack8(__ldg(reinterpret_cast<const uint4*>(p.wq) + vi), wq);
                unpack8(__ldg(reinterpret_cast<const uint4*>(p.wk) + vi), wk);
                const int pair0 = ((vi * 8) % p.head_dim) >> 1;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const double2 cs = cs_s[pair0 + e];
                    // RMSNorm: bf16(x * rsqrt) then bf16(.. * weight)  (components.py:118-126)
                    const double qa = bf16_round(bf16_round(q[2 * e] * rq) * wq[2 * e]);
                    const double qb = bf16_round(bf16_round(q[2 * e + 1] * rq) * wq[2 * e + 1]);
                    const double ka = bf16_round(bf16_round(k[2 * e] * rk) * wk[2 * e]);
                    const double kb = bf16_round(bf16_round(k[2 * e + 1] * rk) * wk[2 * e + 1]);
                    // complex multiply in fp64 (causal_model.py:46-56), rounded once to bf16
                    qo[2 * e] = static_cast<float>(qa * cs.x - qb * cs.y);
                    qo[2 * e + 1] = static_cast<float>(qa * cs.y + qb * cs.x);
                    ko[2 * e] = static_cast<float>(ka * cs.x - kb * cs.y);
                    ko[2 * e + 1] = static_cast<float>(ka * cs.y + kb * cs.x);
                }
                reinterpret_cast<uint4*>(p.q_out + t * p.ld_q)[vi] = pack8(qo);
                const uint4 kk = pack8(ko);
                if (kPeers && p.paged == 2) {
                    if (p.local_only) {
                        reinterpret_cast<uint4*>(p.peer_k[p.sp_rank] + drow * C)[vi] = kk;
                    } else {
                        for (int d = 0; d < p.sp_world; ++d) {
                            const int dst = (p.sp_rank + 1 + d) % p.sp_world;
                            reinterpret_cast<uint4*>(p.peer_k[dst] + drow * C)[vi] = kk;
                        }
                    }
                } else {
                    reinterpret_cast<uint4*>(p.k_dst + drow * C)[vi] = kk;
                }
            }
        }
    }
    if (kPeers && p.paged == 2 && !p.local_only) {
        // every thread's peer stores are ordered before the CTA's arrival; the last CTA to arrive publishes the epoch
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int prev = atomicAdd(p.done_counter, 1u);
            if (prev == gridDim.x - 1) {
                *p.done_counter = 0;                       // next launch on this stream starts from zero
                __threadfence_system();
                for (int d = 0; d < p.sp_world; ++d) st_relaxed_sys(p.peer_flags[d] + p.sp_rank, p.epoch);
            }
        }
    }
}

// kVec dispatch: smallest instantiated vector count per lane that covers the row
#define IFX_WARP_ROW_DISPATCH(nvec, CALL)            \
    do {                                             \
        const int _k = ((nvec) + 31) / 32;           \
        if (_k <= 1) { CALL(1); }                    \
        else if (_k <= 2) { CALL(2); }               \
        else if (_k <= 4) { CALL(4); }               \
        else if (_k <= 6) { CALL(6); }               \
        else { CALL(8); }                            \
    } while (0)

// ================================================================== dynamic per-token quantisation (8-bit linears)
// x[row, :] -> q[row, :] = round(x / s_row) with s_row = max(|x[row, :]|, 1e-12) / qmax  (qmax 448: e4m3 RNE with
// saturation; 127: int8 RNE), scales[row] = s_row.  The activation half of "dynamic per-token activation x per-channel
// weight" quantisation (the qconfig the reference's quantisation examples request from DAX,
// example/quantization/run_causvid_quantized.py:32-37; DAX itself is not vendored — algorithm restated, unpinned).
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint2 quant8_dyn(const float (&v)[8], float inv_s, int int8) {
    uint2 o;
    if (int8) {
        uint32_t b[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int q;
            asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(q) : "f"(v[e] * inv_s));
            q = max(q, -127);
            b[e] = static_cast<uint32_t>(q) & 0xffu;
        }
        o.x = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
        o.y = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
    } else {
        float q[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) q[e] = v[e] * inv_s;
        o.x = static_cast<uint32_t>(pack_e4m3x2(q[0], q[1])) | (static_cast<uint32_t>(pack_e4m3x2(q[2], q[3])) << 16);
        o.y = static_cast<uint32_t>(pack_e4m3x2(q[4], q[5])) | (static_cast<uint32_t>(pack_e4m3x2(q[6], q[7])) << 16);
    }
    return o;
}

// one warp per row, any width (two passes over the row; the second one hits L2)
__global__ void __launch_bounds__(kWarpRowCtaThreads)
quantize_rows_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, uint8_t* __restrict__ out, int64_t ldo,
                     float* __restrict__ scales, int64_t rows, int cols, int int8) {
    griddep_launch();
    griddep_wait();
    const int lane = threadIdx.x & 31;
    const int nvec = cols >> 3;
    const float qmax = int8 ? 127.0f : 448.0f;
    const int64_t step = static_cast<int64_t>(gridDim.x) * (kWarpRowCtaThreads / 32);
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * (kWarpRowCtaThreads / 32) + (threadIdx.x >> 5); row < rows;
         row += step) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
        float amax = 0.f;
        for (int vi = lane; vi < nvec; vi += 32) {
            float v[8];
            unpack8(xr[vi], v);
#pragma unroll
            for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(v[e]));
        }
        amax = warp_max(amax);
        const float s = fmaxf(amax, 1e-12f) / qmax;
        const float inv_s = 1.0f / s;
        if (lane == 0) scales[row] = s;
        uint2* orow = reinterpret_cast<uint2*>(out + row * ldo);
        for (int vi = lane; vi < nvec; vi += 32) {
            float v[8];
            unpack8(xr[vi], v);
            orow[vi] = quant8_dyn(v, inv_s, int8);
        }
    }
}

// ln_modulate_warp_kernel whose bf16 result is quantised per token on the way out (the row never leaves registers)
template <int kVec>
__global__ void __launch_bounds__(kWarpRowCtaThreads, 3)
ln_modulate_quant_warp_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ out, float* __restrict__ scales,
                              const __nv_bfloat16* __restrict__ ln_w, const __nv_bfloat16* __restrict__ ln_b,
                              const __nv_bfloat16* __restrict__ shift, const __nv_bfloat16* __restrict__ scale,
                              int64_t mod_frame_stride, int64_t rows, int cols, int64_t tokens_per_frame, float eps,
                              int int8) {
    griddep_launch();
    griddep_wait();
    const int lane = threadIdx.x & 31;
    const int nvec = cols >> 3;
    const float qmax = int8 ? 127.0f : 448.0f;
    const int64_t step = static_cast<int64_t>(gridDim.x) * (kWarpRowCtaThreads / 32);
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * (kWarpRowCtaThreads / 32) + (threadIdx.x >> 5); row < rows;
         row += step) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + row * cols);
        uint4 raw[kVec];
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
            const int vi = lane + i * 32;
            raw[i] = vi < nvec ? xr[vi] : make_uint4(0, 0, 0, 0);
        }
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; ++i)
            if (lane + i * 32 < nvec) {
                float v[8];
                unpack8(raw[i], v);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc += v[e];
            }
        const float mean = warp_sum(acc) / cols;
        acc = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; ++i)
            if (lane + i * 32 < nvec) {
                float v[8];
                unpack8(raw[i], v);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float d = v[e] - mean;
                    acc += d * d;
                }
            }
        const float rstd = rsqrtf(warp_sum(acc) / cols + eps);
        const int64_t frame = row / tokens_per_frame;
        const uint4* sh = shift ? reinterpret_cast<const uint4*>(shift + frame * mod_frame_stride) : nullptr;
        const uint4* sc = scale ? reinterpret_cast<const uint4*>(scale + frame * mod_frame_stride) : nullptr;
        // the bf16 tensor the reference would hand to the quantised linear, one vector at a time
        auto result = [&](int i, float (&y)[8]) {
            const int vi = lane + i * 32;
            float v[8];
            unpack8(raw[i], v);
            if (ln_w) {
                float w[8], b[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(ln_w) + vi), w);
                unpack8(__ldg(reinterpret_cast<const uint4*>(ln_b) + vi), b);
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] = bf16_round((v[e] - mean) * rstd * w[e] + b[e]);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] = bf16_round((v[e] - mean) * rstd);
            }
            if (sc) {
                float a[8], b[8];
                unpack8(__ldg(sc + vi), a);
                unpack8(__ldg(sh + vi), b);
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] = bf16_round(bf16_round(y[e] * bf16_round(1.0f + a[e])) + b[e]);
            }
        };
        float amax = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; ++i)
            if (lane + i * 32 < nvec) {
                float y[8];
                result(i, y);
#pragma unroll
                for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(y[e]));
            }
        amax = warp_max(amax);
        const float s = fmaxf(amax, 1e-12f) / qmax;
        const float inv_s = 1.0f / s;
        if (lane == 0) scales[row] = s;
        uint2* orow = reinterpret_cast<uint2*>(out + row * cols);
#pragma unroll
        for (int i = 0; i < kVec; ++i)
            if (lane + i * 32 < nvec) {
                float y[8];
                result(i, y);
                orow[lane + i * 32] = quant8_dyn(y, inv_s, int8);
            }
    }
}

// Spin until every rank has published `epoch` in this rank's flag array (one lane per source rank).  Bounded: a rank
// that never arrives traps the kernel after timeout_ns instead of hanging the GPU.
__global__ void peer_wait_kernel(const long long* flags, int world, long long epoch, unsigned long long timeout_ns) {
    griddep_launch();
    griddep_wait();
    const int s = threadIdx.x;
    if (s >= world) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(flags + s) < epoch) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (now - t0 > timeout_ns) {
            printf("ifx peer_wait: rank %d did not publish epoch %lld (have %lld)\n", s, epoch, ld_acquire_sys(flags + s));
            __trap();
        }
        __nanosleep(64);
    }
}

// ------------------------------------------------------------------ paged copy kernels (append / export / import)
__global__ void __launch_bounds__(256)
paged_copy_kernel(const PagedCopyParams p) {
    const int nvec = p.C >> 3;
    const int64_t total = p.rows * nvec;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / nvec;
        const int vi = static_cast<int>(i % nvec);
        const int64_t lt = p.first_logical + r;
        const int64_t lp = lt / p.page_tokens;
        int64_t crow;
        if (p.pl.n == 0) {
            crow = p.linear_row0 + r;
        } else {
            const int phys = p.pl.pages[lp - p.first_logical / p.page_tokens];
            crow = static_cast<int64_t>(phys) * p.page_tokens + lt % p.page_tokens;
        }
        if (p.mode == 0) {
            if (p.lin_k)
                reinterpret_cast<uint4*>(p.cache_k + crow * p.C)[vi] =
                    reinterpret_cast<const uint4*>(p.lin_k + r * p.ld_lin)[vi];
            if (p.lin_v)
                reinterpret_cast<uint4*>(p.cache_v + crow * p.C)[vi] =
                    reinterpret_cast<const uint4*>(p.lin_v + r * p.ld_lin)[vi];
        } else {
            if (p.lin_k)
                reinterpret_cast<uint4*>(p.lin_k + r * p.ld_lin)[vi] =
                    reinterpret_cast<const uint4*>(p.cache_k + crow * p.C)[vi];
            if (p.lin_v)
                reinterpret_cast<uint4*>(p.lin_v + r * p.ld_lin)[vi] =
                    reinterpret_cast<const uint4*>(p.cache_v + crow * p.C)[vi];
        }
    }
}

// Copy this rank's rows of the block's new pages from its own cache to the same rows of every other rank's cache and
// publish the epoch: the exchange half of qk_norm_rope_append_kernel<true>, as a separate small grid that runs on a
// side stream next to the attention over the already-cached pages (which leaves a few SMs free at 4-8 ranks).
__global__ void __launch_bounds__(1024)
peer_push_kernel(const PeerPushParams p) {
    // This grid is itself launched programmatically behind the norm + RoPE kernel that wrote the rows it ships: wait
    // for that kernel first, THEN release the attention kernel queued behind this grid.  The attention does not wait
    // for this grid (it needs nothing written locally here; the peers' rows are ordered by the epoch flags), but it
    // must not start before the norm + RoPE kernel has produced q and the local rows — hence wait before launch.
    griddep_wait();
    griddep_launch();
    const int nvec = p.C >> 3;
    const int64_t rows = static_cast<int64_t>(p.frames) * p.chunk;
    const int64_t fs_full = static_cast<int64_t>(p.world) * p.chunk;
    const int64_t total = rows * nvec * 2;                               // K and V
    const __nv_bfloat16* src[2] = {p.peer_k[p.rank], p.peer_v[p.rank]};
    constexpr int kUnroll = 4;                                           // 16-byte loads in flight per thread
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i0 < total; i0 += stride * kUnroll) {
        uint4 val[kUnroll];
        int64_t off[kUnroll];
        int which[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t i = i0 + u * stride;
            off[u] = -1;
            if (i < total) {
                which[u] = static_cast<int>(i / (rows * nvec));
                const int64_t r = (i / nvec) % rows;
                const int vi = static_cast<int>(i % nvec);
                const int64_t tb = (r / p.chunk) * fs_full + static_cast<int64_t>(p.rank) * p.chunk + r % p.chunk;
                const int64_t crow =
                    static_cast<int64_t>(p.pl.pages[tb / p.page_tokens]) * p.page_tokens + tb % p.page_tokens;
                off[u] = crow * p.C + vi * 8;
                val[u] = *reinterpret_cast<const uint4*>(src[which[u]] + off[u]);
            }
        }
        for (int d = 1; d < p.world; ++d) {
            const int dst = (p.rank + d) % p.world;                      // start at the neighbour: spread the links
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                if (off[u] >= 0)
                    *reinterpret_cast<uint4*>((which[u] ? p.peer_v[dst] : p.peer_k[dst]) + off[u]) = val[u];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(p.done_counter, 1u);
        if (prev == gridDim.x - 1) {
            *p.done_counter = 0;
            __threadfence_system();
            for (int d = 0; d < p.world; ++d) st_relaxed_sys(p.peer_flags[d] + p.rank, p.epoch);
        }
    }
}

struct SpAppendParams {
    __nv_bfloat16* cache_k;
    __nv_bfloat16* cache_v;
    const __nv_bfloat16* src_k;
    const __nv_bfloat16* src_v;
    int64_t rank_stride;
    int32_t world, frames, chunk, page_tokens, C;
    PageList pl;
};

__global__ void __launch_bounds__(256)
sp_append_kernel(const SpAppendParams p) {
    const int nvec = p.C >> 3;
    const int64_t rows_per_rank = static_cast<int64_t>(p.frames) * p.chunk;
    const int64_t total = rows_per_rank * p.world * nvec;
    const int64_t fs = static_cast<int64_t>(p.world) * p.chunk;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t row = i / nvec;
        const int vi = static_cast<int>(i % nvec);
        const int64_t r = row / rows_per_rank, rem = row % rows_per_rank;
        const int64_t t = (rem / p.chunk) * fs + r * p.chunk + rem % p.chunk;  // token index inside the block
        const int64_t crow = static_cast<int64_t>(p.pl.pages[t / p.page_tokens]) * p.page_tokens + t % p.page_tokens;
        const int64_t srow = r * p.rank_stride + rem * p.C;
        reinterpret_cast<uint4*>(p.cache_k + crow * p.C)[vi] = reinterpret_cast<const uint4*>(p.src_k + srow)[vi];
        reinterpret_cast<uint4*>(p.cache_v + crow * p.C)[vi] = reinterpret_cast<const uint4*>(p.src_v + srow)[vi];
    }
}

}  // namespace ifx

using namespace ifx;

extern "C" ifx_status ifx_kv_append_sp(ifx_kv* kv_, const ifx_kv_plan* plan, const void* k_src, const void* v_src,
                                       int64_t src_rank_stride, int32_t world, int32_t frames, int32_t chunk,
                                       void* stream) {
    KvImpl* kv = kv_cast(kv_);
    if (!kv) return set_error(IFX_ERR_HANDLE, "ifx_kv_append_sp: bad kv handle");
    IFX_CHECK_ARG(plan && k_src && v_src, "ifx_kv_append_sp: null pointer");
    IFX_CHECK_ARG(world > 0 && frames > 0 && chunk > 0, "ifx_kv_append_sp: bad geometry");
    const int64_t rows = static_cast<int64_t>(world) * frames * chunk;
    IFX_CHECK_ARG(rows == plan->local_end - plan->local_start && rows == (int64_t)plan->num_pages * kv->page_tokens,
                  "ifx_kv_append_sp: world*frames*chunk (%lld) does not match the plan", (long long)rows);
    SpAppendParams p;
    p.cache_k = static_cast<__nv_bfloat16*>(kv->k_base);
    p.cache_v = static_cast<__nv_bfloat16*>(kv->v_base);
    p.src_k = static_cast<const __nv_bfloat16*>(k_src);
    p.src_v = static_cast<const __nv_bfloat16*>(v_src);
    p.world = world;
    p.frames = frames;
    p.chunk = chunk;
    p.page_tokens = kv->page_tokens;
    p.C = kv->heads * kv->head_dim;
    IFX_CHECK_ARG(src_rank_stride >= static_cast<int64_t>(frames) * chunk * p.C && src_rank_stride % 8 == 0,
                  "ifx_kv_append_sp: bad src_rank_stride");
    p.rank_stride = src_rank_stride;
    p.pl.n = plan->num_pages;
    for (int i = 0; i < plan->num_pages; ++i) p.pl.pages[i] = plan->pages[i];
    const int64_t total = rows * (p.C >> 3);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    {
        ProfScope prof("sp_append_kernel", static_cast<cudaStream_t>(stream));
        sp_append_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    }
    IFX_LAUNCH_OK("sp_append_kernel");
    return IFX_OK;
}

static ifx_status ln_modulate_entry(const void* x, void* out, const void* ln_weight, const void* ln_bias,
                                    const void* shift, const void* scale, int64_t mod_frame_stride, int64_t rows,
                                    int32_t cols, int64_t tokens_per_frame, float eps, bool fp8, float out_scale,
                                    void* stream) {
    IFX_CHECK_ARG(x && out, "ifx_ln_modulate: null pointer");
    IFX_CHECK_ARG(rows > 0 && cols > 0 && cols % 8 == 0 && cols <= kRowThreads * kMaxVecPerThread * 8,
                  "ifx_ln_modulate: cols must be a multiple of 8 and <= %d (got %d)",
                  kRowThreads * kMaxVecPerThread * 8, cols);
    IFX_CHECK_ARG((ln_weight == nullptr) == (ln_bias == nullptr), "ifx_ln_modulate: weight and bias go together");
    IFX_CHECK_ARG((shift == nullptr) == (scale == nullptr), "ifx_ln_modulate: shift and scale go together");
    IFX_CHECK_ARG(!scale || (tokens_per_frame > 0 && mod_frame_stride % 8 == 0),
                  "ifx_ln_modulate: bad modulation layout");
    IFX_CHECK_ARG(!fp8 || out_scale > 0.f, "ifx_ln_modulate_fp8: out_scale must be positive");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t tpf = tokens_per_frame > 0 ? tokens_per_frame : 1;
    {
        ProfScope prof(fp8 ? "ln_modulate_kernel<fp8>" : "ln_modulate_kernel", st);
        const RowLaunch rl = row_launch(rows, cols);
        if (rl.warp_rows) {
#define IFX_LN_CALL(K)                                                                                                \
    IFX_CUDA_OK(launch_kernel(fp8 ? ln_modulate_warp_kernel<K, true> : ln_modulate_warp_kernel<K, false>, dim3(rl.grid),  \
                              dim3(rl.block), 0, st, true, static_cast<const __nv_bfloat16*>(x), out,                 \
                              static_cast<const __nv_bfloat16*>(ln_weight), static_cast<const __nv_bfloat16*>(ln_bias), \
                              static_cast<const __nv_bfloat16*>(shift), static_cast<const __nv_bfloat16*>(scale),     \
                              mod_frame_stride, rows, cols, tpf, eps, fp8 ? out_scale : 1.0f))
            IFX_WARP_ROW_DISPATCH(cols >> 3, IFX_LN_CALL);
#undef IFX_LN_CALL
        } else {
            IFX_CUDA_OK(launch_kernel(fp8 ? ln_modulate_kernel<true> : ln_modulate_kernel<false>, dim3(rl.grid),
                                      dim3(rl.block), 0, st, true, static_cast<const __nv_bfloat16*>(x), out,
                                      static_cast<const __nv_bfloat16*>(ln_weight),
                                      static_cast<const __nv_bfloat16*>(ln_bias), static_cast<const __nv_bfloat16*>(shift),
                                      static_cast<const __nv_bfloat16*>(scale), mod_frame_stride, rows, cols, tpf, eps,
                                      fp8 ? out_scale : 1.0f, 0));
        }
    }
    IFX_LAUNCH_OK("ln_modulate_kernel");
    return IFX_OK;
}

extern "C" ifx_status ifx_ln_modulate(const void* x, void* out, const void* ln_weight, const void* ln_bias,
                                      const void* shift, const void* scale, int64_t mod_frame_stride, int64_t rows,
                                      int32_t cols, int64_t tokens_per_frame, float eps, void* stream) {
    return ln_modulate_entry(x, out, ln_weight, ln_bias, shift, scale, mod_frame_stride, rows, cols, tokens_per_frame,
                             eps, false, 1.0f, stream);
}

extern "C" ifx_status ifx_ln_modulate_fp8(const void* x, void* out, const void* ln_weight, const void* ln_bias,
                                          const void* shift, const void* scale, int64_t mod_frame_stride,
                                          int64_t rows, int32_t cols, int64_t tokens_per_frame, float eps,
                                          float out_scale, void* stream) {
    return ln_modulate_entry(x, out, ln_weight, ln_bias, shift, scale, mod_frame_stride, rows, cols, tokens_per_frame,
                             eps, true, out_scale, stream);
}

static ifx_status quantize_fp8_entry(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int32_t cols,
                                     float scale, const float* col_scale, void* stream) {
    IFX_CHECK_ARG(x && out, "ifx_quantize_fp8: null pointer");
    IFX_CHECK_ARG(rows > 0 && cols > 0 && cols % 8 == 0, "ifx_quantize_fp8: cols must be a multiple of 8");
    IFX_CHECK_ARG(ldx >= cols && ldx % 8 == 0 && ldo >= cols && ldo % 8 == 0, "ifx_quantize_fp8: bad strides");
    IFX_CHECK_ARG(col_scale != nullptr || scale > 0.f, "ifx_quantize_fp8: scale must be positive");
    IFX_CHECK_ARG(col_scale == nullptr || (reinterpret_cast<uintptr_t>(col_scale) & 15) == 0,
                  "ifx_quantize_fp8_cols: col_scale must be 16-byte aligned");
    const int64_t total = rows * (cols >> 3);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ProfScope prof("quantize_fp8_kernel", st);
        IFX_CUDA_OK(launch_kernel(quantize_fp8_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, true,
                                  static_cast<const __nv_bfloat16*>(x), ldx, static_cast<uint8_t*>(out), ldo, rows, cols,
                                  scale, col_scale));
    }
    IFX_LAUNCH_OK("quantize_fp8_kernel");
    return IFX_OK;
}

extern "C" ifx_status ifx_quantize_fp8(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int32_t cols,
                                       float scale, void* stream) {
    return quantize_fp8_entry(x, ldx, out, ldo, rows, cols, scale, nullptr, stream);
}

extern "C" ifx_status ifx_quantize_fp8_cols(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows,
                                            int32_t cols, const float* col_scale, void* stream) {
    IFX_CHECK_ARG(col_scale != nullptr, "ifx_quantize_fp8_cols: null col_scale");
    return quantize_fp8_entry(x, ldx, out, ldo, rows, cols, 1.0f, col_scale, stream);
}

extern "C" ifx_status ifx_rmsnorm(const void* x, int64_t ldx, const void* weight, void* out, int64_t ldo,
                                  int64_t rows, int32_t cols, float eps, void* stream) {
    IFX_CHECK_ARG(x && out && weight, "ifx_rmsnorm: null pointer");
    IFX_CHECK_ARG(rows > 0 && cols > 0 && cols % 8 == 0 && cols <= kRowThreads * kMaxVecPerThread * 8,
                  "ifx_rmsnorm: bad cols %d", cols);
    IFX_CHECK_ARG(ldx % 8 == 0 && ldo % 8 == 0 && ldx >= cols && ldo >= cols, "ifx_rmsnorm: bad strides");
    {
        ProfScope prof("rmsnorm_kernel", static_cast<cudaStream_t>(stream));
        const RowLaunch rl = row_launch(rows, cols);
        if (rl.warp_rows) {
#define IFX_RMS_CALL(K)                                                                                               \
    IFX_CUDA_OK(launch_kernel(rmsnorm_warp_kernel<K>, dim3(rl.grid), dim3(rl.block), 0, static_cast<cudaStream_t>(stream), \
                              true, static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(weight), \
                              static_cast<__nv_bfloat16*>(out), ldo, rows, cols, eps))
            IFX_WARP_ROW_DISPATCH(cols >> 3, IFX_RMS_CALL);
#undef IFX_RMS_CALL
        } else {
            IFX_CUDA_OK(launch_kernel(rmsnorm_kernel, dim3(rl.grid), dim3(rl.block), 0, static_cast<cudaStream_t>(stream),
                                      true, static_cast<const __nv_bfloat16*>(x), ldx,
                                      static_cast<const __nv_bfloat16*>(weight), static_cast<__nv_bfloat16*>(out), ldo, rows,
                                      cols, eps, 0));
        }
    }
    IFX_LAUNCH_OK("rmsnorm_kernel");
    return IFX_OK;
}

static ifx_status norm_rope_append_entry(const void* qkv, int64_t ld_qkv, const void* norm_q_weight,
                                         const void* norm_k_weight, const double* freqs, const ifx_rope_grid* grid,
                                         void* q_out, int64_t ld_q, ifx_kv* kv_, const ifx_kv_plan* plan, void* k_dst,
                                         void* v_dst, const ifx_peer_dst* peers, int64_t rows, int32_t heads,
                                         int32_t head_dim, float eps, void* stream) {
    IFX_CHECK_ARG(qkv && norm_q_weight && norm_k_weight && freqs && grid && q_out, "ifx_qk_norm_rope_append: null");
    const int C = heads * head_dim;
    IFX_CHECK_ARG(rows > 0 && C % 8 == 0 && C <= kRowThreads * kMaxVecPerThread * 8 && head_dim % 8 == 0 && head_dim <= 256,
                  "ifx_qk_norm_rope_append: bad shape heads=%d head_dim=%d", heads, head_dim);
    IFX_CHECK_ARG(ld_qkv >= 3 * C && ld_qkv % 8 == 0 && ld_q >= C && ld_q % 8 == 0,
                  "ifx_qk_norm_rope_append: bad strides");
    IFX_CHECK_ARG(grid->hw_count > 0 && grid->width > 0 && rows == (int64_t)grid->frames * grid->hw_count,
                  "ifx_qk_norm_rope_append: rows (%lld) != frames*hw_count (%d*%d)", (long long)rows, grid->frames,
                  grid->hw_count);
    IFX_CHECK_ARG(grid->start_frame >= 0 && grid->start_frame + grid->frames <= 1024 && grid->height <= 1024 &&
                      grid->width <= 1024,
                  "ifx_qk_norm_rope_append: RoPE table has 1024 positions");
    NormRopeParams p = {};
    p.qkv = static_cast<const __nv_bfloat16*>(qkv);
    p.ld_qkv = ld_qkv;
    p.wq = static_cast<const __nv_bfloat16*>(norm_q_weight);
    p.wk = static_cast<const __nv_bfloat16*>(norm_k_weight);
    p.freqs = reinterpret_cast<const double2*>(freqs);
    p.grid = *grid;
    p.q_out = static_cast<__nv_bfloat16*>(q_out);
    p.ld_q = ld_q;
    p.C = C;
    p.heads = heads;
    p.head_dim = head_dim;
    p.eps = eps;
    if (kv_ != nullptr) {
        KvImpl* kv = kv_cast(kv_);
        if (!kv) return set_error(IFX_ERR_HANDLE, "ifx_qk_norm_rope_append: bad kv handle");
        IFX_CHECK_ARG(plan != nullptr, "ifx_qk_norm_rope_append: plan required with kv");
        IFX_CHECK_ARG(kv->heads == heads && kv->head_dim == head_dim, "ifx_qk_norm_rope_append: kv shape mismatch");
        const int64_t plan_rows = peers ? rows * peers->world : rows;
        IFX_CHECK_ARG(plan->local_end - plan->local_start == plan_rows, "ifx_qk_norm_rope_append: plan covers %lld rows",
                      (long long)(plan->local_end - plan->local_start));
        IFX_CHECK_ARG((int64_t)plan->num_pages * kv->page_tokens == plan_rows && plan->first_offset == 0,
                      "ifx_qk_norm_rope_append: append must be page aligned");
        p.k_dst = static_cast<__nv_bfloat16*>(kv->k_base);
        p.v_dst = static_cast<__nv_bfloat16*>(kv->v_base);
        p.paged = 1;
        if (peers) {
            static unsigned int* g_done_dev[64] = {nullptr};   // one arrival counter per device (stream-ordered launches)
            unsigned int* g_done = nullptr;
            IFX_TRY(device_counter(g_done_dev, &g_done));
            IFX_CHECK_ARG(peers->world >= 1 && peers->world <= IFX_MAX_PEERS && peers->rank >= 0 &&
                              peers->rank < peers->world && peers->epoch > 0,
                          "ifx_qk_norm_rope_append_peers: bad world / rank / epoch");
            IFX_CHECK_ARG(grid->hw_offset == peers->rank * grid->hw_count,
                          "ifx_qk_norm_rope_append_peers: grid.hw_offset must be rank * hw_count");
            p.paged = 2;
            p.local_only = peers->local_only ? 1 : 0;
            p.sp_world = peers->world;
            p.sp_rank = peers->rank;
            p.epoch = peers->epoch;
            p.done_counter = g_done;
            for (int d = 0; d < peers->world; ++d) {
                IFX_CHECK_ARG(peers->k[d] && peers->v[d] && peers->flags[d], "ifx_qk_norm_rope_append_peers: null peer %d", d);
                p.peer_k[d] = static_cast<__nv_bfloat16*>(peers->k[d]);
                p.peer_v[d] = static_cast<__nv_bfloat16*>(peers->v[d]);
                p.peer_flags[d] = reinterpret_cast<long long*>(peers->flags[d]);
            }
            IFX_CHECK_ARG(peers->k[peers->rank] == kv->k_base && peers->v[peers->rank] == kv->v_base,
                          "ifx_qk_norm_rope_append_peers: own entry must be this rank's cache");
        }
        p.page_tokens = kv->page_tokens;
        p.pl.n = plan->num_pages;
        for (int i = 0; i < plan->num_pages; ++i) p.pl.pages[i] = plan->pages[i];
    } else {
        IFX_CHECK_ARG(k_dst && v_dst, "ifx_qk_norm_rope_append: need kv or k_dst/v_dst");
        p.k_dst = static_cast<__nv_bfloat16*>(k_dst);
        p.v_dst = static_cast<__nv_bfloat16*>(v_dst);
        p.paged = 0;
        p.page_tokens = 1;
        p.pl.n = 0;
    }
    {
        ProfScope prof(peers ? "qk_norm_rope_append_kernel<peers>" : "qk_norm_rope_append_kernel",
                       static_cast<cudaStream_t>(stream));
        const RowLaunch rl = row_launch(rows, C);
        p.rows = rows;
        p.warp_rows = 0;
        if (rl.warp_rows) {
#define IFX_QK_CALL(K)                                                                                               \
    IFX_CUDA_OK(launch_kernel(peers ? qk_norm_rope_append_warp_kernel<K, true> : qk_norm_rope_append_warp_kernel<K, false>, \
                              dim3(rl.grid), dim3(rl.block), 0, static_cast<cudaStream_t>(stream), true, p))
            IFX_WARP_ROW_DISPATCH(C >> 3, IFX_QK_CALL);
#undef IFX_QK_CALL
        } else {
            IFX_CUDA_OK(launch_kernel(peers ? qk_norm_rope_append_kernel<true> : qk_norm_rope_append_kernel<false>,
                                      dim3(rl.grid), dim3(rl.block), 0, static_cast<cudaStream_t>(stream), true, p));
        }
    }
    IFX_LAUNCH_OK("qk_norm_rope_append_kernel");
    return IFX_OK;
}

extern "C" ifx_status ifx_qk_norm_rope_append(const void* qkv, int64_t ld_qkv, const void* norm_q_weight,
                                              const void* norm_k_weight, const double* freqs,
                                              const ifx_rope_grid* grid, void* q_out, int64_t ld_q, ifx_kv* kv_,
                                              const ifx_kv_plan* plan, void* k_dst, void* v_dst, int64_t rows,
                                              int32_t heads, int32_t head_dim, float eps, void* stream) {
    return norm_rope_append_entry(qkv, ld_qkv, norm_q_weight, norm_k_weight, freqs, grid, q_out, ld_q, kv_, plan, k_dst,
                                  v_dst, nullptr, rows, heads, head_dim, eps, stream);
}

extern "C" ifx_status ifx_qk_norm_rope_append_peers(const void* qkv, int64_t ld_qkv, const void* norm_q_weight,
                                                    const void* norm_k_weight, const double* freqs,
                                                    const ifx_rope_grid* grid, void* q_out, int64_t ld_q, ifx_kv* kv_,
                                                    const ifx_kv_plan* plan, const ifx_peer_dst* peers, int64_t rows,
                                                    int32_t heads, int32_t head_dim, float eps, void* stream) {
    IFX_CHECK_ARG(kv_ && plan && peers, "ifx_qk_norm_rope_append_peers: kv, plan and peers are required");
    return norm_rope_append_entry(qkv, ld_qkv, norm_q_weight, norm_k_weight, freqs, grid, q_out, ld_q, kv_, plan, nullptr,
                                  nullptr, peers, rows, heads, head_dim, eps, stream);
}

namespace ifx {
ifx_status fill_peer_push(PeerPushParams& p, const KvImpl* kv, const ifx_kv_plan* plan, const ifx_peer_dst* peers,
                          int32_t frames, int32_t chunk) {
    IFX_CHECK_ARG(kv && plan && peers, "peer push: null pointer");
    IFX_CHECK_ARG(peers->world >= 2 && peers->world <= IFX_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world &&
                      peers->epoch > 0, "peer push: bad world / rank / epoch");
    IFX_CHECK_ARG(frames > 0 && chunk > 0, "peer push: bad geometry");
    const int64_t rows = static_cast<int64_t>(peers->world) * frames * chunk;
    IFX_CHECK_ARG(rows == plan->local_end - plan->local_start && rows == (int64_t)plan->num_pages * kv->page_tokens,
                  "peer push: world*frames*chunk (%lld) does not match the plan", (long long)rows);
    p = PeerPushParams{};
    p.world = peers->world;
    p.rank = peers->rank;
    p.frames = frames;
    p.chunk = chunk;
    p.page_tokens = kv->page_tokens;
    p.C = kv->heads * kv->head_dim;
    p.pl.n = plan->num_pages;
    for (int i = 0; i < plan->num_pages; ++i) p.pl.pages[i] = plan->pages[i];
    for (int d = 0; d < peers->world; ++d) {
        IFX_CHECK_ARG(peers->k[d] && peers->v[d] && peers->flags[d], "peer push: null peer %d", d);
        p.peer_k[d] = static_cast<__nv_bfloat16*>(peers->k[d]);
        p.peer_v[d] = static_cast<__nv_bfloat16*>(peers->v[d]);
        p.peer_flags[d] = reinterpret_cast<long long*>(peers->flags[d]);
    }
    IFX_CHECK_ARG(peers->k[peers->rank] == kv->k_base && peers->v[peers->rank] == kv->v_base,
                  "peer push: own entry must be this rank's cache");
    p.epoch = peers->epoch;
    return IFX_OK;
}
}  // namespace ifx

extern "C" ifx_status ifx_peer_push(ifx_kv* kv_, const ifx_kv_plan* plan, const ifx_peer_dst* peers, int32_t frames,
                                    int32_t chunk, int32_t ctas, void* stream) {
    KvImpl* kv = kv_cast(kv_);
    if (!kv) return set_error(IFX_ERR_HANDLE, "ifx_peer_push: bad kv handle");
    IFX_CHECK_ARG(ctas > 0 && ctas <= 1024, "ifx_peer_push: bad CTA count");
    static unsigned int* g_done_push_dev[64] = {nullptr};   // separate from the append kernel's: the two may overlap
    unsigned int* g_done_push = nullptr;
    IFX_TRY(device_counter(g_done_push_dev, &g_done_push));
    PeerPushParams p;
    ifx_status st = fill_peer_push(p, kv, plan, peers, frames, chunk);
    if (st != IFX_OK) return st;
    p.done_counter = g_done_push;
    {
        ProfScope prof("peer_push_kernel", static_cast<cudaStream_t>(stream));
        IFX_CUDA_OK(launch_kernel(peer_push_kernel, dim3(ctas), dim3(1024), 0, static_cast<cudaStream_t>(stream), true, p));
    }
    IFX_LAUNCH_OK("peer_push_kernel");
    return IFX_OK;
}

extern "C" ifx_status ifx_peer_wait(const int64_t* flags, int32_t world, int64_t epoch, int32_t timeout_ms, void* stream) {
    IFX_CHECK_ARG(flags && world >= 1 && world <= IFX_MAX_PEERS && epoch > 0 && timeout_ms > 0, "ifx_peer_wait: bad argument");
    {
        ProfScope prof("peer_wait_kernel", static_cast<cudaStream_t>(stream));
        IFX_CUDA_OK(launch_kernel(peer_wait_kernel, dim3(1), dim3(32), 0, static_cast<cudaStream_t>(stream), true,
                                  reinterpret_cast<const long long*>(flags), world, static_cast<long long>(epoch),
                                  static_cast<unsigned long long>(timeout_ms) * 1000000ull));
    }
    IFX_LAUNCH_OK("peer_wait_kernel");
    return IFX_OK;
}

namespace ifx {
ifx_status launch_paged_copy(const PagedCopyParams& p, cudaStream_t stream) {
    const int64_t total = p.rows * (p.C >> 3);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    {
        ProfScope prof("paged_copy_kernel", stream);
        paged_copy_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
    }
    IFX_LAUNCH_OK("paged_copy_kernel");
    return IFX_OK;
}
}  // namespace ifx

extern "C" ifx_status ifx_kv_append(ifx_kv* kv_, const ifx_kv_plan* plan, const void* k_src, const void* v_src,
                                    int64_t ld_src, int64_t rows, void* stream) {
    KvImpl* kv = kv_cast(kv_);
    if (!kv) return set_error(IFX_ERR_HANDLE, "ifx_kv_append: bad kv handle");
    IFX_CHECK_ARG(plan && (k_src || v_src), "ifx_kv_append: null pointer");
    const int C = kv->heads * kv->head_dim;
    IFX_CHECK_ARG(ld_src >= C && ld_src % 8 == 0, "ifx_kv_append: bad stride");
    IFX_CHECK_ARG(rows == plan->local_end - plan->local_start && rows == (int64_t)plan->num_pages * kv->page_tokens,
                  "ifx_kv_append: rows do not match plan");
    PagedCopyParams p;
    p.cache_k = static_cast<__nv_bfloat16*>(kv->k_base);
    p.cache_v = static_cast<__nv_bfloat16*>(kv->v_base);
    p.lin_k = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(k_src));
    p.lin_v = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(v_src));
    p.ld_lin = ld_src;
    p.rows = rows;
    p.first_logical = 0;
    p.page_tokens = kv->page_tokens;
    p.C = C;
    p.mode = 0;
    p.linear_row0 = 0;
    p.pl.n = plan->num_pages;
    for (int i = 0; i < plan->num_pages; ++i) p.pl.pages[i] = plan->pages[i];
    return launch_paged_copy(p, static_cast<cudaStream_t>(stream));
}

extern "C" ifx_status ifx_quantize_rows(const void* x, int64_t ldx, void* out, int64_t ldo, float* scales, int64_t rows,
                                        int32_t cols, int32_t kind, void* stream) {
    IFX_CHECK_ARG(x && out && scales, "ifx_quantize_rows: null pointer");
    IFX_CHECK_ARG(rows > 0 && cols > 0 && cols % 16 == 0, "ifx_quantize_rows: cols must be a multiple of 16");
    IFX_CHECK_ARG(ldx >= cols && ldx % 8 == 0 && ldo >= cols && ldo % 16 == 0, "ifx_quantize_rows: bad strides");
    IFX_CHECK_ARG(kind == IFX_Q8_E4M3 || kind == IFX_Q8_INT8, "ifx_quantize_rows: kind must be IFX_Q8_E4M3 or IFX_Q8_INT8");
    int64_t ctas = (rows + 7) / 8;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    if (ctas > cap) ctas = cap;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ProfScope prof("quantize_rows_kernel", st);
        IFX_CUDA_OK(launch_kernel(quantize_rows_kernel, dim3(static_cast<unsigned>(ctas)), dim3(kWarpRowCtaThreads), 0, st,
                                  true, static_cast<const __nv_bfloat16*>(x), ldx, static_cast<uint8_t*>(out), ldo, scales,
                                  rows, cols, kind == IFX_Q8_INT8 ? 1 : 0));
    }
    IFX_LAUNCH_OK("quantize_rows_kernel");
    return IFX_OK;
}

extern "C" ifx_status ifx_ln_modulate_quant(const void* x, void* out, float* scales, const void* ln_weight,
                                            const void* ln_bias, const void* shift, const void* scale,
                                            int64_t mod_frame_stride, int64_t rows, int32_t cols,
                                            int64_t tokens_per_frame, float eps, int32_t kind, void* stream) {
    IFX_CHECK_ARG(x && out && scales, "ifx_ln_modulate_quant: null pointer");
    IFX_CHECK_ARG(rows > 0 && cols > 0 && cols % 16 == 0 && (cols >> 3) <= 256,
                  "ifx_ln_modulate_quant: cols must be a multiple of 16 and <= 2048 (got %d)", cols);
    IFX_CHECK_ARG((ln_weight == nullptr) == (ln_bias == nullptr), "ifx_ln_modulate_quant: weight and bias go together");
    IFX_CHECK_ARG((shift == nullptr) == (scale == nullptr), "ifx_ln_modulate_quant: shift and scale go together");
    IFX_CHECK_ARG(!scale || (tokens_per_frame > 0 && mod_frame_stride % 8 == 0), "ifx_ln_modulate_quant: bad modulation layout");
    IFX_CHECK_ARG(kind == IFX_Q8_E4M3 || kind == IFX_Q8_INT8, "ifx_ln_modulate_quant: bad kind");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t tpf = tokens_per_frame > 0 ? tokens_per_frame : 1;
    const RowLaunch rl = row_launch(rows, cols);
    {
        ProfScope prof("ln_modulate_quant_kernel", st);
#define IFX_LNQ_CALL(K)                                                                                              \
    IFX_CUDA_OK(launch_kernel(ln_modulate_quant_warp_kernel<K>, dim3(rl.grid), dim3(rl.block), 0, st, true,          \
                              static_cast<const __nv_bfloat16*>(x), static_cast<uint8_t*>(out), scales,              \
                              static_cast<const __nv_bfloat16*>(ln_weight), static_cast<const __nv_bfloat16*>(ln_bias), \
                              static_cast<const __nv_bfloat16*>(shift), static_cast<const __nv_bfloat16*>(scale),    \
                              mod_frame_stride, rows, cols, tpf, eps, kind == IFX_Q8_INT8 ? 1 : 0))
        IFX_WARP_ROW_DISPATCH(cols >> 3, IFX_LNQ_CALL);
#undef IFX_LNQ_CALL
    }
    IFX_LAUNCH_OK("ln_modulate_quant_kernel");
    return IFX_OK;
}
