// Row kernels of the MAGI-1 transformer layer (inferix/models/magi/dit/dit_module.py).  All HBM/L2-bound:
// 8/16-byte vector accesses, fp32 statistics, one rounding per reference op.
//
//   magi_qkv_post_kernel        get_q / get_k / get_v / get_xqkv (:902-970): per-head fp32 LayerNorm + partial
//                               non-interleaved rotary for q and k, raw copy of v straight into the KV rows, per-head
//                               bf16 LayerNorm for the cross-attention query — one pass over the fused projection
//   head_layernorm_kernel       k_layernorm_xattn on the caption keys (:968)
//   gate_norm_residual_kernel   bias_modulate_add (:295-313) = range_mod_triton (:205-292) -> fp32 FusedLayerNorm ->
//                               + residual -> bf16
//   silu_mul_kernel             flashinfer.activation.silu_and_mul (:549)
#include "ifx_internal.h"
#include "ifx_ptx.cuh"

namespace ifx {

namespace {

constexpr int kHeadDim = 128;       // every MAGI / Wan model: kv_channels = 128
constexpr int kRowThreadsMax = 256;
constexpr int kVecMax = 8;          // cols <= 256 * 8 * 8

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float bsum(float v, float* scratch /* [32] */) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    v = wsum(v);
    if (nwarps == 1) return v;
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float t = lane < nwarps ? scratch[lane] : 0.f;
    t = wsum(t);
    __syncthreads();
    return t;
}

__device__ __forceinline__ void load4(const __nv_bfloat16* p, float (&f)[4]) {
    const uint2 raw = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
    float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float (&f)[4]) {
    uint2 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    *reinterpret_cast<uint2*>(p) = o;
}
__device__ __forceinline__ void unpack8f(const uint4& raw, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float2 t = __bfloat1622float2(h[e]);
        f[2 * e] = t.x;
        f[2 * e + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8f(const float (&f)[8]) {
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    return o;
}

// LayerNorm of the 128 values a warp holds (4 per lane): two-pass fp32 statistics.
__device__ __forceinline__ void warp_ln128(float (&x)[4], float eps) {
    const float mean = wsum(x[0] + x[1] + x[2] + x[3]) * (1.0f / kHeadDim);
    float d[4], ss = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        d[e] = x[e] - mean;
        ss += d[e] * d[e];
    }
    const float rstd = rsqrtf(wsum(ss) * (1.0f / kHeadDim) + eps);
#pragma unroll
    for (int e = 0; e < 4; ++e) x[e] = d[e] * rstd;
}

struct QkvPostParams {
    const __nv_bfloat16* in;     // [rows, ld]: q | k | v | qx
    int64_t ld;
    int64_t rows;
    int32_t q_heads, kv_heads;
    const float* q_w; const float* q_b;     // fp32 [128]
    const float* k_w; const float* k_b;
    const __nv_bfloat16* x_w; const __nv_bfloat16* x_b;   // bf16 [128] (q_layernorm_xattn)
    const float* rope;           // [rows, ld_rope]: sin[half] | cos[half]
    int64_t ld_rope;
    int32_t half;                // rotary_dim / 2 (multiple of 4, <= 64)
    float eps;
    __nv_bfloat16* q_out; int64_t ld_q;
    __nv_bfloat16* k_dst; __nv_bfloat16* v_dst; int64_t ld_kv;
    __nv_bfloat16* x_out; int64_t ld_x;
    // context parallel (Ulysses): heads are written in destination-rank groups, group g at base + g * group_stride,
    // so the all-to-all send buffer "(cp seq) hn hd" (context_parallel.py:397) needs no rearrange pass
    int32_t q_group_heads, kv_group_heads;
    int64_t q_group_stride, kv_group_stride;
};

// one warp per (token, head slot); lane l owns dims [4l, 4l+4)
__global__ void __launch_bounds__(256) magi_qkv_post_kernel(const QkvPostParams p) {
    const int slots = 2 * p.q_heads + 2 * p.kv_heads;
    const int64_t gw = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= p.rows * slots) return;
    const int64_t row = gw / slots;
    const int slot = static_cast<int>(gw - row * slots);
    const int lane = threadIdx.x & 31;
    const __nv_bfloat16* src = p.in + row * p.ld + static_cast<int64_t>(slot) * kHeadDim + lane * 4;
    float x[4];
    load4(src, x);
    const int kq = p.q_heads, kk = kq + p.kv_heads, kv = kk + p.kv_heads;
    if (slot >= kk && slot < kv) {                         // value: raw copy into the KV rows (get_v :936-938)
        const int h = slot - kk;
        store4(p.v_dst + (h / p.kv_group_heads) * p.kv_group_stride + row * p.ld_kv +
                   static_cast<int64_t>(h % p.kv_group_heads) * kHeadDim + lane * 4, x);
        return;
    }
    if (slot >= kv) {                                      // cross-attention query: bf16 LayerNorm (:958)
        warp_ln128(x, p.eps);
        float w[4], b[4];
        load4(p.x_w + lane * 4, w);
        load4(p.x_b + lane * 4, b);
#pragma unroll
        for (int e = 0; e < 4; ++e) x[e] = x[e] * w[e] + b[e];
        store4(p.x_out + row * p.ld_x + static_cast<int64_t>(slot - kv) * kHeadDim + lane * 4, x);
        return;
    }
    // q / k: fp32 LayerNorm (fp32 affine) -> rotary over dims [0, 2*half) pairing j with j + half -> bf16
    const bool is_q = slot < kq;
    warp_ln128(x, p.eps);
    {
        const float4 w = *reinterpret_cast<const float4*>((is_q ? p.q_w : p.k_w) + lane * 4);
        const float4 b = *reinterpret_cast<const float4*>((is_q ? p.q_b : p.k_b) + lane * 4);
        x[0] = x[0] * w.x + b.x; x[1] = x[1] * w.y + b.y; x[2] = x[2] * w.z + b.z; x[3] = x[3] * w.w + b.w;
    }
    const int hl = p.half >> 2;                            // lanes per rotary half
    const bool lo = lane < hl, hi = lane >= hl && lane < 2 * hl;
    const int partner = lo ? lane + hl : (hi ? lane - hl : lane);
    float y[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) y[e] = __shfl_sync(0xffffffffu, x[e], partner);
    if (lo || hi) {
        const int j = (lo ? lane : lane - hl) * 4;
        const float* rp = p.rope + row * p.ld_rope;
        const float4 sn = *reinterpret_cast<const float4*>(rp + j);
        const float4 cs = *reinterpret_cast<const float4*>(rp + p.half + j);
        const float s[4] = {sn.x, sn.y, sn.z, sn.w}, c[4] = {cs.x, cs.y, cs.z, cs.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            // x1' = x1 cos - x2 sin ; x2' = x2 cos + x1 sin   (separately rounded products, as the torch statement)
            const float a = __fmul_rn(x[e], c[e]);
            const float b = __fmul_rn(y[e], s[e]);
            x[e] = lo ? __fsub_rn(a, b) : __fadd_rn(a, b);
        }
    }
    if (is_q) {
        store4(p.q_out + (slot / p.q_group_heads) * p.q_group_stride + row * p.ld_q +
                   static_cast<int64_t>(slot % p.q_group_heads) * kHeadDim + lane * 4, x);
    } else {
        const int h = slot - kq;
        store4(p.k_dst + (h / p.kv_group_heads) * p.kv_group_stride + row * p.ld_kv +
                   static_cast<int64_t>(h % p.kv_group_heads) * kHeadDim + lane * 4, x);
    }
}

// per-head LayerNorm with bf16 affine on [rows, heads, 128] (row stride ldx / ldo), one warp per (row, head)
__global__ void __launch_bounds__(256)
head_layernorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out, int64_t ldo,
                      int64_t rows, int heads, const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ b,
                      float eps) {
    const int64_t gw = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= rows * heads) return;
    const int64_t row = gw / heads;
    const int head = static_cast<int>(gw - row * heads);
    const int lane = threadIdx.x & 31;
    float v[4], wf[4], bf[4];
    load4(x + row * ldx + static_cast<int64_t>(head) * kHeadDim + lane * 4, v);
    warp_ln128(v, eps);
    load4(w + lane * 4, wf);
    load4(b + lane * 4, bf);
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = v[e] * wf[e] + bf[e];
    store4(out + row * ldo + static_cast<int64_t>(head) * kHeadDim + lane * 4, v);
}

// out[row] = bf16( LN_fp32( x[row] * gate[map[row]] ) * w + b + residual[row] );  x is bf16 or (kXF32) fp32
template <bool kXF32>
__global__ void __launch_bounds__(kRowThreadsMax)
gate_norm_residual_kernel(const void* __restrict__ x_, int64_t ldx, const __nv_bfloat16* __restrict__ gate,
                          int64_t gate_stride, const int32_t* __restrict__ row_map, const float* __restrict__ nw,
                          const float* __restrict__ nb, const __nv_bfloat16* residual, int64_t ldr,
                          __nv_bfloat16* out, int64_t ldo, int cols, float eps) {
    __shared__ float scratch[32];
    const int64_t row = blockIdx.x;
    const int nvec = cols >> 3;
    const uint4* xr = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(x_) + (kXF32 ? 0 : row * ldx));
    const float4* xf = reinterpret_cast<const float4*>(static_cast<const float*>(x_) + (kXF32 ? row * ldx : 0));
    const uint4* gr = reinterpret_cast<const uint4*>(gate + static_cast<int64_t>(__ldg(row_map + row)) * gate_stride);
    float v[kVecMax][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kVecMax; ++i) {
        const int vi = threadIdx.x + i * blockDim.x;
        if (vi < nvec) {
            float g[8];
            if (kXF32) {
                const float4 a = xf[2 * vi], b = xf[2 * vi + 1];
                v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
                v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
            } else {
                unpack8f(xr[vi], v[i]);
            }
            unpack8f(__ldg(gr + vi), g);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                v[i][e] *= g[e];             // x.float() * gate.float()  (exact when x is bf16)
                s += v[i][e];
            }
        }
    }
    const float mean = bsum(s, scratch) / cols;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kVecMax; ++i) {
        const int vi = threadIdx.x + i * blockDim.x;
        if (vi < nvec)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float d = v[i][e] - mean;
                ss += d * d;
            }
    }
    const float rstd = rsqrtf(bsum(ss, scratch) / cols + eps);
    const uint4* rr = reinterpret_cast<const uint4*>(residual + row * ldr);
    uint4* orow = reinterpret_cast<uint4*>(out + row * ldo);
#pragma unroll
    for (int i = 0; i < kVecMax; ++i) {
        const int vi = threadIdx.x + i * blockDim.x;
        if (vi < nvec) {
            float r[8], y[8];
            unpack8f(rr[vi], r);
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(nw) + 2 * vi);
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(nw) + 2 * vi + 1);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(nb) + 2 * vi);
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(nb) + 2 * vi + 1);
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] = ((v[i][e] - mean) * rstd * w[e] + b[e]) + r[e];
            orow[vi] = pack8f(y);
        }
    }
}

// out[r, j] = bf16( silu(x[r, j]) * x[r, cols_out + j] ), grid-stride over 8-wide vectors
__global__ void __launch_bounds__(256)
silu_mul_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out, int64_t ldo,
                int64_t rows, int cols_out) {
    const int nvec = cols_out >> 3;
    const int64_t total = rows * nvec;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / nvec;
        const int c = static_cast<int>(i - r * nvec) * 8;
        float a[8], b[8], y[8];
        unpack8f(*reinterpret_cast<const uint4*>(x + r * ldx + c), a);
        unpack8f(*reinterpret_cast<const uint4*>(x + r * ldx + cols_out + c), b);
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = (a[e] / (1.0f + expf(-a[e]))) * b[e];
        *reinterpret_cast<uint4*>(out + r * ldo + c) = pack8f(y);
    }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline bool al8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }

int row_threads_for(int cols) {
    const int nvec = cols >> 3;
    int t = ((nvec + kVecMax - 1) / kVecMax + 31) / 32 * 32;
    if (nvec <= 32 * kVecMax) t = 32;          // one warp keeps up to 8 x 16 B loads in flight; shuffle-only statistics
    return t > kRowThreadsMax ? kRowThreadsMax : t;
}

}  // namespace
}  // namespace ifx

using namespace ifx;

extern "C" ifx_status ifx_magi_qkv_post(const void* qkvx, int64_t ld, int64_t rows, int32_t q_heads, int32_t kv_heads,
                                        int32_t head_dim, const float* q_ln_w, const float* q_ln_b,
                                        const float* k_ln_w, const float* k_ln_b, const void* qx_ln_w,
                                        const void* qx_ln_b, const float* rope, int64_t ld_rope, int32_t rotary_half,
                                        float eps, void* q_out, int64_t ld_q, int32_t q_group_heads,
                                        int64_t q_group_stride, void* k_dst, void* v_dst, int64_t ld_kv,
                                        int32_t kv_group_heads, int64_t kv_group_stride, void* qx_out, int64_t ld_qx,
                                        void* stream) {
    IFX_CHECK_ARG(qkvx && q_ln_w && q_ln_b && k_ln_w && k_ln_b && qx_ln_w && qx_ln_b && rope && q_out && k_dst && v_dst &&
                      qx_out, "ifx_magi_qkv_post: null pointer");
    if (head_dim != kHeadDim)
        return set_error(IFX_ERR_UNSUPPORTED, "ifx_magi_qkv_post: head_dim %d (built for 128)", head_dim);
    IFX_CHECK_ARG(rows > 0 && q_heads > 0 && kv_heads > 0, "ifx_magi_qkv_post: bad geometry");
    const int64_t width = static_cast<int64_t>(2 * q_heads + 2 * kv_heads) * kHeadDim;
    IFX_CHECK_ARG(q_group_heads > 0 && q_heads % q_group_heads == 0 && kv_group_heads > 0 &&
                      kv_heads % kv_group_heads == 0 && q_group_stride % 4 == 0 && kv_group_stride % 4 == 0 &&
                      q_group_stride >= 0 && kv_group_stride >= 0,
                  "ifx_magi_qkv_post: head groups must divide the head counts");
    IFX_CHECK_ARG(ld >= width && ld % 4 == 0 && ld_q >= q_group_heads * kHeadDim && ld_q % 4 == 0 &&
                      ld_qx >= q_heads * kHeadDim && ld_qx % 4 == 0 && ld_kv >= kv_group_heads * kHeadDim && ld_kv % 4 == 0,
                  "ifx_magi_qkv_post: bad strides");
    IFX_CHECK_ARG(rotary_half > 0 && rotary_half % 4 == 0 && rotary_half <= 64 && ld_rope >= 2 * rotary_half &&
                      ld_rope % 4 == 0, "ifx_magi_qkv_post: rotary_half must be a multiple of 4 in (0, 64]");
    IFX_CHECK_ARG(al8(qkvx) && al8(q_out) && al8(k_dst) && al8(v_dst) && al8(qx_out) && al8(qx_ln_w) && al8(qx_ln_b) &&
                      al16(q_ln_w) && al16(q_ln_b) && al16(k_ln_w) && al16(k_ln_b) && al16(rope) && (rotary_half * 4) % 16 == 0,
                  "ifx_magi_qkv_post: misaligned pointer");
    QkvPostParams p;
    p.in = static_cast<const __nv_bfloat16*>(qkvx);
    p.ld = ld;
    p.rows = rows;
    p.q_heads = q_heads;
    p.kv_heads = kv_heads;
    p.q_w = q_ln_w; p.q_b = q_ln_b; p.k_w = k_ln_w; p.k_b = k_ln_b;
    p.x_w = static_cast<const __nv_bfloat16*>(qx_ln_w);
    p.x_b = static_cast<const __nv_bfloat16*>(qx_ln_b);
    p.rope = rope;
    p.ld_rope = ld_rope;
    p.half = rotary_half;
    p.eps = eps;
    p.q_out = static_cast<__nv_bfloat16*>(q_out); p.ld_q = ld_q;
    p.k_dst = static_cast<__nv_bfloat16*>(k_dst); p.v_dst = static_cast<__nv_bfloat16*>(v_dst); p.ld_kv = ld_kv;
    p.x_out = static_cast<__nv_bfloat16*>(qx_out); p.ld_x = ld_qx;
    p.q_group_heads = q_group_heads; p.q_group_stride = q_group_stride;
    p.kv_group_heads = kv_group_heads; p.kv_group_stride = kv_group_stride;
    const int64_t warps = rows * (2 * q_heads + 2 * kv_heads);
    const int64_t blocks = (warps + 7) / 8;
    IFX_CHECK_ARG(blocks <= 0x7fffffffLL, "ifx_magi_qkv_post: too many rows");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ProfScope prof("magi_qkv_post_kernel", st);
        magi_qkv_post_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(p);
    }
    IFX_LAUNCH_OK("magi_qkv_post_kernel");
    return IFX_OK;
}

extern "C" ifx_status ifx_head_layernorm(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows,
                                         int32_t heads, int32_t head_dim, const void* weight, const void* bias,
                                         float eps, void* stream) {
    IFX_CHECK_ARG(x && out && weight && bias, "ifx_head_layernorm: null pointer");
    if (head_dim != kHeadDim)
        return set_error(IFX_ERR_UNSUPPORTED, "ifx_head_layernorm: head_dim %d (built for 128)", head_dim);
    IFX_CHECK_ARG(rows > 0 && heads > 0 && ldx >= heads * kHeadDim && ldo >= heads * kHeadDim && ldx % 4 == 0 &&
                      ldo % 4 == 0, "ifx_head_layernorm: bad geometry");
    IFX_CHECK_ARG(al8(x) && al8(out) && al8(weight) && al8(bias), "ifx_head_layernorm: misaligned pointer");
    const int64_t blocks = (rows * heads + 7) / 8;
    IFX_CHECK_ARG(blocks <= 0x7fffffffLL, "ifx_head_layernorm: too many rows");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ProfScope prof("head_layernorm_kernel", st);
        head_layernorm_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out), ldo, rows, heads,
            static_cast<const __nv_bfloat16*>(weight), static_cast<const __nv_bfloat16*>(bias), eps);
    }
    IFX_LAUNCH_OK("head_layernorm_kernel");
    return IFX_OK;
}

extern "C" ifx_status ifx_gate_norm_residual(const void* x, int64_t ldx, int32_t x_is_f32, const void* gate,
                                             int64_t gate_stride, int32_t num_gates, const int32_t* row_map,
                                             const float* norm_w, const float* norm_b, const void* residual,
                                             int64_t ldr, void* out, int64_t ldo, int64_t rows, int32_t cols, float eps,
                                             void* stream) {
    IFX_CHECK_ARG(x && gate && row_map && norm_w && norm_b && residual && out, "ifx_gate_norm_residual: null pointer");
    IFX_CHECK_ARG(rows > 0 && cols > 0 && cols % 8 == 0 && cols <= kRowThreadsMax * kVecMax * 8,
                  "ifx_gate_norm_residual: cols must be a multiple of 8 and <= %d (got %d)", kRowThreadsMax * kVecMax * 8,
                  cols);
    IFX_CHECK_ARG(num_gates > 0 && gate_stride >= cols && gate_stride % 8 == 0 && ldx >= cols && ldx % 8 == 0 &&
                      ldr >= cols && ldr % 8 == 0 && ldo >= cols && ldo % 8 == 0, "ifx_gate_norm_residual: bad strides");
    IFX_CHECK_ARG(al16(x) && al16(gate) && al16(norm_w) && al16(norm_b) && al16(residual) && al16(out),
                  "ifx_gate_norm_residual: pointers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ProfScope prof("gate_norm_residual_kernel", st);
        if (x_is_f32)
            gate_norm_residual_kernel<true><<<static_cast<unsigned>(rows), row_threads_for(cols), 0, st>>>(
                x, ldx, static_cast<const __nv_bfloat16*>(gate), gate_stride, row_map, norm_w, norm_b,
                static_cast<const __nv_bfloat16*>(residual), ldr, static_cast<__nv_bfloat16*>(out), ldo, cols, eps);
        else
            gate_norm_residual_kernel<false><<<static_cast<unsigned>(rows), row_threads_for(cols), 0, st>>>(
                x, ldx, static_cast<const __nv_bfloat16*>(gate), gate_stride, row_map, norm_w, norm_b,
                static_cast<const __nv_bfloat16*>(residual), ldr, static_cast<__nv_bfloat16*>(out), ldo, cols, eps);
    }
    IFX_LAUNCH_OK("gate_norm_residual_kernel");
    return IFX_OK;
}

extern "C" ifx_status ifx_silu_mul(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int32_t cols_out,
                                   void* stream) {
    IFX_CHECK_ARG(x && out, "ifx_silu_mul: null pointer");
    IFX_CHECK_ARG(rows > 0 && cols_out > 0 && cols_out % 8 == 0 && ldx >= 2 * cols_out && ldx % 8 == 0 &&
                      ldo >= cols_out && ldo % 8 == 0, "ifx_silu_mul: bad geometry");
    IFX_CHECK_ARG(al16(x) && al16(out), "ifx_silu_mul: pointers must be 16-byte aligned");
    const int64_t total = rows * (cols_out >> 3);
    const int64_t want = (total + 255) / 256;
    const int blocks = static_cast<int>(want < static_cast<int64_t>(sm_count()) * 16 ? want
                                                                                      : static_cast<int64_t>(sm_count()) * 16);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ProfScope prof("silu_mul_kernel", st);
        silu_mul_kernel<<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), ldx,
                                                static_cast<__nv_bfloat16*>(out), ldo, rows, cols_out);
    }
    IFX_LAUNCH_OK("silu_mul_kernel");
    return IFX_OK;
}
