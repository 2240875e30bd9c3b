// Prologue and epilogue of one DiT forward: everything between the latent tensor and the first
// block, and between the last block and the denoised latent, as a handful of small kernels instead of ~45 eager ops.
//
//   patchify_kernel         latents [C_in, F, H, W] -> token rows [rows, C_in * pt * ph * pw] (this rank's hw slice of
//                           every frame), the A operand of the patch-embedding GEMM (Conv3d with kernel == stride,
//                           causal_model.py:916-921; under sequence parallelism the scatter of :939-942 is folded in)
//   sinusoid_kernel         sinusoidal_embedding_1d in fp64 (wan_base/components.py:11-31) -> bf16
//   linear_small_kernel     out[m, :] = bf16(act_in(x[m, :]) @ W^T + b) for a handful of rows (time MLP, :922-936):
//                           one warp per output column, weights streamed once; optional SiLU on the input; optional
//                           modulation-table epilogue writing every layer's `modulation + e0` (:412) in one pass
//   unpatchify_x0_kernel    head tokens [S, ph*pw*C_out] -> flow [F, C, H, W] (unpatchify, :1196-1219) and
//                           x0 = x_t - sigma_t * flow in fp64 (wrapper.py:259-283), sigma looked up by argmin over the
//                           scheduler's timestep table as the reference does
//   add_noise_kernel        (1 - sigma) * x0 + sigma * noise in fp32 -> bf16 (flow_match.py:159-176)
#include "ifx_internal.h"
#include "ifx_ptx.cuh"

namespace ifx {

__global__ void __launch_bounds__(256)
patchify_kernel(const __nv_bfloat16* __restrict__ x, int64_t sc, int64_t sf, int64_t sh, int64_t sw, int C_in, int pt,
                int ph, int pw, int frames, int gh, int gw, int hw_offset, int hw_count, __nv_bfloat16* __restrict__ out) {
    griddep_launch();
    griddep_wait();
    const int K = C_in * pt * ph * pw;
    const int64_t total = static_cast<int64_t>(frames) * hw_count * K;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % K);
        const int64_t row = i / K;
        const int f = static_cast<int>(row / hw_count);
        const int hw = hw_offset + static_cast<int>(row % hw_count);
        const int h = hw / gw, w = hw % gw;
        // k = ((c * pt + dt) * ph + dy) * pw + dx : the flattening of Conv3d's weight [out, C_in, pt, ph, pw]
        const int dx = k % pw, dy = (k / pw) % ph, dt = (k / (pw * ph)) % pt, c = k / (pw * ph * pt);
        out[i] = x[c * sc + (f * pt + dt) * sf + (h * ph + dy) * sh + (w * pw + dx) * sw];
    }
}

__global__ void sinusoid_kernel(const double* __restrict__ t, int n, int dim, __nv_bfloat16* __restrict__ out) {
    griddep_launch();
    griddep_wait();
    const int half = dim / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * half; i += gridDim.x * blockDim.x) {
        const int r = i / half, j = i % half;
        // position * 10000^(-j / half), all in fp64 like the reference
        const double ang = t[r] * pow(10000.0, -static_cast<double>(j) / static_cast<double>(half));
        // torch casts double -> bf16 through float
        out[r * dim + j] = __float2bfloat16_rn(static_cast<float>(cos(ang)));
        out[r * dim + half + j] = __float2bfloat16_rn(static_cast<float>(sin(ang)));
    }
}

constexpr int kSmallRows = 8;

// one warp per output column n: dot products of W[n, :] with up to kSmallRows input rows (fp32 accumulate, one bf16
// rounding after the bias, as nn.Linear in bf16).  silu_in: the input is SiLU(x) rounded to bf16 (nn.SiLU before the
// Linear).  mod_table != nullptr: out is [layers, M, N] and receives bf16(mod_table[l, n] + bf16(acc + b[n])).
__global__ void __launch_bounds__(256)
linear_small_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ w, int64_t ldw,
                    const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out, int64_t ldo, int M, int N,
                    int K, int silu_in, const __nv_bfloat16* __restrict__ mod_table, int layers, int64_t mod_stride,
                    int64_t out_layer_stride) {
    griddep_launch();
    griddep_wait();
    extern __shared__ __nv_bfloat16 xs[];      // [M][K] activated input
    for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
        float v = __bfloat162float(x[(i / K) * ldx + (i % K)]);
        if (silu_in) v = bf16_round(v / (1.0f + expf(-v)));
        xs[i] = __float2bfloat16_rn(v);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps) {
        float acc[kSmallRows];
#pragma unroll
        for (int m = 0; m < kSmallRows; ++m) acc[m] = 0.f;
        const uint4* wr = reinterpret_cast<const uint4*>(w + n * ldw);
        for (int kv = lane; kv < (K >> 3); kv += 32) {
            const uint4 wraw = __ldg(wr + kv);
            const __nv_bfloat162* w2 = reinterpret_cast<const __nv_bfloat162*>(&wraw);
            float wf[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(w2[e]);
                wf[2 * e] = f.x;
                wf[2 * e + 1] = f.y;
            }
#pragma unroll
            for (int m = 0; m < kSmallRows; ++m)
                if (m < M) {
                    const uint4 xraw = *reinterpret_cast<const uint4*>(xs + m * K + kv * 8);
                    const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&xraw);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = __bfloat1622float2(x2[e]);
                        acc[m] = fmaf(f.x, wf[2 * e], acc[m]);
                        acc[m] = fmaf(f.y, wf[2 * e + 1], acc[m]);
                    }
                }
        }
#pragma unroll
        for (int m = 0; m < kSmallRows; ++m)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
        if (lane == 0) {
            const float b = bias ? __bfloat162float(bias[n]) : 0.f;
            for (int m = 0; m < M; ++m) {
                const float y = bf16_round(acc[m] + b);
                if (mod_table == nullptr) {
                    out[m * ldo + n] = __float2bfloat16_rn(y);
                } else {
                    for (int l = 0; l < layers; ++l)
                        out[l * out_layer_stride + m * ldo + n] =
                            __float2bfloat16_rn(__bfloat162float(mod_table[l * mod_stride + n]) + y);
                }
            }
        }
    }
}

struct SigmaTable {
    const float* timesteps;   // [n] scheduler.timesteps (fp32)
    const float* sigmas;      // [n] scheduler.sigmas (fp32)
    int n;
};

// sigma of frame f: argmin_i |timesteps[i] - t[f]| (first minimum, as torch.argmin), one warp.  The reference takes
// the difference in fp64 in the flow -> x0 conversion (wrapper.py:270-276: everything .double()) and in fp32 in
// add_noise (flow_match.py:166-170: fp32 table minus an int64 tensor promotes to fp32).
__device__ __forceinline__ double frame_sigma(const SigmaTable& tab, double t, int lane, bool fp32_diff) {
    double best = 1e300;
    int best_i = 0x7fffffff;
    for (int i = lane; i < tab.n; i += 32) {
        const double d = fp32_diff ? static_cast<double>(fabsf(tab.timesteps[i] - static_cast<float>(t)))
                                   : fabs(static_cast<double>(tab.timesteps[i]) - t);
        if (d < best) {
            best = d;
            best_i = i;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ob < best || (ob == best && oi < best_i)) {
            best = ob;
            best_i = oi;
        }
    }
    return static_cast<double>(tab.sigmas[best_i]);
}

// grid.y = frame.  y: head tokens [F * gh * gw, ph * pw * C] ('f h w (p q r c)' with p = 1); xt / flow / x0: frame f at
// base + f * frame stride, element (c, Y, X) at c * sc + Y * sh + X * sw (flow and x0 are contiguous [F, C, H, W]).
__global__ void __launch_bounds__(256)
unpatchify_x0_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ xt, int64_t xf, int64_t xc,
                     int64_t xh, int64_t xw, const double* __restrict__ t, SigmaTable tab, int C, int gh, int gw, int ph,
                     int pw, __nv_bfloat16* __restrict__ flow, __nv_bfloat16* __restrict__ x0) {
    griddep_launch();
    griddep_wait();
    __shared__ double sigma_s;
    const int f = blockIdx.y;
    if (threadIdx.x < 32) {
        const double s = frame_sigma(tab, t[f], threadIdx.x, false);
        if (threadIdx.x == 0) sigma_s = s;
    }
    __syncthreads();
    const double sigma = sigma_s;
    const int H = gh * ph, W = gw * pw;
    const int64_t per_frame = static_cast<int64_t>(C) * H * W;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < per_frame;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        // i enumerates the TOKEN layout (coalesced reads of y): i = ((h * gw + w) * ph * pw + (q * pw + r)) * C + c
        const int c = static_cast<int>(i % C);
        const int qr = static_cast<int>((i / C) % (ph * pw));
        const int64_t tok = i / (static_cast<int64_t>(C) * ph * pw);
        const int h = static_cast<int>(tok / gw), w = static_cast<int>(tok % gw);
        const int Y = h * ph + qr / pw, X = w * pw + qr % pw;
        const __nv_bfloat16 fl = y[(static_cast<int64_t>(f) * gh * gw + tok) * (ph * pw * C) + qr * C + c];
        const int64_t o = (static_cast<int64_t>(f) * C + c) * H * W + static_cast<int64_t>(Y) * W + X;
        if (flow != nullptr) flow[o] = fl;
        const double xv = static_cast<double>(__bfloat162float(xt[f * xf + c * xc + Y * xh + X * xw]));
        // fp64 arithmetic, then torch's double -> float -> bf16 cast chain
        x0[o] = __float2bfloat16_rn(static_cast<float>(xv - sigma * static_cast<double>(__bfloat162float(fl))));
    }
}

// out = bf16( (1 - sigma_f) * x0 + sigma_f * noise ) in fp32 (sigma is the scheduler's fp32 table entry); grid.y = frame
__global__ void __launch_bounds__(256)
add_noise_kernel(const __nv_bfloat16* __restrict__ x0, const __nv_bfloat16* __restrict__ noise,
                 const double* __restrict__ t, SigmaTable tab, int64_t per_frame, __nv_bfloat16* __restrict__ out) {
    griddep_launch();
    griddep_wait();
    __shared__ float sigma_s;
    const int f = blockIdx.y;
    if (threadIdx.x < 32) {
        const double s = frame_sigma(tab, t[f], threadIdx.x, true);
        if (threadIdx.x == 0) sigma_s = static_cast<float>(s);
    }
    __syncthreads();
    const float sigma = sigma_s;
    const float one_minus = 1.0f - sigma;
    const int64_t base = f * per_frame;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < per_frame;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float a = __fmul_rn(one_minus, __bfloat162float(x0[base + i]));     // no FMA contraction: two eager ops
        const float b = __fmul_rn(sigma, __bfloat162float(noise[base + i]));
        out[base + i] = __float2bfloat16_rn(__fadd_rn(a, b));
    }
}

}  // namespace ifx

using namespace ifx;

extern "C" ifx_status ifx_patchify(const void* x, int64_t stride_c, int64_t stride_f, int64_t stride_h, int64_t stride_w,
                                   int32_t c_in, int32_t pt, int32_t ph, int32_t pw, int32_t frames, int32_t grid_h,
                                   int32_t grid_w, int32_t hw_offset, int32_t hw_count, void* out, void* stream) {
    IFX_CHECK_ARG(x && out, "ifx_patchify: null pointer");
    IFX_CHECK_ARG(c_in > 0 && pt > 0 && ph > 0 && pw > 0 && frames > 0 && grid_h > 0 && grid_w > 0,
                  "ifx_patchify: bad geometry");
    IFX_CHECK_ARG(hw_offset >= 0 && hw_count > 0 && hw_offset + hw_count <= grid_h * grid_w,
                  "ifx_patchify: hw slice [%d, %d) outside the %d x %d grid", hw_offset, hw_offset + hw_count, grid_h, grid_w);
    const int64_t total = static_cast<int64_t>(frames) * hw_count * c_in * pt * ph * pw;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ProfScope prof("patchify_kernel", st);
        IFX_CUDA_OK(launch_kernel(patchify_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, true,
                                  static_cast<const __nv_bfloat16*>(x), stride_c, stride_f, stride_h, stride_w, c_in, pt, ph,
                                  pw, frames, grid_h, grid_w, hw_offset, hw_count, static_cast<__nv_bfloat16*>(out)));
    }
    IFX_LAUNCH_OK("patchify_kernel");
    return IFX_OK;
}

extern "C" ifx_status ifx_sinusoidal_embedding(const double* positions, int32_t n, int32_t dim, void* out, void* stream) {
    IFX_CHECK_ARG(positions && out && n > 0 && dim > 0 && dim % 2 == 0, "ifx_sinusoidal_embedding: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ProfScope prof("sinusoid_kernel", st);
        IFX_CUDA_OK(launch_kernel(sinusoid_kernel, dim3((n * dim / 2 + 127) / 128), dim3(128), 0, st, true, positions, n, dim,
                                  static_cast<__nv_bfloat16*>(out)));
    }
    IFX_LAUNCH_OK("sinusoid_kernel");
    return IFX_OK;
}

extern "C" ifx_status ifx_linear_small(const void* x, int64_t ldx, const void* w, int64_t ldw, const void* bias, void* out,
                                       int64_t ldo, int32_t M, int32_t N, int32_t K, int32_t silu_input,
                                       const void* mod_table, int32_t layers, int64_t mod_layer_stride,
                                       int64_t out_layer_stride, void* stream) {
    IFX_CHECK_ARG(x && w && out, "ifx_linear_small: null pointer");
    IFX_CHECK_ARG(M >= 1 && M <= kSmallRows, "ifx_linear_small: 1..%d rows (got %d); use ifx_gemm_bf16 beyond", kSmallRows, M);
    IFX_CHECK_ARG(N > 0 && K > 0 && K % 8 == 0 && ldw % 8 == 0 && ldx >= K && ldw >= K && ldo >= N,
                  "ifx_linear_small: bad shape / strides");
    IFX_CHECK_ARG((reinterpret_cast<uintptr_t>(w) & 15) == 0, "ifx_linear_small: w must be 16-byte aligned");
    IFX_CHECK_ARG(mod_table == nullptr || (layers > 0 && mod_layer_stride >= N && out_layer_stride >= static_cast<int64_t>(M) * ldo),
                  "ifx_linear_small: bad modulation-table layout");
    const size_t smem = static_cast<size_t>(M) * K * sizeof(__nv_bfloat16);
    IFX_CHECK_ARG(smem <= 48 * 1024, "ifx_linear_small: M * K too large for the staging buffer");
    int blocks = (N + 7) / 8;                    // 8 warps per CTA, one column per warp per pass
    const int cap = sm_count() * 4;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ProfScope prof("linear_small_kernel", st);
        IFX_CUDA_OK(launch_kernel(linear_small_kernel, dim3(blocks), dim3(256), smem, st, true,
                                  static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(w), ldw,
                                  static_cast<const __nv_bfloat16*>(bias), static_cast<__nv_bfloat16*>(out), ldo, M, N, K,
                                  silu_input, static_cast<const __nv_bfloat16*>(mod_table), layers, mod_layer_stride,
                                  out_layer_stride));
    }
    IFX_LAUNCH_OK("linear_small_kernel");
    return IFX_OK;
}

static ifx_status check_table(const float* timesteps, const float* sigmas, int32_t n, const char* who) {
    IFX_CHECK_ARG(timesteps && sigmas && n > 0, "%s: null sigma table", who);
    return IFX_OK;
}

extern "C" ifx_status ifx_unpatchify_x0(const void* head_tokens, const void* xt, int64_t xt_stride_f, int64_t xt_stride_c,
                                        int64_t xt_stride_h, int64_t xt_stride_w, const double* timestep,
                                        const float* table_timesteps, const float* table_sigmas, int32_t table_len,
                                        int32_t frames, int32_t channels, int32_t grid_h, int32_t grid_w, int32_t ph,
                                        int32_t pw, void* flow_out, void* x0_out, void* stream) {
    IFX_CHECK_ARG(head_tokens && xt && timestep && x0_out, "ifx_unpatchify_x0: null pointer");
    IFX_CHECK_ARG(frames > 0 && channels > 0 && grid_h > 0 && grid_w > 0 && ph > 0 && pw > 0, "ifx_unpatchify_x0: bad geometry");
    ifx_status s = check_table(table_timesteps, table_sigmas, table_len, "ifx_unpatchify_x0");
    if (s != IFX_OK) return s;
    SigmaTable tab{table_timesteps, table_sigmas, table_len};
    const int64_t per_frame = static_cast<int64_t>(channels) * grid_h * ph * grid_w * pw;
    unsigned bx = static_cast<unsigned>((per_frame + 255) / 256);
    const unsigned cap = static_cast<unsigned>(sm_count() * 4 / frames + 1);
    if (bx > cap) bx = cap;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ProfScope prof("unpatchify_x0_kernel", st);
        IFX_CUDA_OK(launch_kernel(unpatchify_x0_kernel, dim3(bx, frames), dim3(256), 0, st, true,
                                  static_cast<const __nv_bfloat16*>(head_tokens), static_cast<const __nv_bfloat16*>(xt),
                                  xt_stride_f, xt_stride_c, xt_stride_h, xt_stride_w, timestep, tab, channels, grid_h, grid_w,
                                  ph, pw, static_cast<__nv_bfloat16*>(flow_out), static_cast<__nv_bfloat16*>(x0_out)));
    }
    IFX_LAUNCH_OK("unpatchify_x0_kernel");
    return IFX_OK;
}

extern "C" ifx_status ifx_add_noise(const void* x0, const void* noise, const double* timestep, const float* table_timesteps,
                                    const float* table_sigmas, int32_t table_len, int32_t frames, int64_t per_frame,
                                    void* out, void* stream) {
    IFX_CHECK_ARG(x0 && noise && timestep && out && frames > 0 && per_frame > 0, "ifx_add_noise: bad argument");
    ifx_status s = check_table(table_timesteps, table_sigmas, table_len, "ifx_add_noise");
    if (s != IFX_OK) return s;
    SigmaTable tab{table_timesteps, table_sigmas, table_len};
    unsigned bx = static_cast<unsigned>((per_frame + 255) / 256);
    const unsigned cap = static_cast<unsigned>(sm_count() * 4 / frames + 1);
    if (bx > cap) bx = cap;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ProfScope prof("add_noise_kernel", st);
        IFX_CUDA_OK(launch_kernel(add_noise_kernel, dim3(bx, frames), dim3(256), 0, st, true,
                                  static_cast<const __nv_bfloat16*>(x0), static_cast<const __nv_bfloat16*>(noise), timestep, tab,
                                  per_frame, static_cast<__nv_bfloat16*>(out)));
    }
    IFX_LAUNCH_OK("add_noise_kernel");
    return IFX_OK;
}
