// Shared by the 1-CTA (ifx_gemm.cu) and 2-CTA (ifx_gemm2.cu) tcgen05 GEMM kernels: parameters and the fused epilogue.
#pragma once
#include "ifx_internal.h"
#include "ifx_ptx.cuh"

namespace ifx {

struct GemmParams {
    int64_t M;
    int32_t N, K;
    float alpha;  // FP8 path: input_scale * weight_scale applied to the fp32 accumulator before the bias
    // 8-bit operand family (the kFp8 instantiations), dynamic quantisation: per-row activation scales and per-column
    // weight scales replace alpha when non-null; int8 = 1 selects kind::i8 (S8 x S8 -> S32 accumulators in TMEM)
    const float* row_scale;   // [M]
    const float* col_scale;   // [N]
    int32_t int8;
    const __nv_bfloat16* bias;
    __nv_bfloat16* out;
    int64_t ldo;
    const __nv_bfloat16* residual;
    int64_t ldr;
    const __nv_bfloat16* gate;
    int64_t gate_frame_stride;
    int64_t tokens_per_frame;
    int32_t num_m_tiles, num_n_tiles;
};

__device__ __forceinline__ float gelu_tanh_f(float x) {
    // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))), tanh(u) = 1 - 2 / (1 + e^{2u})
    const float kBeta = 0.7978845608028654f;
    const float kKappa = 0.044715f;
    float u = kBeta * (x + kKappa * x * x * x);
    float e = __expf(2.0f * u);
    float t = 1.0f - __fdividef(2.0f, 1.0f + e);
    return 0.5f * x * (1.0f + t);
}

// accumulator bits -> dequantised fp32 value (8-bit operand family)
__device__ __forceinline__ float dequant8(const GemmParams& p, uint32_t bits, float row_s, int col) {
    const float a = p.int8 ? __int2float_rn(static_cast<int>(bits)) : __uint_as_float(bits);
    return a * row_s * (p.col_scale ? __ldg(p.col_scale + col) : p.alpha);
}

// Epilogue of 32 accumulator columns of one output row: (dequant) + bias -> bf16 -> GELU | gate + residual -> bf16,
// one rounding per reference op (causal_model.py:378-379,444,455-456), four 16-byte stores.
template <int kEpi, bool kFp8>
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmParams& p, const uint32_t (&acc)[32], int64_t row, int col0,
                                                    const __nv_bfloat16* gate_row) {
    __nv_bfloat16* optr = p.out + row * p.ldo + col0;
    const float row_s = (kFp8 && p.row_scale != nullptr) ? __ldg(p.row_scale + row) : 1.0f;
    if (kEpi == IFX_EPI_BIAS_F32) {
        // fp32 result, no bf16 rounding of the accumulator: MAGI's output projection runs under
        // torch.autocast(dtype=float32) (dit_module.py:1291-1293) and feeds the fp32 gate / post-norm directly
        float* fo = reinterpret_cast<float*>(p.out) + row * p.ldo + col0;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            if (col0 + v * 4 >= p.N) break;
            float4 o;
            float* oe = reinterpret_cast<float*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float a = kFp8 ? dequant8(p, acc[v * 4 + e], row_s, col0 + v * 4 + e) : __uint_as_float(acc[v * 4 + e]);
                oe[e] = a + (p.bias ? __bfloat162float(p.bias[col0 + v * 4 + e]) : 0.f);
            }
            *reinterpret_cast<float4*>(fo + v * 4) = o;
        }
        return;
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        if (col0 + v * 8 >= p.N) break;
        float bv[8];
        {
            uint4 braw = p.bias ? __ldg(reinterpret_cast<const uint4*>(p.bias + col0 + v * 8))
                                : make_uint4(0, 0, 0, 0);
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&braw);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 f = __bfloat1622float2(b2[e]);
                bv[2 * e] = f.x;
                bv[2 * e + 1] = f.y;
            }
        }
        float val[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float a = kFp8 ? dequant8(p, acc[v * 8 + e], row_s, col0 + v * 8 + e) : __uint_as_float(acc[v * 8 + e]);
            val[e] = bf16_round(a + bv[e]);
        }
        if (kEpi == IFX_EPI_BIAS_GELU) {
#pragma unroll
            for (int e = 0; e < 8; ++e) val[e] = gelu_tanh_f(val[e]);
        }
        if (kEpi == IFX_EPI_BIAS_GELU_ERF) {
#pragma unroll
            for (int e = 0; e < 8; ++e) val[e] = 0.5f * val[e] * (1.0f + erff(val[e] * 0.70710678118654752f));
        }
        if (kEpi == IFX_EPI_BIAS_GATE_RES) {
            if (gate_row != nullptr) {
                uint4 graw = __ldg(reinterpret_cast<const uint4*>(gate_row + col0 + v * 8));
                const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&graw);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float2 f = __bfloat1622float2(g2[e]);
                    val[2 * e] = bf16_round(val[2 * e] * f.x);
                    val[2 * e + 1] = bf16_round(val[2 * e + 1] * f.y);
                }
            }
            uint4 rraw = *reinterpret_cast<const uint4*>(p.residual + row * p.ldr + col0 + v * 8);
            const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rraw);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 f = __bfloat1622float2(r2[e]);
                val[2 * e] = f.x + val[2 * e];
                val[2 * e + 1] = f.y + val[2 * e + 1];
            }
        }
        uint4 o;
        o.x = pack_bf16x2(val[0], val[1]);
        o.y = pack_bf16x2(val[2], val[3]);
        o.z = pack_bf16x2(val[4], val[5]);
        o.w = pack_bf16x2(val[6], val[7]);
        *reinterpret_cast<uint4*>(optr + v * 8) = o;
    }
}

}  // namespace ifx
