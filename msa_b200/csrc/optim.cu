// Weight maintenance kernels (HBM-bound, vectorised): fp32 master -> bf16 operand copies, small transposes,
// and the fused AdamW step with the semantics of transformers' (<= 4.x) AdamW that the reference constructs
// (train.py:76-92: betas (0.9, 0.999), eps 1e-6, correct_bias=True, weight decay applied after the update).
#include "common.cuh"

namespace mmb {

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
    const size_t nvec = n / 8;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (size_t)gridDim.x * 256) {
        const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
        uint4 o;
        o.x = pack_bf16x2(a.x, a.y);
        o.y = pack_bf16x2(a.z, a.w);
        o.z = pack_bf16x2(b.x, b.y);
        o.w = pack_bf16x2(b.z, b.w);
        reinterpret_cast<uint4*>(dst)[i] = o;
    }
    if (blockIdx.x == 0)
        for (size_t i = nvec * 8 + threadIdx.x; i < n; i += 256) dst[i] = __float2bfloat16_rn(src[i]);
}

// dst[c][r] = src[r][c]
__global__ void transpose_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int C) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < C) ? src[(size_t)r * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < C) dst[(size_t)c * R + r] = tile[threadIdx.x][i];
    }
}

// One fused pass: Adam moments, bias-corrected step, decoupled weight decay, bf16 operand refresh.
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             __nv_bfloat16* __restrict__ p_bf16, size_t n, float lr, float beta1, float beta2, float eps, float step_size,
             float weight_decay, float grad_scale) {
    const size_t nvec = n / 4;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (size_t)gridDim.x * 256) {
        float4 P = reinterpret_cast<float4*>(p)[i];
        const float4 G = reinterpret_cast<const float4*>(g)[i];
        float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
        float pp[4] = {P.x, P.y, P.z, P.w}, gg[4] = {G.x, G.y, G.z, G.w}, mm[4] = {M.x, M.y, M.z, M.w},
              vv[4] = {V.x, V.y, V.z, V.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gr = gg[j] * grad_scale;
            mm[j] = mm[j] * beta1 + (1.f - beta1) * gr;
            vv[j] = vv[j] * beta2 + (1.f - beta2) * gr * gr;
            pp[j] -= step_size * mm[j] / (sqrtf(vv[j]) + eps);
            pp[j] -= lr * weight_decay * pp[j];
        }
        reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
        reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
        if (p_bf16) {
            uint2 o;
            o.x = pack_bf16x2(pp[0], pp[1]);
            o.y = pack_bf16x2(pp[2], pp[3]);
            reinterpret_cast<uint2*>(p_bf16)[i] = o;
        }
    }
}

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_cast_bf16(const float* src, void* dst, size_t n, void* stream) {
    MMB_REQUIRE(src && dst && n > 0, "cast_bf16: bad arguments");
    MMB_REQUIRE(((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0, "cast_bf16: pointers must be 16-byte aligned");
    const size_t nvec = n / 8 + 1;
    const int grid = (int)((nvec + 255) / 256 < (size_t)num_sms() * 8 ? (nvec + 255) / 256 : (size_t)num_sms() * 8);
    cast_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
    return check_launch("cast_bf16_kernel");
}

extern "C" int mmb_transpose_f32(const float* src, float* dst, int R, int C, void* stream) {
    MMB_REQUIRE(src && dst && R > 0 && C > 0, "transpose_f32: bad arguments");
    dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
    transpose_f32_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, dst, R, C);
    return check_launch("transpose_f32_kernel");
}

extern "C" int mmb_adamw(const mmb_adamw_args* a, void* stream) {
    MMB_REQUIRE(a && a->p && a->g && a->m && a->v, "adamw: null pointer");
    MMB_REQUIRE(a->n > 0 && a->n % 4 == 0, "adamw: n=%zu must be a positive multiple of 4", a->n);
    MMB_REQUIRE(a->step >= 1, "adamw: step must be >= 1");
    float step_size = a->lr;
    if (a->correct_bias) {
        const double bc1 = 1.0 - pow((double)a->beta1, (double)a->step);
        const double bc2 = 1.0 - pow((double)a->beta2, (double)a->step);
        step_size = (float)((double)a->lr * sqrt(bc2) / bc1);
    }
    const size_t nvec = a->n / 4;
    const int grid = (int)((nvec + 255) / 256 < (size_t)num_sms() * 8 ? (nvec + 255) / 256 : (size_t)num_sms() * 8);
    adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a->p, a->g, a->m, a->v, (__nv_bfloat16*)a->p_bf16, a->n, a->lr,
                                                        a->beta1, a->beta2, a->eps, step_size, a->weight_decay,
                                                        a->grad_scale);
    return check_launch("adamw_kernel");
}
