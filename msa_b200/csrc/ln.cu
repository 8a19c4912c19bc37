// HBM-bound row kernels of the encoder block: dropout + residual + LayerNorm (forward / backward) and the
// bf16 column-sum used for bias gradients.  One warp per row, 16-byte vector accesses, warp-shuffle statistics.
//
// Replaces the ATen kernel chains behind
//   BertSelfOutput.forward  modeling_bert.py:295-297   LayerNorm(dropout(dense(x)) + input)
//   BertOutput.forward      modeling_bert.py:353-355
//   BertPredictionHeadTransform LayerNorm :484 (res == NULL, p == 0)
// (the dense bias is already added by the GEMM epilogue).
//
// Precision: the residual stream (LayerNorm outputs that feed the next residual add) and its gradient are kept
// in fp32 next to the bf16 copy that the GEMMs read — the same split torch.autocast(bf16) makes in the reference
// (LayerNorm runs and returns fp32, Linear casts its input to bf16).  Measured on the CPU emulation of this
// path at 12 layers: bf16 residual stream 1.4e-2 max-rel error on the encoder output, fp32 stream 5.7e-3
// (reference autocast: 6.0e-3).
#include "common.cuh"
#include "rowops.cuh"

namespace mmb {

constexpr int kLnWarps = 8;

template <int NCH, bool kYF32 = false>
__global__ void __launch_bounds__(kLnWarps * 32)
drln_fwd_kernel(const void* __restrict__ y_, const float* __restrict__ res,
                const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                float* __restrict__ out_f32, float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int H, float eps, uint32_t thresh,
                float inv_keep, uint64_t seed, uint32_t stream) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = gridDim.x * kLnWarps;
    for (int row = blockIdx.x * kLnWarps + warp; row < M; row += nw) {
        RowF<NCH> z;
        if (kYF32) row_load_f32(z, reinterpret_cast<const float*>(y_) + (size_t)row * H, H, lane);
        else row_load_bf16(z, reinterpret_cast<const __nv_bfloat16*>(y_) + (size_t)row * H, H, lane);
        row_dropout(z, H, lane, seed, stream, (uint64_t)row, thresh, inv_keep);
        if (res != nullptr) {
            RowF<NCH> r;
            row_load_f32(r, res + (size_t)row * H, H, lane);
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int i = 0; i < 8; ++i) z.v[c][i] += r.v[c][i];
        }
        float mean, rstd;
        row_stats(z, H, lane, eps, mean, rstd);
        row_affine(z, H, lane, mean, rstd, gamma, beta);
        if (out != nullptr) row_store_bf16(z, out + (size_t)row * H, H, lane);
        if (out_f32 != nullptr) row_store_f32(z, out_f32 + (size_t)row * H, H, lane);
        if (lane == 0 && mean_out != nullptr) {
            mean_out[row] = mean;
            rstd_out[row] = rstd;
        }
    }
}

// Column partials (dgamma, dbeta, dbias) live in per-warp private shared memory instead of registers: that drops
// the kernel from 151 to <100 registers, doubles the resident warps per SM and with them the loads in flight
// (the register version was latency-bound at ~2.5x its HBM time).
template <int NCH>
__device__ __forceinline__ void smem_accumulate(float* __restrict__ acc, const RowF<NCH>& v, int H, int lane) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int e = (c * 32 + lane) * 8;
        if (e < H) {
            float4* p = reinterpret_cast<float4*>(acc + e);
            float4 a = p[0], b = p[1];
            a.x += v.v[c][0]; a.y += v.v[c][1]; a.z += v.v[c][2]; a.w += v.v[c][3];
            b.x += v.v[c][4]; b.y += v.v[c][5]; b.z += v.v[c][6]; b.w += v.v[c][7];
            p[0] = a;
            p[1] = b;
        }
    }
}

template <int NCH>
__global__ void __launch_bounds__(kLnWarps * 32, 2)
drln_bwd_kernel(const __nv_bfloat16* __restrict__ g1, const float* __restrict__ g2,
                const __nv_bfloat16* __restrict__ y, const float* __restrict__ res,
                const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ gamma,
                __nv_bfloat16* __restrict__ d_y, float* __restrict__ d_res, float* __restrict__ dgamma,
                float* __restrict__ dbeta, float* __restrict__ dbias, const __nv_bfloat16* __restrict__ gelu_aux, int M,
                int H, uint32_t thresh, float inv_keep, uint64_t seed, uint32_t stream) {
    extern __shared__ __align__(16) float smem[];   // [kLnWarps][3][H]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = gridDim.x * kLnWarps;
    float* acc_g = smem + (size_t)warp * 3 * H;
    float* acc_b = acc_g + H;
    float* acc_bias = acc_b + H;
    for (int i = lane; i < 3 * H; i += 32) acc_g[i] = 0.f;
    __syncwarp();
    for (int row = blockIdx.x * kLnWarps + warp; row < M; row += nw) {
        // every load of the row goes in flight before the first one is consumed (the kernel is latency-bound otherwise)
        RowRawB<NCH> y_raw, g1_raw;
        RowRawF<NCH> res_raw, g2_raw;
        row_fetch_bf16(y_raw, y + (size_t)row * H, H, lane);
        if (res != nullptr) row_fetch_f32(res_raw, res + (size_t)row * H, H, lane);
        row_fetch_bf16(g1_raw, g1 + (size_t)row * H, H, lane);
        if (g2 != nullptr) row_fetch_f32(g2_raw, g2 + (size_t)row * H, H, lane);
        const uint32_t mask = row_dropout_mask<NCH>(H, lane, seed, stream, (uint64_t)row, thresh);   // hashed under the loads
        // recompute the LayerNorm input exactly as the forward did: z = dropout(y) + res
        RowF<NCH> z;
        row_unpack_bf16(z, y_raw);
        row_apply_mask(z, mask, inv_keep);
        if (res != nullptr) row_add_f32(z, res_raw);
        RowF<NCH> g;
        row_unpack_bf16(g, g1_raw);
        if (g2 != nullptr) row_add_f32(g, g2_raw);
        const float mean = mean_in[row], rstd = rstd_in[row];
        // dbeta += dy ; dgamma += dy * xhat ; dz = (dy*gamma - mean(dy*gamma) - xhat * mean(dy*gamma*xhat)) * rstd
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int e = (c * 32 + lane) * 8;
            if (e < H) {
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + e));
                const float4 g1v = __ldg(reinterpret_cast<const float4*>(gamma + e + 4));
                const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1v.x, g1v.y, g1v.z, g1v.w};
                float4* pg = reinterpret_cast<float4*>(acc_g + e);
                float4* pb = reinterpret_cast<float4*>(acc_b + e);
                const float4 ag0 = pg[0], ag1 = pg[1], ab0 = pb[0], ab1 = pb[1];
                float ag[8] = {ag0.x, ag0.y, ag0.z, ag0.w, ag1.x, ag1.y, ag1.z, ag1.w};
                float ab[8] = {ab0.x, ab0.y, ab0.z, ab0.w, ab1.x, ab1.y, ab1.z, ab1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float xhat = (z.v[c][i] - mean) * rstd;
                    z.v[c][i] = xhat;
                    const float dy = g.v[c][i];
                    ag[i] = fmaf(dy, xhat, ag[i]);
                    ab[i] += dy;
                    const float dg = dy * gm[i];
                    g.v[c][i] = dg;
                    s1 += dg;
                    s2 = fmaf(dg, xhat, s2);
                }
                pg[0] = make_float4(ag[0], ag[1], ag[2], ag[3]);
                pg[1] = make_float4(ag[4], ag[5], ag[6], ag[7]);
                pb[0] = make_float4(ab[0], ab[1], ab[2], ab[3]);
                pb[1] = make_float4(ab[4], ab[5], ab[6], ab[7]);
            }
        }
        s1 = warp_sum(s1) / (float)H;
        s2 = warp_sum(s2) / (float)H;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int e = (c * 32 + lane) * 8;
            if (e < H) {
#pragma unroll
                for (int i = 0; i < 8; ++i) g.v[c][i] = (g.v[c][i] - s1 - z.v[c][i] * s2) * rstd;
            }
        }
        // g now holds dz: gradient of the residual branch
        if (d_res != nullptr) row_store_f32(g, d_res + (size_t)row * H, H, lane);
        // gradient of the dense output (pre-dropout): dz * mask / keep
        row_apply_mask(g, mask, inv_keep);   // same mask, same scaling
        if (gelu_aux != nullptr) {  // y = gelu(aux): chain through the activation (LM-head transform, :482-484)
            RowF<NCH> u;
            row_load_bf16(u, gelu_aux + (size_t)row * H, H, lane);
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int i = 0; i < 8; ++i) g.v[c][i] *= gelu_erf_grad(u.v[c][i]);
        }
        row_round_bf16(g);  // the bias gradient sums exactly what the wgrad GEMM will read
        row_store_bf16(g, d_y + (size_t)row * H, H, lane);
        if (dbias != nullptr) smem_accumulate(acc_bias, g, H, lane);
    }
    __syncthreads();
    for (int col = threadIdx.x; col < H; col += kLnWarps * 32) {
        float sg = 0.f, sb = 0.f, sbias = 0.f;
#pragma unroll
        for (int w = 0; w < kLnWarps; ++w) {
            const float* a = smem + (size_t)w * 3 * H;
            sg += a[col];
            sb += a[H + col];
            sbias += a[2 * H + col];
        }
        atomicAdd(dgamma + col, sg);
        atomicAdd(dbeta + col, sb);
        if (dbias != nullptr) atomicAdd(dbias + col, sbias);
    }
}

// out[n] += sum_m X[m, n]   (X bf16, row stride ld).  Each thread owns 8 adjacent columns.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ X, float* __restrict__ out, int M, int N, int64_t ld,
                   int rows_per_cta) {
    __shared__ float red[8][32 * 8 + 1];
    const int cg = threadIdx.x & 31;  // column group within the CTA's 256-column strip
    const int rr = threadIdx.x >> 5;  // row lane 0..7
    const int col0 = blockIdx.x * 256 + cg * 8;
    const int r0 = blockIdx.y * rows_per_cta;
    const int r1 = min(M, r0 + rows_per_cta);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (col0 < N) {
        int r = r0 + rr;
        for (; r + 24 < r1; r += 32) {   // four independent 16-byte loads in flight per thread
            uint4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(X + (size_t)(r + 8 * u) * ld + col0));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float2 a = unpack_bf16x2(q[u].x), b = unpack_bf16x2(q[u].y), c = unpack_bf16x2(q[u].z), d = unpack_bf16x2(q[u].w);
                acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
                acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
            }
        }
        for (; r < r1; r += 8) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(X + (size_t)r * ld + col0));
            const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
            acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
            acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[rr][cg * 8 + i] = acc[i];
    __syncthreads();
    const int col = blockIdx.x * 256 + threadIdx.x;
    if (col < N) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += red[r][threadIdx.x];
        atomicAdd(out + col, s);
    }
}

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_dropout_residual_ln_fwd(const mmb_drln_fwd_args* a, void* stream) {
    MMB_REQUIRE(a && a->y && a->gamma && a->beta, "drln_fwd: null pointer");
    MMB_REQUIRE(a->y_f32 ? (a->out_f32 != nullptr) : (a->out && a->mean && a->rstd), "drln_fwd: null output");
    MMB_REQUIRE(a->M > 0 && a->H > 0 && a->H % 8 == 0 && a->H <= 1024, "drln_fwd: bad shape M=%d H=%d", a->M, a->H);
    const uint32_t thresh = dropout_threshold(a->p_drop);
    const float inv_keep = dropout_inv_keep(a->p_drop);
    const int grid = min((a->M + kLnWarps - 1) / kLnWarps, num_sms() * 4);
    if (a->y_f32) {
        MMB_DISPATCH_NCH(a->H, (drln_fwd_kernel<NCH, true><<<grid, kLnWarps * 32, 0, (cudaStream_t)stream>>>(
                                   a->y, a->res, a->gamma, a->beta, (__nv_bfloat16*)a->out, a->out_f32, a->mean, a->rstd, a->M,
                                   a->H, a->eps, thresh, inv_keep, a->seed, a->rng_stream)));
        return check_launch("drln_fwd_kernel<f32>");
    }
    MMB_DISPATCH_NCH(a->H, (drln_fwd_kernel<NCH><<<grid, kLnWarps * 32, 0, (cudaStream_t)stream>>>(
                               a->y, a->res, a->gamma, a->beta,
                               (__nv_bfloat16*)a->out, a->out_f32, a->mean, a->rstd, a->M, a->H, a->eps, thresh, inv_keep, a->seed,
                               a->rng_stream)));
    return check_launch("drln_fwd_kernel");
}

extern "C" int mmb_dropout_residual_ln_bwd(const mmb_drln_bwd_args* a, void* stream) {
    MMB_REQUIRE(a && a->g1 && a->y && a->mean && a->rstd && a->gamma && a->d_y && a->dgamma && a->dbeta,
                "drln_bwd: null pointer");
    MMB_REQUIRE(a->M > 0 && a->H > 0 && a->H % 8 == 0 && a->H <= 1024, "drln_bwd: bad shape M=%d H=%d", a->M, a->H);
    const uint32_t thresh = dropout_threshold(a->p_drop);
    const float inv_keep = dropout_inv_keep(a->p_drop);
    const int grid = min((a->M + kLnWarps - 1) / kLnWarps, num_sms() * 2);
    const size_t smem = (size_t)kLnWarps * 3 * a->H * sizeof(float);
    MMB_DISPATCH_NCH(a->H, {
        MMB_ENSURE_SMEM(100 * 1024, drln_bwd_kernel<NCH>);
    });
    MMB_DISPATCH_NCH(a->H, (drln_bwd_kernel<NCH><<<grid, kLnWarps * 32, smem, (cudaStream_t)stream>>>(
                               (const __nv_bfloat16*)a->g1, a->g2, (const __nv_bfloat16*)a->y,
                               a->res, a->mean, a->rstd, a->gamma, (__nv_bfloat16*)a->d_y,
                               a->d_res, a->dgamma, a->dbeta, a->dbias,
                               (const __nv_bfloat16*)a->gelu_aux, a->M, a->H, thresh, inv_keep, a->seed, a->rng_stream)));
    return check_launch("drln_bwd_kernel");
}

extern "C" int mmb_colsum_bf16(const mmb_colsum_args* a, void* stream) {
    MMB_REQUIRE(a && a->X && a->out, "colsum: null pointer");
    MMB_REQUIRE(a->M > 0 && a->N > 0 && a->N % 8 == 0 && a->ld % 8 == 0, "colsum: bad shape M=%d N=%d ld=%lld", a->M,
                a->N, (long long)a->ld);
    const int strips = (a->N + 255) / 256;
    int ysplit = (num_sms() * 8 + strips - 1) / strips;
    if (ysplit > (a->M + 63) / 64) ysplit = (a->M + 63) / 64;
    if (ysplit < 1) ysplit = 1;
    const int rows_per_cta = (a->M + ysplit - 1) / ysplit;
    dim3 grid(strips, (a->M + rows_per_cta - 1) / rows_per_cta);
    colsum_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a->X, a->out, a->M, a->N, a->ld,
                                                              rows_per_cta);
    return check_launch("colsum_bf16_kernel");
}
