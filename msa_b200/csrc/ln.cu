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

// forward: the fp32 residual row keeps its two 16-byte loads — measured 110.9 us against 122.0 with one 32-byte .nc load
// per chunk (same box, 73 600 x 768); the backward's four row streams gain from the 32-byte form (183 -> 170 us)
constexpr bool kFwdLd256 = false;
constexpr int kLnWarps = 8;

template <int NCH, bool kYF32 = false>
__global__ void __launch_bounds__(kLnWarps * 32)
drln_fwd_kernel(const void* __restrict__ y_, const float* __restrict__ res,
                const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                float* __restrict__ out_f32, float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int H, float eps, uint32_t thresh,
                float inv_keep, uint64_t seed, uint32_t stream, const int* __restrict__ row_list) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = gridDim.x * kLnWarps;
    // row list (mmb_attn_schedule): only the live rows are computed; padding rows keep their previous contents
    const int nrows = row_list != nullptr ? __ldg(row_list) : M;
    int ri = blockIdx.x * kLnWarps + warp;
    int row_next = (row_list != nullptr && ri < nrows) ? __ldg(row_list + 4 + ri) : ri;      // (the index load runs one row ahead)
    for (; ri < nrows; ri += nw) {
        const int row = row_next;
        row_next = (row_list != nullptr && ri + nw < nrows) ? __ldg(row_list + 4 + ri + nw) : ri + nw;
        RowF<NCH> z;
        if (kYF32) row_load_f32<NCH, kFwdLd256>(z, reinterpret_cast<const float*>(y_) + (size_t)row * H, H, lane);
        else row_load_bf16(z, reinterpret_cast<const __nv_bfloat16*>(y_) + (size_t)row * H, H, lane);
        row_dropout(z, H, lane, seed, stream, (uint64_t)row, thresh, inv_keep);
        if (res != nullptr) {
            RowF<NCH> r;
            row_load_f32<NCH, kFwdLd256>(r, res + (size_t)row * H, H, lane);
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

// Backward.  All element-wise arithmetic runs on PAIRS of adjacent columns with the packed fp32x2 forms (FFMA2 / FADD2 /
// FMUL2, sm_100): the scalar version issued ~1 180 instructions per row and was bound by issue slots at 61-70 % of the
// copy bandwidth; the keep decisions become a float2 multiplier (keep ? 1 / (1 - p) : 0) built once per pair and used for
// both the recomputed LayerNorm input and the outgoing gradient.
// 2 CTAs x 6 warps per SM: 12 resident warps leave ~170 registers per thread — with 16 (128 registers) the 18 row vectors in
// flight plus the two 24-value arrays kept across the warp reduction spilled ~360 bytes per thread and row
constexpr int kLnBwdWarps = 6;
template <int NCH>
__global__ void __launch_bounds__(kLnBwdWarps * 32, 2)
drln_bwd_kernel(const __nv_bfloat16* __restrict__ g1, const float* __restrict__ g2,
                const __nv_bfloat16* __restrict__ y, const float* __restrict__ res,
                const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ gamma,
                __nv_bfloat16* __restrict__ d_y, float* __restrict__ d_res, float* __restrict__ dgamma,
                float* __restrict__ dbeta, float* __restrict__ dbias, const __nv_bfloat16* __restrict__ gelu_aux, int M,
                int H, uint32_t thresh, float inv_keep, uint64_t seed, uint32_t stream, const int* __restrict__ row_list,
                int dead_rows_zeroed) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    extern __shared__ __align__(16) float smem[];   // [kLnBwdWarps][3][H]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = gridDim.x * kLnBwdWarps;
    float* acc_g = smem + (size_t)warp * 3 * H;
    float* acc_b = acc_g + H;
    float* acc_bias = acc_b + H;
    for (int i = lane; i < 3 * H; i += 32) acc_g[i] = 0.f;
    __syncwarp();
    const bool drop = thresh != 0u;
    // row list (mmb_attn_schedule): the live rows are computed; the gradient of every other row is exactly zero and is
    // written as such without reading anything (the wgrad / dgrad GEMMs read every row of d_y, the next LayerNorm
    // backward every live row of d_res)
    const int nrows = row_list != nullptr ? __ldg(row_list) : M;
    if (row_list != nullptr && !dead_rows_zeroed) {
        const int ndead = __ldg(row_list + 2) - nrows;
        for (int ri = blockIdx.x * kLnBwdWarps + warp; ri < ndead; ri += nw) {
            const int row = __ldg(row_list + 4 + nrows + ri);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int e = (c * 32 + lane) * 8;
                if (e < H) {
                    *reinterpret_cast<uint4*>(d_y + (size_t)row * H + e) = make_uint4(0u, 0u, 0u, 0u);
                    if (d_res != nullptr)
                        stg256_f32(d_res + (size_t)row * H + e, make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f));
                }
            }
        }
    }
    int ri = blockIdx.x * kLnBwdWarps + warp;
    int row_next = (row_list != nullptr && ri < nrows) ? __ldg(row_list + 4 + ri) : ri;      // (the index load runs one row ahead)
    for (; ri < nrows; ri += nw) {
        const int row = row_next;
        row_next = (row_list != nullptr && ri + nw < nrows) ? __ldg(row_list + 4 + ri + nw) : ri + nw;
        // every load of the row goes in flight before the first one is consumed (the kernel is latency-bound otherwise)
        RowRawB<NCH> y_raw, g1_raw;
        RowRawF<NCH> res_raw, g2_raw;
        row_fetch_bf16(y_raw, y + (size_t)row * H, H, lane);
        if (res != nullptr) row_fetch_f32(res_raw, res + (size_t)row * H, H, lane);
        row_fetch_bf16(g1_raw, g1 + (size_t)row * H, H, lane);
        if (g2 != nullptr) row_fetch_f32(g2_raw, g2 + (size_t)row * H, H, lane);
        const float mean = mean_in[row], rstd = rstd_in[row];
        const uint32_t key = drop ? rng_row_key(seed, stream, (uint32_t)row) : 0u;     // hashed under the loads
        const float2 rstd2 = make_float2(rstd, rstd), nmr2 = make_float2(-mean * rstd, -mean * rstd);
        float2 xh[NCH][4], dg[NCH][4];
        uint32_t keep = 0u;          // one bit per element of this lane (bit c * 8 + 2 i + {0, 1})
        float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int e = (c * 32 + lane) * 8;
            if (e < H) {
                const uint32_t yw[4] = {y_raw.q[c].x, y_raw.q[c].y, y_raw.q[c].z, y_raw.q[c].w};
                const uint32_t gw[4] = {g1_raw.q[c].x, g1_raw.q[c].y, g1_raw.q[c].z, g1_raw.q[c].w};
                const float2 rs[4] = {make_float2(res_raw.a[c].x, res_raw.a[c].y), make_float2(res_raw.a[c].z, res_raw.a[c].w),
                                      make_float2(res_raw.b[c].x, res_raw.b[c].y), make_float2(res_raw.b[c].z, res_raw.b[c].w)};
                const float2 gs[4] = {make_float2(g2_raw.a[c].x, g2_raw.a[c].y), make_float2(g2_raw.a[c].z, g2_raw.a[c].w),
                                      make_float2(g2_raw.b[c].x, g2_raw.b[c].y), make_float2(g2_raw.b[c].z, g2_raw.b[c].w)};
                const float4 gm0 = __ldg(reinterpret_cast<const float4*>(gamma + e));
                const float4 gm1 = __ldg(reinterpret_cast<const float4*>(gamma + e + 4));
                const float2 gm[4] = {make_float2(gm0.x, gm0.y), make_float2(gm0.z, gm0.w), make_float2(gm1.x, gm1.y),
                                      make_float2(gm1.z, gm1.w)};
                float4* pg = reinterpret_cast<float4*>(acc_g + e);
                float4* pb = reinterpret_cast<float4*>(acc_b + e);
                const float4 ag0 = pg[0], ag1 = pg[1], ab0 = pb[0], ab1 = pb[1];
                float2 ag[4] = {make_float2(ag0.x, ag0.y), make_float2(ag0.z, ag0.w), make_float2(ag1.x, ag1.y), make_float2(ag1.z, ag1.w)};
                float2 ab[4] = {make_float2(ab0.x, ab0.y), make_float2(ab0.z, ab0.w), make_float2(ab1.x, ab1.y), make_float2(ab1.z, ab1.w)};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    // recompute the LayerNorm input exactly as the forward did: z = dropout(y) + res
                    float2 z = unpack_bf16x2(yw[i]);
                    if (drop) {
                        const uint32_t bits = rng_pair(key, (uint32_t)(e >> 1) + i);
                        const bool k0 = rng_keep_lo(bits, thresh), k1 = rng_keep_hi(bits, thresh);
                        keep |= ((k0 ? 1u : 0u) | (k1 ? 2u : 0u)) << (c * 8 + 2 * i);
                        z = mul2(z, make_float2(k0 ? inv_keep : 0.f, k1 ? inv_keep : 0.f));
                    }
                    if (res != nullptr) z = add2(z, rs[i]);
                    float2 dy = unpack_bf16x2(gw[i]);
                    if (g2 != nullptr) dy = add2(dy, gs[i]);
                    // dbeta += dy ; dgamma += dy * xhat ; dz = (dy*gamma - mean(dy*gamma) - xhat * mean(dy*gamma*xhat)) * rstd
                    const float2 x = fma2(z, rstd2, nmr2);
                    xh[c][i] = x;
                    ag[i] = fma2(dy, x, ag[i]);
                    ab[i] = add2(ab[i], dy);
                    const float2 d = mul2(dy, gm[i]);
                    dg[c][i] = d;
                    s1 = add2(s1, d);
                    s2 = fma2(d, x, s2);
                }
                pg[0] = make_float4(ag[0].x, ag[0].y, ag[1].x, ag[1].y);
                pg[1] = make_float4(ag[2].x, ag[2].y, ag[3].x, ag[3].y);
                pb[0] = make_float4(ab[0].x, ab[0].y, ab[1].x, ab[1].y);
                pb[1] = make_float4(ab[2].x, ab[2].y, ab[3].x, ab[3].y);
            }
        }
        const float m1 = warp_sum(s1.x + s1.y) / (float)H;
        const float m2 = warp_sum(s2.x + s2.y) / (float)H;
        const float2 c1 = make_float2(-m1 * rstd, -m1 * rstd), c2 = make_float2(-m2 * rstd, -m2 * rstd);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int e = (c * 32 + lane) * 8;
            if (e < H) {
                float2 dz[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) dz[i] = fma2(xh[c][i], c2, fma2(dg[c][i], rstd2, c1));    // (dg - m1 - xhat m2) rstd
                // dz: gradient of the residual branch
                if (d_res != nullptr)
                    stg256_f32(d_res + (size_t)row * H + e, make_float4(dz[0].x, dz[0].y, dz[1].x, dz[1].y),
                               make_float4(dz[2].x, dz[2].y, dz[3].x, dz[3].y));
                // gradient of the dense output (pre-dropout): dz * mask / keep — same mask, same scaling
                if (drop) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        dz[i] = mul2(dz[i], make_float2(((keep >> (c * 8 + 2 * i)) & 1u) ? inv_keep : 0.f,
                                                        ((keep >> (c * 8 + 2 * i + 1)) & 1u) ? inv_keep : 0.f));
                }
                if (gelu_aux != nullptr) {  // y = gelu(aux): chain through the activation (LM-head transform, :482-484)
                    const uint4 u = __ldg(reinterpret_cast<const uint4*>(gelu_aux + (size_t)row * H + e));
                    const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 uv = unpack_bf16x2(uw[i]);
                        dz[i].x *= gelu_erf_grad(uv.x);
                        dz[i].y *= gelu_erf_grad(uv.y);
                    }
                }
                uint4 q;
                q.x = pack_bf16x2(dz[0].x, dz[0].y);
                q.y = pack_bf16x2(dz[1].x, dz[1].y);
                q.z = pack_bf16x2(dz[2].x, dz[2].y);
                q.w = pack_bf16x2(dz[3].x, dz[3].y);
                *reinterpret_cast<uint4*>(d_y + (size_t)row * H + e) = q;
                if (dbias != nullptr) {     // the bias gradient sums exactly what the wgrad GEMM will read (bf16-rounded)
                    const float2 r0 = unpack_bf16x2(q.x), r1 = unpack_bf16x2(q.y), r2 = unpack_bf16x2(q.z), r3 = unpack_bf16x2(q.w);
                    float4* pz = reinterpret_cast<float4*>(acc_bias + e);
                    float4 a = pz[0], b = pz[1];
                    const float2 a0 = add2(make_float2(a.x, a.y), r0), a1 = add2(make_float2(a.z, a.w), r1);
                    const float2 b0 = add2(make_float2(b.x, b.y), r2), b1 = add2(make_float2(b.z, b.w), r3);
                    pz[0] = make_float4(a0.x, a0.y, a1.x, a1.y);
                    pz[1] = make_float4(b0.x, b0.y, b1.x, b1.y);
                }
            }
        }
    }
    __syncthreads();
    for (int col = threadIdx.x; col < H; col += kLnBwdWarps * 32) {
        float sg = 0.f, sb = 0.f, sbias = 0.f;
#pragma unroll
        for (int w = 0; w < kLnBwdWarps; ++w) {
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
                   int rows_per_cta, const int* __restrict__ row_list) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    __shared__ float red[8][32 * 8 + 1];
    const int cg = threadIdx.x & 31;  // column group within the CTA's 256-column strip
    const int rr = threadIdx.x >> 5;  // row lane 0..7
    const int col0 = blockIdx.x * 256 + cg * 8;
    // row list (mmb_attn_schedule): only the live rows are read (the others hold zeros); the grid was sized for M rows
    const int* rl = row_list != nullptr ? row_list + 4 : nullptr;
    if (rl != nullptr) {
        M = __ldg(row_list);
        rows_per_cta = ((M + (int)gridDim.y - 1) / (int)gridDim.y + 7) & ~7;
    }
    auto row_of = [&](int i) { return rl != nullptr ? __ldg(rl + i) : i; };
    const int r0 = blockIdx.y * rows_per_cta;
    const int r1 = min(M, r0 + rows_per_cta);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (col0 < N) {
        int r = r0 + rr;
        for (; r + 24 < r1; r += 32) {   // four independent 16-byte loads in flight per thread
            uint4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(X + (size_t)row_of(r + 8 * u) * ld + col0));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float2 a = unpack_bf16x2(q[u].x), b = unpack_bf16x2(q[u].y), c = unpack_bf16x2(q[u].z), d = unpack_bf16x2(q[u].w);
                acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
                acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
            }
        }
        for (; r < r1; r += 8) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(X + (size_t)row_of(r) * ld + col0));
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
        MMB_DISPATCH_NCH(a->H, (launch_pdl(drln_fwd_kernel<NCH, true>, dim3(grid), dim3(kLnWarps * 32), (size_t)(0), (cudaStream_t)stream, 
                                   a->y, a->res, a->gamma, a->beta, (__nv_bfloat16*)a->out, a->out_f32, a->mean, a->rstd, a->M,
                                   a->H, a->eps, thresh, inv_keep, a->seed, a->rng_stream, a->row_list)));
        return check_launch("drln_fwd_kernel<f32>");
    }
    MMB_DISPATCH_NCH(a->H, (launch_pdl(drln_fwd_kernel<NCH>, dim3(grid), dim3(kLnWarps * 32), (size_t)(0), (cudaStream_t)stream, 
                               a->y, a->res, a->gamma, a->beta,
                               (__nv_bfloat16*)a->out, a->out_f32, a->mean, a->rstd, a->M, a->H, a->eps, thresh, inv_keep, a->seed,
                               a->rng_stream, a->row_list)));
    return check_launch("drln_fwd_kernel");
}

extern "C" int mmb_dropout_residual_ln_bwd(const mmb_drln_bwd_args* a, void* stream) {
    MMB_REQUIRE(a && a->g1 && a->y && a->mean && a->rstd && a->gamma && a->d_y && a->dgamma && a->dbeta,
                "drln_bwd: null pointer");
    MMB_REQUIRE(a->M > 0 && a->H > 0 && a->H % 8 == 0 && a->H <= 1024, "drln_bwd: bad shape M=%d H=%d", a->M, a->H);
    const uint32_t thresh = dropout_threshold(a->p_drop);
    const float inv_keep = dropout_inv_keep(a->p_drop);
    const int grid = min((a->M + kLnBwdWarps - 1) / kLnBwdWarps, num_sms() * 2);
    const size_t smem = (size_t)kLnBwdWarps * 3 * a->H * sizeof(float);
    MMB_DISPATCH_NCH(a->H, {
        MMB_ENSURE_SMEM(100 * 1024, drln_bwd_kernel<NCH>);
    });
    MMB_DISPATCH_NCH(a->H, (launch_pdl(drln_bwd_kernel<NCH>, dim3(grid), dim3(kLnBwdWarps * 32), (size_t)(smem), (cudaStream_t)stream, 
                               (const __nv_bfloat16*)a->g1, a->g2, (const __nv_bfloat16*)a->y,
                               a->res, a->mean, a->rstd, a->gamma, (__nv_bfloat16*)a->d_y,
                               a->d_res, a->dgamma, a->dbeta, a->dbias,
                               (const __nv_bfloat16*)a->gelu_aux, a->M, a->H, thresh, inv_keep, a->seed, a->rng_stream, a->row_list,
                               a->dead_rows_zeroed)));
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
    launch_pdl(colsum_bf16_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, (const __nv_bfloat16*)a->X, a->out, a->M, a->N, a->ld,
                                                              rows_per_cta, a->row_list);
    return check_launch("colsum_bf16_kernel");
}
