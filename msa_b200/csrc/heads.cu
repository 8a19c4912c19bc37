// Pooler, alignment / NSP heads, score-attention fusion, sentiment head, CPC and the loss combination,
// forward and backward, on the 3B [CLS] rows of the packed batch.  Everything here is O(B * H^2) with B ~ 64
// (< 0.1 % of the path's FLOPs), so it runs in fp32 on CUDA cores as a short fixed sequence of small kernels
// launched from one C entry point (no host synchronisation, scratch in a caller-provided workspace).
//
// Replaces
//   BertPooler.forward                modeling_bert.py:462-468
//   MMBertPreTrainingHeads.forward    MMBertForPretraining.py:295-302 (align on seq[:,0]; seq_relationship on pooled)
//   fusion / classifier               MMBertForPretraining.py:406-415
//   CPC.forward x3                    MMBertEmbedding.py:21-32 (in-batch negatives)
//   losses                            MMBertForPretraining.py:386-388 (AP CE), :427-443
// and their autograd backward.
#include "common.cuh"

namespace mmb {

enum { ACT_NONE = MMB_ACT_NONE, ACT_TANH = MMB_ACT_TANH, ACT_RELU = MMB_ACT_RELU, ACT_GELU = MMB_ACT_GELU };

// Generic fp32 tiled GEMM for the head linears and their backward (64 x 64 tiles, 16-deep k-steps, 4 x 4 outputs
// per thread).  C[m][n] (+)= act(sum_k A(m,k) B(k,n) + bias[n]) with A(m,k) = A[m*sam + k*sak] and
// B(k,n) = B[k*sbk + n*sbn], so that Y = X W^T, dX = dY W and dW = dY^T X are the same kernel.
constexpr int kSgBK = 64;   // deep k-steps: fewer load -> sync -> compute rounds for these latency-bound GEMMs
constexpr int kSgT = 32;    // 32 x 32 output tiles: the largest head GEMM (192 x 768) still yields 144 CTAs
__global__ void __launch_bounds__(256)
sgemm64_kernel(const float* __restrict__ A, int sam, int sak, const float* __restrict__ B, int sbk, int sbn,
               float* __restrict__ C, int ldc, const float* __restrict__ bias, int act, int M, int N, int K, int accumulate) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    __shared__ __align__(16) float As[kSgBK][kSgT + 2];
    __shared__ __align__(16) float Bs[kSgBK][kSgT + 2];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * kSgT, n0 = blockIdx.x * kSgT;
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    float ra[8], rb[8];
    // all 16 loads of a k-tile are issued together, and the NEXT tile's loads fly under the current tile's math
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + i * 256;  // 2048 = 32 x 64 elements per operand tile
            int m, k;
            if (sak == 1) { m = idx >> 6; k = idx & 63; } else { m = idx & 31; k = idx >> 5; }
            ra[i] = (m0 + m < M && k0 + k < K) ? __ldg(A + (size_t)(m0 + m) * sam + (size_t)(k0 + k) * sak) : 0.f;
            int kb, n;
            if (sbn == 1) { kb = idx >> 5; n = idx & 31; } else { kb = idx & 63; n = idx >> 6; }
            rb[i] = (n0 + n < N && k0 + kb < K) ? __ldg(B + (size_t)(k0 + kb) * sbk + (size_t)(n0 + n) * sbn) : 0.f;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += kSgBK) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + i * 256;
            int m, k;
            if (sak == 1) { m = idx >> 6; k = idx & 63; } else { m = idx & 31; k = idx >> 5; }
            As[k][m] = ra[i];
            int kb, n;
            if (sbn == 1) { kb = idx >> 5; n = idx & 31; } else { kb = idx & 63; n = idx >> 6; }
            Bs[kb][n] = rb[i];
        }
        __syncthreads();
        if (k0 + kSgBK < K) fetch(k0 + kSgBK);
#pragma unroll 16
        for (int k = 0; k < kSgBK; ++k) {
            const float2 a = *reinterpret_cast<const float2*>(&As[k][ty * 2]);
            const float2 bv = *reinterpret_cast<const float2*>(&Bs[k][tx * 2]);
            acc[0][0] = fmaf(a.x, bv.x, acc[0][0]);
            acc[0][1] = fmaf(a.x, bv.y, acc[0][1]);
            acc[1][0] = fmaf(a.y, bv.x, acc[1][0]);
            acc[1][1] = fmaf(a.y, bv.y, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + ty * 2 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int n = n0 + tx * 2 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            if (act == ACT_TANH) v = tanhf(v);
            else if (act == ACT_RELU) v = fmaxf(v, 0.f);
            else if (act == ACT_GELU) v = gelu_erf(v);
            float* c = C + (size_t)m * ldc + n;
            *c = accumulate ? *c + v : v;
        }
    }
}
// db[n] += sum_r dY[r][n]
__global__ void bias_grad_kernel(const float* __restrict__ dY, int ldy, float* __restrict__ db, int R, int N) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += dY[(size_t)r * ldy + n];
    db[n] += s;
}

// ---------------------------------------------------------------- small fused elementwise kernels
__global__ void gather_cls_kernel(const __nv_bfloat16* __restrict__ seq, const int* __restrict__ cu, float* __restrict__ X0,
                                  int R, int H) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int r = blockIdx.x;
    const __nv_bfloat16* src = seq + (size_t)cu[r] * H;
    for (int j = threadIdx.x; j < H; j += blockDim.x) X0[(size_t)r * H + j] = __bfloat162float(src[j]);
}
__global__ void gather_cls_f32_kernel(const float* __restrict__ seq, const int* __restrict__ cu, float* __restrict__ X0, int R,
                                      int H) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int r = blockIdx.x;
    const float* src = seq + (size_t)cu[r] * H;
    for (int j = threadIdx.x; j < H; j += blockDim.x) X0[(size_t)r * H + j] = src[j];
}
__global__ void scatter_cls_grad_kernel(const float* __restrict__ dX0, const int* __restrict__ cu, __nv_bfloat16* __restrict__ g,
                                        int R, int H) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int r = blockIdx.x;
    __nv_bfloat16* dst = g + (size_t)cu[r] * H;
    for (int j = threadIdx.x; j < H; j += blockDim.x)
        dst[j] = __float2bfloat16_rn(__bfloat162float(dst[j]) + dX0[(size_t)r * H + j]);
}
__global__ void dup_cols_kernel(const float* __restrict__ P, float* __restrict__ PP, int R, int H) {  // PP = [P, P]
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int r = blockIdx.x;
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        const float v = P[(size_t)r * H + j];
        PP[(size_t)r * 2 * H + j] = v;
        PP[(size_t)r * 2 * H + H + j] = v;
    }
}
struct VPtrs {  // per-modality score vectors (0 = vt, 1 = vv, 2 = vs) and their gradients
    const float* w[3];
    const float* b[3];
    float* gw[3];
    float* gb[3];
};
// s[r] = U[r] . v_m + b_m ; PC[b][m*H + j] = P[r][j] * s[r]      (r = m*B + b; m: 0 = vt, 1 = vv, 2 = vs)
__global__ void __launch_bounds__(256)
score_scale_kernel(const float* __restrict__ U, const float* __restrict__ P, const VPtrs vp, float* __restrict__ s,
                   float* __restrict__ PC, int B, int H) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    __shared__ float red[8];
    const int r = blockIdx.x, m = r / B, b = r - m * B;
    const float* v = vp.w[m];
    float acc = 0.f;
    for (int j = threadIdx.x; j < H; j += 256) acc = fmaf(U[(size_t)r * H + j], v[j], acc);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w];
    tot += vp.b[m][0];
    if (threadIdx.x == 0) s[r] = tot;
    for (int j = threadIdx.x; j < H; j += 256) PC[(size_t)b * 3 * H + m * H + j] = P[(size_t)r * H + j] * tot;
}
// backward of the above: ds[r] = sum_j dPC[b][mH+j] P[r][j]; dP[r][j] += dPC * s[r]; dU[r][j] = ds v_m[j] (masked by U>0);
// dv_m[j] += ds U[r][j]; dbv_m += ds
__global__ void __launch_bounds__(256)
score_scale_bwd_kernel(const float* __restrict__ dPC, const float* __restrict__ P, const float* __restrict__ U,
                       const float* __restrict__ s, const VPtrs vp, float* __restrict__ dP, float* __restrict__ dA, int B,
                       int H) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    __shared__ float red[8];
    const int r = blockIdx.x, m = r / B, b = r - m * B;
    float acc = 0.f;
    for (int j = threadIdx.x; j < H; j += 256) acc = fmaf(dPC[(size_t)b * 3 * H + m * H + j], P[(size_t)r * H + j], acc);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    float ds = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) ds += red[w];
    const float sr = s[r];
    const float* v = vp.w[m];
    for (int j = threadIdx.x; j < H; j += 256) {
        const size_t idx = (size_t)r * H + j;
        dP[idx] += dPC[(size_t)b * 3 * H + m * H + j] * sr;
        const float u = U[idx];
        dA[idx] = u > 0.f ? ds * v[j] : 0.f;
        atomicAdd(vp.gw[m] + j, ds * u);
    }
    if (threadIdx.x == 0) atomicAdd(vp.gb[m], ds);
}

// CPC forward for one modality: row i -> xn_i, an_i (unit vectors), softmax row Sm_i over G_ij = xn_i . an_j,
// accumulates nce += -(pos_i - neg_i) / B.   Two kernels: normalise, then rows.
__global__ void __launch_bounds__(256)
normalize_rows_kernel(const float* __restrict__ X, float* __restrict__ Y, float* __restrict__ norms, int H) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    __shared__ float red[8];
    const int r = blockIdx.x;
    float acc = 0.f;
    for (int j = threadIdx.x; j < H; j += 256) {
        const float v = X[(size_t)r * H + j];
        acc = fmaf(v, v, acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w];
    const float nrm = sqrtf(tot);
    if (threadIdx.x == 0) norms[r] = nrm;
    for (int j = threadIdx.x; j < H; j += 256) Y[(size_t)r * H + j] = X[(size_t)r * H + j] / nrm;
}
// one CTA per row i; dynamic smem: B floats
__global__ void __launch_bounds__(256)
cpc_rows_kernel(const float* __restrict__ xn, const float* __restrict__ an, float* __restrict__ Sm, float* __restrict__ nce,
                int B, int H) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    extern __shared__ float g[];
    __shared__ float red[8];
    const int i = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // g[j] = <xn_i, an_j>: row i is read once into registers (H <= 1024: 8 float4 per lane), two rows j per pass with
    // 16-byte loads so that many loads are in flight (the scalar one-row-at-a-time loop was latency-bound: 58 us)
    const float4* xi = reinterpret_cast<const float4*>(xn + (size_t)i * H);
    const int nv = H >> 2;                       // H % 4 == 0 (H = heads * 64)
    float4 xr[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) xr[t] = (lane + 32 * t) < nv ? __ldg(xi + lane + 32 * t) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 2 * warp; j < B; j += 16) {
        const float4* a0 = reinterpret_cast<const float4*>(an + (size_t)j * H);
        const float4* a1 = reinterpret_cast<const float4*>(an + (size_t)min(j + 1, B - 1) * H);
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if ((lane + 32 * t) < nv) {
                const float4 u = __ldg(a0 + lane + 32 * t), v = __ldg(a1 + lane + 32 * t);
                acc0 += xr[t].x * u.x + xr[t].y * u.y + xr[t].z * u.z + xr[t].w * u.w;
                acc1 += xr[t].x * v.x + xr[t].y * v.y + xr[t].z * v.z + xr[t].w * v.w;
            }
        }
        acc0 = warp_sum(acc0);
        acc1 = warp_sum(acc1);
        if (lane == 0) {
            g[j] = acc0;
            if (j + 1 < B) g[j + 1] = acc1;
        }
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < B; j += 256) mx = fmaxf(mx, g[j]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int j = threadIdx.x; j < B; j += 256) sum += expf(g[j] - mx);
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w];
    const float neg = mx + logf(sum);
    for (int j = threadIdx.x; j < B; j += 256) Sm[(size_t)i * B + j] = expf(g[j] - neg);
    if (threadIdx.x == 0) atomicAdd(nce, -(g[i] - neg) / (float)B);
}
// dG_ij = c * (delta_ij - Sm_ij)   (c = beta * gs / B);  in place over Sm
__global__ void cpc_dg_kernel(float* __restrict__ Sm, const float* __restrict__ gscale, float beta, int B) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * B) return;
    const float c = beta * (gscale ? *gscale : 1.f) / (float)B;
    const int i = idx / B, j = idx - i * B;
    Sm[idx] = c * ((i == j ? 1.f : 0.f) - Sm[idx]);
}
// dv = (dy - y (y . dy)) / ||v||   per row;  out (+)= dv
__global__ void __launch_bounds__(256)
normalize_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, const float* __restrict__ norms,
                     float* __restrict__ out, int H, int accumulate) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    __shared__ float red[8];
    const int r = blockIdx.x;
    float acc = 0.f;
    for (int j = threadIdx.x; j < H; j += 256) acc = fmaf(y[(size_t)r * H + j], dy[(size_t)r * H + j], acc);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    float dot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) dot += red[w];
    const float inv = 1.f / norms[r];
    for (int j = threadIdx.x; j < H; j += 256) {
        const size_t idx = (size_t)r * H + j;
        const float v = (dy[idx] - y[idx] * dot) * inv;
        out[idx] = accumulate ? out[idx] + v : v;
    }
}
// G-shaped products for CPC backward: out[i][k] = sum_j M[i][j] Y[j][k]  (transpose = 0) or sum_j M[j][i] Y[j][k] (1)
__global__ void __launch_bounds__(256)
bb_matmul_kernel(const float* __restrict__ Mx, const float* __restrict__ Y, float* __restrict__ out, int B, int H, int transpose) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int i = blockIdx.x;
    extern __shared__ float mrow[];              // row (or column) i of Mx: B floats
    for (int j = threadIdx.x; j < B; j += 256) mrow[j] = transpose ? Mx[(size_t)j * B + i] : Mx[(size_t)i * B + j];
    __syncthreads();
    // up to 4 output columns per thread, 4 rows of Y per pass: 16 independent loads in flight
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int j = 0;
    for (; j + 4 <= B; j += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float m = mrow[j + u];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int k = threadIdx.x + 256 * c;
                if (k < H) acc[c] = fmaf(m, __ldg(Y + (size_t)(j + u) * H + k), acc[c]);
            }
        }
    }
    for (; j < B; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int k = threadIdx.x + 256 * c;
            if (k < H) acc[c] = fmaf(mrow[j], __ldg(Y + (size_t)j * H + k), acc[c]);
        }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int k = threadIdx.x + 256 * c;
        if (k < H) out[(size_t)i * H + k] = acc[c];
    }
}

struct LossParams {
    const float* al;          // [2B,2] alignment scores (visual rows then speech rows)
    const long long* ap[2];   // [B] labels
    const float* logit;       // [B] classifier output
    const float* sentiment;   // [B]
    const float* ce_loss_sum; // [3]
    const int* label_count;   // [4]: labelled rows per pass, [3] = out-of-range label / id count
    const float* nce;         // [1]
    float* losses;            // [8]: joint, mlm, ap, label, nce, mlm_t, mlm_v, mlm_s
    float* logits_out;        // [B] (tanh applied iff num_labels == 1; zeros in the classification branch)
    float* dal;               // [2B,2]  backward
    float* dlogit;            // [B]     backward
    const float* gscale;
    float alpha, beta;
    int B, num_labels;
};
// single CTA: AP cross entropies, MSE, loss combination (MMBertForPretraining.py:386-388, 427-443)
__global__ void __launch_bounds__(256)
final_losses_kernel(const LossParams p) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    __shared__ float red[8];
    float ap_acc = 0.f, mse_acc = 0.f;
    for (int i = threadIdx.x; i < 2 * p.B; i += 256) {
        const int mod = i / p.B, b = i - mod * p.B;
        const float a0 = p.al[2 * i], a1 = p.al[2 * i + 1];
        const float mx = fmaxf(a0, a1);
        const float lse = mx + logf(expf(a0 - mx) + expf(a1 - mx));
        const long long lab = p.ap[mod][b];
        ap_acc += (lse - (lab == 0 ? a0 : a1)) / (float)p.B;
    }
    const bool classify = p.num_labels != 1 && p.num_labels != 7;
    for (int i = threadIdx.x; i < p.B; i += 256) {
        float o = p.logit[i];
        if (classify) {
            // MMBertForPretraining.py:437-442 with classifier1_2 = Linear(H, 1) (:311-314): cross entropy over ONE class is
            // lse - logit[target] = 0 for target 0; any other target is out of bounds (torch raises) -> NaN here.
            // Returned "logits" = argmax(sigmoid(logits), dim=1) = 0.
            p.logits_out[i] = 0.f;
            if (p.sentiment[i] != 0.f) mse_acc += __int_as_float(0x7fc00000);
            continue;
        }
        if (p.num_labels == 1) o = tanhf(o);
        p.logits_out[i] = o;
        const float d = o - p.sentiment[i];
        mse_acc += d * d / (float)p.B;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ap_acc = warp_sum(ap_acc);
    if (lane == 0) red[warp] = ap_acc;
    __syncthreads();
    float ap_tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) ap_tot += red[w];
    __syncthreads();
    mse_acc = warp_sum(mse_acc);
    if (lane == 0) red[warp] = mse_acc;
    __syncthreads();
    float mse = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) mse += red[w];
    if (threadIdx.x == 0) {
        float mlm = 0.f;
        for (int i = 0; i < 3; ++i) {
            const float li = p.ce_loss_sum[i] / (float)p.label_count[i];  // NaN when a pass has no label, as the reference
            p.losses[5 + i] = li;
            mlm += li;
        }
        mlm /= 3.f;
        // label_count[3]: labels / token ids outside the vocabulary seen by mmb_pack_prepare / mmb_embed_fwd (torch's
        // CrossEntropyLoss / nn.Embedding device-assert there): poison the loss instead of training on garbage
        if (p.label_count[3] != 0) mlm = __int_as_float(0x7fc00000);
        const float ap = ap_tot * 0.5f, nce = p.nce[0];
        p.losses[0] = p.alpha * mlm + ap + mse - p.beta * nce;
        p.losses[1] = mlm;
        p.losses[2] = ap;
        p.losses[3] = mse;
        p.losses[4] = nce;
    }
}
__global__ void __launch_bounds__(256)
final_losses_bwd_kernel(const LossParams p) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const float gs = p.gscale ? *p.gscale : 1.f;
    for (int i = threadIdx.x; i < 2 * p.B; i += 256) {
        const int mod = i / p.B, b = i - mod * p.B;
        const float a0 = p.al[2 * i], a1 = p.al[2 * i + 1];
        const float mx = fmaxf(a0, a1);
        const float e0 = expf(a0 - mx), e1 = expf(a1 - mx);
        const float inv = 1.f / (e0 + e1);
        const long long lab = p.ap[mod][b];
        const float c = gs * 0.5f / (float)p.B;
        p.dal[2 * i] = c * (e0 * inv - (lab == 0 ? 1.f : 0.f));
        p.dal[2 * i + 1] = c * (e1 * inv - (lab == 1 ? 1.f : 0.f));
    }
    const bool classify = p.num_labels != 1 && p.num_labels != 7;
    for (int i = threadIdx.x; i < p.B; i += 256) {
        const float o = p.logits_out[i];
        float d = gs * 2.f * (o - p.sentiment[i]) / (float)p.B;
        if (p.num_labels == 1) d *= (1.f - o * o);
        p.dlogit[i] = classify ? 0.f : d;     // one-class cross entropy: softmax - onehot == 0
    }
}
__global__ void tanh_bwd_kernel(float* __restrict__ dP, const float* __restrict__ P, int n) {  // dZ = dP (1 - P^2), in place
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dP[i] *= (1.f - P[i] * P[i]);
}
__global__ void fold_dup_grad_kernel(const float* __restrict__ dPP, float* __restrict__ dP, int R, int H) {  // dP += dPP[:, :H] + dPP[:, H:]
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int r = blockIdx.x;
    for (int j = threadIdx.x; j < H; j += blockDim.x)
        dP[(size_t)r * H + j] += dPP[(size_t)r * 2 * H + j] + dPP[(size_t)r * 2 * H + H + j];
}

// ---------------------------------------------------------------- workspace layout (floats)
struct HeadsWs {
    size_t X0, P, PP, U, s, PC, temp, logit, XH, xn, an, nx, na, Sm, al, rel, nce, ptrs;
    size_t dX0, dP, dPP, dA, dPC, dtemp, dlogit, dXH, dxn, dan, dal, total;
};
static HeadsWs heads_ws(int B, int H) {
    HeadsWs w;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) / 4 * 4; return r; };
    const size_t R = 3 * (size_t)B;
    w.X0 = take(R * H); w.P = take(R * H); w.PP = take(R * 2 * H); w.U = take(R * H); w.s = take(R);
    w.PC = take((size_t)B * 3 * H); w.temp = take((size_t)B * H); w.logit = take(B);
    w.XH = take(3 * (size_t)B * H); w.xn = take(3 * (size_t)B * H); w.an = take(3 * (size_t)B * H);
    w.nx = take(3 * (size_t)B); w.na = take(3 * (size_t)B); w.Sm = take(3 * (size_t)B * B);
    w.al = take(4 * (size_t)B); w.rel = take(2 * (size_t)B); w.nce = take(4); w.ptrs = take(32);
    w.dX0 = take(R * H); w.dP = take(R * H); w.dPP = take(R * 2 * H); w.dA = take(R * H);
    w.dPC = take((size_t)B * 3 * H); w.dtemp = take((size_t)B * H); w.dlogit = take(B);
    w.dXH = take((size_t)B * H); w.dxn = take((size_t)B * H); w.dan = take((size_t)B * H); w.dal = take(4 * (size_t)B);
    w.total = o;
    return w;
}

// Y[R,N] = act(X[R,K] W[N,K]^T + b)
static inline void linear(cudaStream_t st, const float* X, int ldx, const float* W, int ldw, const float* b, float* Y, int ldy,
                          int R, int N, int K, int act) {
    dim3 g((N + kSgT - 1) / kSgT, (R + kSgT - 1) / kSgT);
    launch_pdl(sgemm64_kernel, dim3(g), dim3(256), (size_t)(0), st, X, ldx, 1, W, 1, ldw, Y, ldy, b, act, R, N, K, 0);
}
// dX[R,K] (+)= dY[R,N] W[N,K]
static inline void dx(cudaStream_t st, const float* dY, int ldy, const float* W, int ldw, float* dX, int ldx, int R, int N,
                      int K, int accumulate) {
    dim3 g((K + kSgT - 1) / kSgT, (R + kSgT - 1) / kSgT);
    launch_pdl(sgemm64_kernel, dim3(g), dim3(256), (size_t)(0), st, dY, ldy, 1, W, ldw, 1, dX, ldx, nullptr, ACT_NONE, R, K, N, accumulate);
}
// dW[N,K] += dY[R,N]^T X[R,K] ; db[N] += colsum(dY)
static inline void dw(cudaStream_t st, const float* dY, int ldy, const float* X, int ldx, float* dW, int ldw, float* db, int R,
                      int N, int K) {
    dim3 g((K + kSgT - 1) / kSgT, (N + kSgT - 1) / kSgT);
    launch_pdl(sgemm64_kernel, dim3(g), dim3(256), (size_t)(0), st, dY, 1, ldy, X, ldx, 1, dW, ldw, nullptr, ACT_NONE, N, K, R, 1);
    if (db) launch_pdl(bias_grad_kernel, dim3((N + 127) / 128), dim3(128), (size_t)(0), st, dY, ldy, db, R, N);
}

}  // namespace mmb

using namespace mmb;

extern "C" size_t mmb_heads_workspace_bytes(int B, int H) { return heads_ws(B, H).total * sizeof(float); }

static int heads_check(const mmb_heads_args* a) {
    MMB_REQUIRE(a && a->seq_out && a->cu_seqlens && a->workspace && a->losses && a->logits_out, "heads: null pointer");
    MMB_REQUIRE(a->B > 0 && a->H > 0 && a->H % 4 == 0 && a->H <= 1024, "heads: bad dims (H must be a multiple of 4, <= 1024)");
    MMB_REQUIRE(a->w_pooler && a->b_pooler && a->w_align && a->b_align && a->w_attn && a->b_attn && a->w_c11 && a->b_c11 &&
                    a->w_c12 && a->b_c12 && a->ap_label[0] && a->ap_label[1] && a->sentiment && a->ce_loss_sum &&
                    a->label_count,
                "heads: null parameter");
    for (int m = 0; m < 3; ++m) MMB_REQUIRE(a->w_v[m] && a->b_v[m] && a->w_cpc[m] && a->b_cpc[m], "heads: null v/cpc %d", m);
    return MMB_OK;
}

static VPtrs vptrs(const mmb_heads_args* a) {
    VPtrs v;
    for (int m = 0; m < 3; ++m) {
        v.w[m] = a->w_v[m];
        v.b[m] = a->b_v[m];
        v.gw[m] = a->g_w_v[m];
        v.gb[m] = a->g_b_v[m];
    }
    return v;
}

static LossParams loss_params(const mmb_heads_args* a, float* ws, const HeadsWs& w) {
    LossParams lp;
    lp.al = ws + w.al;
    lp.ap[0] = (const long long*)a->ap_label[0];
    lp.ap[1] = (const long long*)a->ap_label[1];
    lp.logit = ws + w.logit;
    lp.sentiment = a->sentiment;
    lp.ce_loss_sum = a->ce_loss_sum;
    lp.label_count = a->label_count;
    lp.nce = ws + w.nce;
    lp.losses = a->losses;
    lp.logits_out = a->logits_out;
    lp.dal = ws + w.dal;
    lp.dlogit = ws + w.dlogit;
    lp.gscale = a->gscale;
    lp.alpha = a->alpha;
    lp.beta = a->beta;
    lp.B = a->B;
    lp.num_labels = a->num_labels;
    return lp;
}

extern "C" int mmb_heads_fwd(const mmb_heads_args* a, void* stream) {
    int rc = heads_check(a);
    if (rc != MMB_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int B = a->B, H = a->H, R = 3 * B;
    const HeadsWs w = heads_ws(B, H);
    float* ws = (float*)a->workspace;
    const VPtrs vp = vptrs(a);

    if (a->seq_out_f32) launch_pdl(gather_cls_f32_kernel, dim3(R), dim3(256), (size_t)(0), st, (const float*)a->seq_out, a->cu_seqlens, ws + w.X0, R, H);
    else launch_pdl(gather_cls_kernel, dim3(R), dim3(256), (size_t)(0), st, (const __nv_bfloat16*)a->seq_out, a->cu_seqlens, ws + w.X0, R, H);
    linear(st, ws + w.X0, H, a->w_pooler, H, a->b_pooler, ws + w.P, H, R, H, H, ACT_TANH);
    if (a->w_seqrel && a->b_seqrel) linear(st, ws + w.P, H, a->w_seqrel, H, a->b_seqrel, ws + w.rel, 2, B, 2, H, ACT_NONE);
    linear(st, ws + w.X0 + (size_t)B * H, H, a->w_align, H, a->b_align, ws + w.al, 2, 2 * B, 2, H, ACT_NONE);
    launch_pdl(dup_cols_kernel, dim3(R), dim3(256), (size_t)(0), st, ws + w.P, ws + w.PP, R, H);
    linear(st, ws + w.PP, 2 * H, a->w_attn, 2 * H, a->b_attn, ws + w.U, H, R, H, 2 * H, ACT_RELU);
    launch_pdl(score_scale_kernel, dim3(R), dim3(256), (size_t)(0), st, ws + w.U, ws + w.P, vp, ws + w.s, ws + w.PC, B, H);
    linear(st, ws + w.PC, 3 * H, a->w_c11, 3 * H, a->b_c11, ws + w.temp, H, B, H, 3 * H, ACT_NONE);
    linear(st, ws + w.temp, H, a->w_c12, H, a->b_c12, ws + w.logit, 1, B, 1, H, ACT_NONE);
    MMB_CUDA(cudaMemsetAsync(ws + w.nce, 0, 4 * sizeof(float), st));
    for (int m = 0; m < 3; ++m) {
        float* XH = ws + w.XH + (size_t)m * B * H;
        float* xn = ws + w.xn + (size_t)m * B * H;
        float* an = ws + w.an + (size_t)m * B * H;
        linear(st, ws + w.temp, H, a->w_cpc[m], H, a->b_cpc[m], XH, H, B, H, H, ACT_NONE);
        launch_pdl(normalize_rows_kernel, dim3(B), dim3(256), (size_t)(0), st, XH, an, ws + w.na + (size_t)m * B, H);
        launch_pdl(normalize_rows_kernel, dim3(B), dim3(256), (size_t)(0), st, ws + w.P + (size_t)m * B * H, xn, ws + w.nx + (size_t)m * B, H);
        launch_pdl(cpc_rows_kernel, dim3(B), dim3(256), (size_t)(B * sizeof(float)), st, xn, an, ws + w.Sm + (size_t)m * B * B, ws + w.nce, B, H);
    }
    LossParams lp = loss_params(a, ws, w);
    launch_pdl(final_losses_kernel, dim3(1), dim3(256), (size_t)(0), st, lp);
    // user-visible score outputs
    if (a->rel_out) MMB_CUDA(cudaMemcpyAsync(a->rel_out, ws + w.rel, 2 * B * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (a->align_out) MMB_CUDA(cudaMemcpyAsync(a->align_out, ws + w.al, 4 * B * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return check_launch("heads_fwd", 21);
}

extern "C" int mmb_heads_bwd(const mmb_heads_args* a, void* stream) {
    int rc = heads_check(a);
    if (rc != MMB_OK) return rc;
    MMB_REQUIRE(a->dseq_out && a->g_w_pooler && a->g_b_pooler && a->g_w_align && a->g_b_align && a->g_w_attn && a->g_b_attn &&
                    a->g_w_c11 && a->g_b_c11 && a->g_w_c12 && a->g_b_c12,
                "heads_bwd: null gradient pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int B = a->B, H = a->H, R = 3 * B;
    const HeadsWs w = heads_ws(B, H);
    float* ws = (float*)a->workspace;
    const VPtrs vp = vptrs(a);
    LossParams lp = loss_params(a, ws, w);

    MMB_CUDA(cudaMemsetAsync(ws + w.dP, 0, (size_t)R * H * sizeof(float), st));
    MMB_CUDA(cudaMemsetAsync(ws + w.dX0, 0, (size_t)R * H * sizeof(float), st));
    launch_pdl(final_losses_bwd_kernel, dim3(1), dim3(256), (size_t)(0), st, lp);
    // classifier1_2
    dx(st, ws + w.dlogit, 1, a->w_c12, H, ws + w.dtemp, H, B, 1, H, 0);
    dw(st, ws + w.dlogit, 1, ws + w.temp, H, a->g_w_c12, H, a->g_b_c12, B, 1, H);
    // CPC x3 (joint loss has -beta * nce)
    for (int m = 0; m < 3; ++m) {
        MMB_REQUIRE(a->g_w_cpc[m] && a->g_b_cpc[m] && a->g_w_v[m] && a->g_b_v[m], "heads_bwd: null cpc/v grads");
        float* Sm = ws + w.Sm + (size_t)m * B * B;
        const float* xn = ws + w.xn + (size_t)m * B * H;
        const float* an = ws + w.an + (size_t)m * B * H;
        launch_pdl(cpc_dg_kernel, dim3((B * B + 255) / 256), dim3(256), (size_t)(0), st, Sm, a->gscale, a->beta, B);          // Sm <- dG
        launch_pdl(bb_matmul_kernel, dim3(B), dim3(256), (size_t)(B * sizeof(float)), st, Sm, an, ws + w.dxn, B, H, 0);                          // dXn = dG An
        launch_pdl(bb_matmul_kernel, dim3(B), dim3(256), (size_t)(B * sizeof(float)), st, Sm, xn, ws + w.dan, B, H, 1);                          // dAn = dG^T Xn
        launch_pdl(normalize_bwd_kernel, dim3(B), dim3(256), (size_t)(0), st, xn, ws + w.dxn, ws + w.nx + (size_t)m * B, ws + w.dP + (size_t)m * B * H, H, 1);
        launch_pdl(normalize_bwd_kernel, dim3(B), dim3(256), (size_t)(0), st, an, ws + w.dan, ws + w.na + (size_t)m * B, ws + w.dXH, H, 0);
        dx(st, ws + w.dXH, H, a->w_cpc[m], H, ws + w.dtemp, H, B, H, H, 1);
        dw(st, ws + w.dXH, H, ws + w.temp, H, a->g_w_cpc[m], H, a->g_b_cpc[m], B, H, H);
    }
    // classifier1_1
    dx(st, ws + w.dtemp, H, a->w_c11, 3 * H, ws + w.dPC, 3 * H, B, H, 3 * H, 0);
    dw(st, ws + w.dtemp, H, ws + w.PC, 3 * H, a->g_w_c11, 3 * H, a->g_b_c11, B, H, 3 * H);
    // score scaling, v_m, relu
    launch_pdl(score_scale_bwd_kernel, dim3(R), dim3(256), (size_t)(0), st, ws + w.dPC, ws + w.P, ws + w.U, ws + w.s, vp, ws + w.dP, ws + w.dA, B, H);
    // attn Linear on [P, P]
    dx(st, ws + w.dA, H, a->w_attn, 2 * H, ws + w.dPP, 2 * H, R, H, 2 * H, 0);
    dw(st, ws + w.dA, H, ws + w.PP, 2 * H, a->g_w_attn, 2 * H, a->g_b_attn, R, H, 2 * H);
    launch_pdl(fold_dup_grad_kernel, dim3(R), dim3(256), (size_t)(0), st, ws + w.dPP, ws + w.dP, R, H);
    // pooler: P = tanh(X0 Wp^T + bp)
    launch_pdl(tanh_bwd_kernel, dim3((R * H + 255) / 256), dim3(256), (size_t)(0), st, ws + w.dP, ws + w.P, R * H);
    dx(st, ws + w.dP, H, a->w_pooler, H, ws + w.dX0, H, R, H, H, 1);
    dw(st, ws + w.dP, H, ws + w.X0, H, a->g_w_pooler, H, a->g_b_pooler, R, H, H);
    // align on seq[:,0] of the two joint passes
    dx(st, ws + w.dal, 2, a->w_align, H, ws + w.dX0 + (size_t)B * H, H, 2 * B, 2, H, 1);
    dw(st, ws + w.dal, 2, ws + w.X0 + (size_t)B * H, H, a->g_w_align, H, a->g_b_align, 2 * B, 2, H);
    // add into the gradient of the encoder output at the [CLS] rows
    launch_pdl(scatter_cls_grad_kernel, dim3(R), dim3(256), (size_t)(0), st, ws + w.dX0, a->cu_seqlens, (__nv_bfloat16*)a->dseq_out, R, H);
    return check_launch("heads_bwd", 36 + 8);
}

// fp32 validation path: nn.Linear forward on fp32 activations (include/mmbert_sm100.h: mmb_linear_f32)
extern "C" int mmb_linear_f32(const mmb_linear_f32_args* a, void* stream) {
    MMB_REQUIRE(a && a->X && a->W && a->Y, "linear_f32: null pointer");
    MMB_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0 && a->ldx >= a->K && a->ldw >= a->K && a->ldy >= a->N, "linear_f32: bad shape");
    MMB_REQUIRE(a->act >= MMB_ACT_NONE && a->act <= MMB_ACT_GELU, "linear_f32: bad activation %d", a->act);
    dim3 g((a->N + kSgT - 1) / kSgT, (a->M + kSgT - 1) / kSgT);
    launch_pdl(sgemm64_kernel, dim3(g), dim3(256), (size_t)(0), (cudaStream_t)stream, a->X, (int)a->ldx, 1, a->W, 1, (int)a->ldw, a->Y, (int)a->ldy, a->bias,
                                                       a->act, a->M, a->N, a->K, 0);
    return check_launch("sgemm64_kernel");
}
