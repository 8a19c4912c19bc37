// Warp-per-row building blocks for the HBM-bound kernels (LayerNorm family, embeddings).
// A row of H elements (H % 8 == 0, H <= 8 * 32 * NCH) is held by one warp: lane l owns the
// 8-element chunks l, l+32, l+64, ... so every global access is a coalesced 16-byte (bf16) or 2x16-byte
// (f32) vector per lane.
#pragma once
#include "common.cuh"

namespace mmb {

// 32-byte global accesses (LDG.256 / STG.256, sm_100) for the fp32 rows: one instruction per 8-float chunk instead of two.
// Every fp32 row of this path starts 32-byte aligned (H % 8 == 0, buffers 256-byte aligned: torch allocations, flat store).
__device__ __forceinline__ void ldg256_f32(const float* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ void stg256_f32(float* p, const float4& a, const float4& b) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
                 "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
                 : "memory");
}

// NCH = ceil(H / 256) chunks per lane; kernels are instantiated for NCH in {1,2,3,4} (H <= 1024).
template <int NCH>
struct RowF {
    float v[NCH][8];
};

#define MMB_DISPATCH_NCH(H, ...)                         \
    do {                                                 \
        const int nch__ = ((H) + 255) / 256;             \
        if (nch__ == 1) { constexpr int NCH = 1; __VA_ARGS__; }      \
        else if (nch__ == 2) { constexpr int NCH = 2; __VA_ARGS__; } \
        else if (nch__ == 3) { constexpr int NCH = 3; __VA_ARGS__; } \
        else { constexpr int NCH = 4; __VA_ARGS__; }                 \
    } while (0)

template <int NCH>
__device__ __forceinline__ void row_zero(RowF<NCH>& r) {
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[c][i] = 0.f;
}

// bf16 row -> registers (chunks beyond H are left at zero)
template <int NCH>
__device__ __forceinline__ void row_load_bf16(RowF<NCH>& r, const __nv_bfloat16* __restrict__ p, int H, int lane) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int e = (c * 32 + lane) * 8;
        if (e < H) {
            const uint4 q = *reinterpret_cast<const uint4*>(p + e);
            const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), cc = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
            r.v[c][0] = a.x; r.v[c][1] = a.y; r.v[c][2] = b.x; r.v[c][3] = b.y;
            r.v[c][4] = cc.x; r.v[c][5] = cc.y; r.v[c][6] = d.x; r.v[c][7] = d.y;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) r.v[c][i] = 0.f;
        }
    }
}
// Split loads: row_fetch_* only ISSUES the 16-byte loads (raw vectors), row_unpack_* converts them later, so a kernel can
// put every load of a row in flight before it consumes the first one.
template <int NCH>
struct RowRawB {
    uint4 q[NCH];
};
template <int NCH>
struct RowRawF {
    float4 a[NCH], b[NCH];
};
template <int NCH>
__device__ __forceinline__ void row_fetch_bf16(RowRawB<NCH>& r, const __nv_bfloat16* __restrict__ p, int H, int lane) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int e = (c * 32 + lane) * 8;
        r.q[c] = e < H ? __ldg(reinterpret_cast<const uint4*>(p + e)) : make_uint4(0, 0, 0, 0);
    }
}
template <int NCH>
__device__ __forceinline__ void row_fetch_f32(RowRawF<NCH>& r, const float* __restrict__ p, int H, int lane) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int e = (c * 32 + lane) * 8;
        r.a[c] = r.b[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < H) ldg256_f32(p + e, r.a[c], r.b[c]);
    }
}
template <int NCH>
__device__ __forceinline__ void row_unpack_bf16(RowF<NCH>& r, const RowRawB<NCH>& raw) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const float2 a = unpack_bf16x2(raw.q[c].x), b = unpack_bf16x2(raw.q[c].y), cc = unpack_bf16x2(raw.q[c].z),
                     d = unpack_bf16x2(raw.q[c].w);
        r.v[c][0] = a.x; r.v[c][1] = a.y; r.v[c][2] = b.x; r.v[c][3] = b.y;
        r.v[c][4] = cc.x; r.v[c][5] = cc.y; r.v[c][6] = d.x; r.v[c][7] = d.y;
    }
}
template <int NCH>
__device__ __forceinline__ void row_add_f32(RowF<NCH>& r, const RowRawF<NCH>& raw) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        r.v[c][0] += raw.a[c].x; r.v[c][1] += raw.a[c].y; r.v[c][2] += raw.a[c].z; r.v[c][3] += raw.a[c].w;
        r.v[c][4] += raw.b[c].x; r.v[c][5] += raw.b[c].y; r.v[c][6] += raw.b[c].z; r.v[c][7] += raw.b[c].w;
    }
}

// kGlobal: ``p`` is known to point to GLOBAL memory (32-byte loads); the default also takes shared-memory rows
template <int NCH, bool kGlobal = false>
__device__ __forceinline__ void row_load_f32(RowF<NCH>& r, const float* __restrict__ p, int H, int lane) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int e = (c * 32 + lane) * 8;
        if (e < H) {
            float4 a, b;
            if (kGlobal) {
                ldg256_f32(p + e, a, b);
            } else {
                a = *reinterpret_cast<const float4*>(p + e);
                b = *reinterpret_cast<const float4*>(p + e + 4);
            }
            r.v[c][0] = a.x; r.v[c][1] = a.y; r.v[c][2] = a.z; r.v[c][3] = a.w;
            r.v[c][4] = b.x; r.v[c][5] = b.y; r.v[c][6] = b.z; r.v[c][7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) r.v[c][i] = 0.f;
        }
    }
}
template <int NCH>
__device__ __forceinline__ void row_store_bf16(const RowF<NCH>& r, __nv_bfloat16* __restrict__ p, int H, int lane) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int e = (c * 32 + lane) * 8;
        if (e < H) {
            uint4 q;
            q.x = pack_bf16x2(r.v[c][0], r.v[c][1]);
            q.y = pack_bf16x2(r.v[c][2], r.v[c][3]);
            q.z = pack_bf16x2(r.v[c][4], r.v[c][5]);
            q.w = pack_bf16x2(r.v[c][6], r.v[c][7]);
            *reinterpret_cast<uint4*>(p + e) = q;
        }
    }
}
template <int NCH>
__device__ __forceinline__ void row_store_f32(const RowF<NCH>& r, float* __restrict__ p, int H, int lane) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int e = (c * 32 + lane) * 8;
        if (e < H) {
            stg256_f32(p + e, make_float4(r.v[c][0], r.v[c][1], r.v[c][2], r.v[c][3]),
                       make_float4(r.v[c][4], r.v[c][5], r.v[c][6], r.v[c][7]));
        }
    }
}
template <int NCH>
__device__ __forceinline__ void row_round_bf16(RowF<NCH>& r) {
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[c][i] = bf16_round(r.v[c][i]);
}

// dropout on a row (identical in forward and backward): one hash per column pair
template <int NCH>
__device__ __forceinline__ void row_dropout(RowF<NCH>& r, int H, int lane, uint64_t seed, uint32_t stream, uint64_t row,
                                            uint32_t thresh, float inv_keep) {
    if (thresh == 0u) return;
    const uint32_t key = rng_row_key(seed, stream, (uint32_t)row);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int e = (c * 32 + lane) * 8;
        if (e < H) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t bits = rng_pair(key, (uint32_t)(e >> 1) + i);
                r.v[c][2 * i] = rng_keep_lo(bits, thresh) ? r.v[c][2 * i] * inv_keep : 0.f;
                r.v[c][2 * i + 1] = rng_keep_hi(bits, thresh) ? r.v[c][2 * i + 1] * inv_keep : 0.f;
            }
        }
    }
}

// The same decisions as row_dropout, as one bit per element of this lane (bit c * 8 + i), so that a kernel that needs
// the mask twice (backward: rebuild the LayerNorm input, then mask the gradient) hashes once.
template <int NCH>
__device__ __forceinline__ uint32_t row_dropout_mask(int H, int lane, uint64_t seed, uint32_t stream, uint64_t row,
                                                     uint32_t thresh) {
    if (thresh == 0u) return 0xFFFFFFFFu;
    const uint32_t key = rng_row_key(seed, stream, (uint32_t)row);
    uint32_t m = 0u;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int e = (c * 32 + lane) * 8;
        if (e < H) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t bits = rng_pair(key, (uint32_t)(e >> 1) + i);
                m |= (rng_keep_lo(bits, thresh) ? 1u : 0u) << (c * 8 + 2 * i);
                m |= (rng_keep_hi(bits, thresh) ? 1u : 0u) << (c * 8 + 2 * i + 1);
            }
        }
    }
    return m;
}
template <int NCH>
__device__ __forceinline__ void row_apply_mask(RowF<NCH>& r, uint32_t mask, float inv_keep) {
    if (mask == 0xFFFFFFFFu && inv_keep == 1.0f) return;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[c][i] = ((mask >> (c * 8 + i)) & 1u) ? r.v[c][i] * inv_keep : 0.f;
}

// mean / rstd of a row (two-pass in registers: exact mean first, then centred second moment)
template <int NCH>
__device__ __forceinline__ void row_stats(const RowF<NCH>& r, int H, int lane, float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) s += r.v[c][i];  // chunks beyond H hold zeros
    s = warp_sum(s);
    mean = s / (float)H;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int e = (c * 32 + lane) * 8;
        if (e < H) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float d = r.v[c][i] - mean;
                q += d * d;
            }
        }
    }
    q = warp_sum(q);
    rstd = rsqrtf(q / (float)H + eps);
}

// y = (x - mean) * rstd * gamma + beta, in place
template <int NCH>
__device__ __forceinline__ void row_affine(RowF<NCH>& r, int H, int lane, float mean, float rstd,
                                           const float* __restrict__ gamma, const float* __restrict__ beta) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int e = (c * 32 + lane) * 8;
        if (e < H) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + e));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + e + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + e));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + e + 4));
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) r.v[c][i] = (r.v[c][i] - mean) * rstd * g[i] + b[i];
        }
    }
}

// LayerNorm backward for one row.  In: x (pre-LN input), dy (grad of LN output).  Out: dx in place of dy.
// Accumulates per-lane column partials dgamma += dy * xhat, dbeta += dy.
template <int NCH>
__device__ __forceinline__ void row_ln_bwd(const RowF<NCH>& x, RowF<NCH>& dy, int H, int lane, float mean, float rstd,
                                           const float* __restrict__ gamma, RowF<NCH>& dgamma, RowF<NCH>& dbeta) {
    float s1 = 0.f, s2 = 0.f;  // sum(dy*g), sum(dy*g*xhat)
    RowF<NCH> xh;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int e = (c * 32 + lane) * 8;
        if (e < H) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + e));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + e + 4));
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float xhat = (x.v[c][i] - mean) * rstd;
                xh.v[c][i] = xhat;
                dgamma.v[c][i] += dy.v[c][i] * xhat;
                dbeta.v[c][i] += dy.v[c][i];
                const float dg = dy.v[c][i] * g[i];
                dy.v[c][i] = dg;
                s1 += dg;
                s2 += dg * xhat;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) xh.v[c][i] = 0.f;
        }
    }
    s1 = warp_sum(s1) / (float)H;
    s2 = warp_sum(s2) / (float)H;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int e = (c * 32 + lane) * 8;
        if (e < H) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dy.v[c][i] = (dy.v[c][i] - s1 - xh.v[c][i] * s2) * rstd;
        }
    }
}

// CTA-level flush of per-lane column partials into global fp32 vectors with atomics.
// smem must hold nwarps * H floats; every warp of the CTA must call.
template <int NCH>
__device__ __forceinline__ void cta_flush_columns(const RowF<NCH>& part, float* __restrict__ dst, int H, float* smem,
                                                  int warp, int lane, int nwarps) {
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int e = (c * 32 + lane) * 8;
        if (e < H) {
#pragma unroll
            for (int i = 0; i < 8; ++i) smem[warp * H + e + i] = part.v[c][i];
        }
    }
    __syncthreads();
    for (int col = threadIdx.x; col < H; col += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nwarps; ++w) s += smem[w * H + col];
        atomicAdd(dst + col, s);
    }
}

}  // namespace mmb
