// Masked self-attention for short sequences (S <= 256, head dim <= 64), fp32 SIMT path.
// The (B,1,S,S) mask of the reference is never materialised: validity is computed in-kernel from the
// (B,S) key-padding mask and the causal flag; masked scores are filled with -1e9 exactly as the
// reference does, so a fully masked query row attends uniformly to all S keys (quirk Q4).
//
// One CTA = one (batch, head) x a tile of 32 "owner" rows; K/V (or Q/dO for the key-owned backward)
// of the whole sequence are staged once in shared memory (row stride d+1 -> conflict-free for the
// lane-owns-a-row access pattern).  Each warp processes 4 owner rows at a time.
#include "common.cuh"
#include <cuda_bf16.h>

#define ATT_WARPS 4
#define ATT_RQ 4                         // rows per warp pass
#define ATT_ROWS (ATT_WARPS * ATT_RQ * 2) // 32 owner rows per CTA
#define MASK_FILL (-1e9f)

// floats taken by the two staged [S][d+1] matrices, rounded up so that what follows is 16-byte aligned
__host__ __device__ __forceinline__ size_t staged_floats(int S, int d) { return (((size_t)2 * S * (d + 1) + 3) / 4) * 4; }

struct AttnParams {
    const float* qkv;
    const uint8_t* key_valid;
    int B, S, heads, d, H, causal;
    float scale, p_drop, inv_keep;
    uint64_t seed;
    uint32_t site;
};

// out[r][i] = sum_c Xs[(lane + 32 i)][c] * Yw[c][r]   (lane-owned rows of Xs against the warp's 4 vectors)
template <int KPL>
__device__ __forceinline__ void dot_rows(const float* __restrict__ Xs, int stride, int nrows, const float* __restrict__ Yw,
                                         int d, int lane, float (&out)[ATT_RQ][KPL]) {
#pragma unroll
    for (int r = 0; r < ATT_RQ; ++r)
#pragma unroll
        for (int i = 0; i < KPL; ++i) out[r][i] = 0.f;
    for (int c = 0; c < d; ++c) {
        const float4 y = *reinterpret_cast<const float4*>(Yw + c * 4);
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            const int j = lane + 32 * i;
            const float x = j < nrows ? Xs[j * stride + c] : 0.f;
            out[0][i] = fmaf(x, y.x, out[0][i]);
            out[1][i] = fmaf(x, y.y, out[1][i]);
            out[2][i] = fmaf(x, y.z, out[2][i]);
            out[3][i] = fmaf(x, y.w, out[3][i]);
        }
    }
}

// out[r][cc] = sum_{j<n} Pw[j][r] * Zs[j][lane + 32 cc]
__device__ __forceinline__ void accum_rows(const float* __restrict__ Pw, const float* __restrict__ Zs, int stride, int n,
                                           int d, int lane, float (&out)[ATT_RQ][2]) {
#pragma unroll
    for (int r = 0; r < ATT_RQ; ++r) { out[r][0] = 0.f; out[r][1] = 0.f; }
    const bool c0 = lane < d, c1 = lane + 32 < d;
    for (int j = 0; j < n; ++j) {
        const float4 p = *reinterpret_cast<const float4*>(Pw + j * 4);
        const float z0 = c0 ? Zs[j * stride + lane] : 0.f;
        out[0][0] = fmaf(p.x, z0, out[0][0]); out[1][0] = fmaf(p.y, z0, out[1][0]);
        out[2][0] = fmaf(p.z, z0, out[2][0]); out[3][0] = fmaf(p.w, z0, out[3][0]);
        if (c1) {
            const float z1 = Zs[j * stride + lane + 32];
            out[0][1] = fmaf(p.x, z1, out[0][1]); out[1][1] = fmaf(p.y, z1, out[1][1]);
            out[2][1] = fmaf(p.z, z1, out[2][1]); out[3][1] = fmaf(p.w, z1, out[3][1]);
        }
    }
}

// stage `rows` rows of one head (column offset col0 inside the (T,3H) qkv buffer or a (T,H) buffer) into smem
__device__ __forceinline__ void stage_rows(const float* __restrict__ src, long long row0, int ld, int col0, int rows, int d,
                                           float* __restrict__ dst, int stride) {
    const int d4 = d / 4;
    for (int i = threadIdx.x; i < rows * d4; i += blockDim.x) {
        const int r = i / d4, c = (i % d4) * 4;
        const float4 v = ldg4(src + (row0 + r) * ld + col0 + c);
        float* o = dst + r * stride + c;
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
}

// the warp's 4 owner vectors, transposed to [c][4]; rows beyond `limit` are zero
__device__ __forceinline__ void stage_warp_vectors(const float* __restrict__ src, long long row0, int ld, int col0,
                                                   int first, int limit, int d, float* __restrict__ Yw, int lane) {
    for (int i = lane; i < d * ATT_RQ; i += 32) {
        const int c = i / ATT_RQ, r = i % ATT_RQ;
        const int row = first + r;
        Yw[c * 4 + r] = row < limit ? __ldg(src + (row0 + row) * ld + col0 + c) : 0.f;
    }
}

__device__ __forceinline__ bool key_ok(const AttnParams& p, int b, int q, int j) {
    bool ok = true;
    if (p.key_valid) ok = p.key_valid[(long long)b * p.S + j] != 0;
    if (p.causal) ok = ok && (j <= q);
    return ok;
}

// ---------------------------------------------------------------------------------------------
// forward (MODE 0) and dQ backward (MODE 1): query-owned
// ---------------------------------------------------------------------------------------------
template <int KPL, int MODE>
__global__ void __launch_bounds__(ATT_WARPS * 32) attn_query_kernel(AttnParams p, float* __restrict__ ctx_out,
                                                                    float* __restrict__ stats, const float* __restrict__ ctx_in,
                                                                    const float* __restrict__ d_ctx, float* __restrict__ d_qkv,
                                                                    float* __restrict__ Dbuf) {
    extern __shared__ __align__(16) float smem[];
    const int S = p.S, d = p.d, H = p.H, stride = d + 1;
    const int bh = blockIdx.x, b = bh / p.heads, h = bh % p.heads;
    const int q_tile0 = blockIdx.y * ATT_ROWS;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    // With the -1e9 fill, masked keys still take part in the softmax of a FULLY masked row (uniform over all S
    // keys), so all S keys are staged whenever a padding mask is present; causal rows without a padding mask
    // always see key 0, masked keys then contribute exactly 0 and only keys < tile end are needed.
    const int nk = (p.causal && p.key_valid == nullptr) ? min(S, q_tile0 + ATT_ROWS) : S;

    float* Ks = smem;                          // [S][d+1]
    float* Vs = Ks + (size_t)S * stride;       // [S][d+1]
    float* tail = smem + staged_floats(S, d);  // 16-byte aligned
    float* Yw = tail + warp * (2 * 64 * 4);    // per warp: q vectors [d][4] and dO vectors [d][4]
    float* Pw = tail + ATT_WARPS * (2 * 64 * 4) + warp * (ATT_RQ * 256);  // per warp: [S][4]
    const long long row0 = (long long)b * S;
    stage_rows(p.qkv, row0, 3 * H, H + h * d, nk, d, Ks, stride);
    stage_rows(p.qkv, row0, 3 * H, 2 * H + h * d, nk, d, Vs, stride);
    __syncthreads();

    for (int pass = 0; pass < 2; ++pass) {
        const int q0 = q_tile0 + (pass * ATT_WARPS + warp) * ATT_RQ;
        if (q0 >= S) continue;   // warp-uniform
        stage_warp_vectors(p.qkv, row0, 3 * H, h * d, q0, S, d, Yw, lane);
        if (MODE == 1) stage_warp_vectors(d_ctx, row0, H, h * d, q0, S, d, Yw + 64 * 4, lane);
        __syncwarp();
        float s[ATT_RQ][KPL];
        dot_rows<KPL>(Ks, stride, nk, Yw, d, lane, s);
        float mx[ATT_RQ], sm[ATT_RQ];
#pragma unroll
        for (int r = 0; r < ATT_RQ; ++r) {
            const int q = q0 + r;
            float m = -INFINITY;
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int j = lane + 32 * i;
                if (j < nk) {
                    s[r][i] = (q < S && key_ok(p, b, q, j)) ? s[r][i] * p.scale : MASK_FILL;
                    m = fmaxf(m, s[r][i]);
                } else {
                    s[r][i] = -INFINITY;
                }
            }
            if (MODE == 0) {
                m = warp_max(m);
                float e = 0.f;
#pragma unroll
                for (int i = 0; i < KPL; ++i) { s[r][i] = expf(s[r][i] - m); e += s[r][i]; }
                e = warp_sum(e);
                mx[r] = m; sm[r] = e;
                const float inv = 1.0f / e;
#pragma unroll
                for (int i = 0; i < KPL; ++i) s[r][i] *= inv;
            } else {
                const long long si = ((long long)bh * S + min(q, S - 1));
                mx[r] = stats[si];
                sm[r] = stats[(long long)p.B * p.heads * S + si];
                const float inv = 1.0f / sm[r];
#pragma unroll
                for (int i = 0; i < KPL; ++i) s[r][i] = expf(s[r][i] - mx[r]) * inv;
            }
        }
        if (MODE == 0) {
            if (stats) {
#pragma unroll
                for (int r = 0; r < ATT_RQ; ++r)
                    if (lane == r && q0 + r < S) {
                        const long long si = (long long)bh * S + q0 + r;
                        stats[si] = mx[r];
                        stats[(long long)p.B * p.heads * S + si] = sm[r];
                    }
            }
            // dropout on the probabilities, then P.V
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int j = lane + 32 * i;
                if (j < nk) {
                    float4 pv = make_float4(s[0][i], s[1][i], s[2][i], s[3][i]);
                    if (p.p_drop > 0.f) {
                        const uint64_t base = ((uint64_t)bh * S + q0) * S + j;
                        pv.x *= dropout_scale(asme_seed(p.seed), p.site, base, p.p_drop, p.inv_keep);
                        pv.y *= dropout_scale(asme_seed(p.seed), p.site, base + S, p.p_drop, p.inv_keep);
                        pv.z *= dropout_scale(asme_seed(p.seed), p.site, base + 2 * (uint64_t)S, p.p_drop, p.inv_keep);
                        pv.w *= dropout_scale(asme_seed(p.seed), p.site, base + 3 * (uint64_t)S, p.p_drop, p.inv_keep);
                    }
                    *reinterpret_cast<float4*>(Pw + j * 4) = pv;
                }
            }
            __syncwarp();
            float o[ATT_RQ][2];
            accum_rows(Pw, Vs, stride, nk, d, lane, o);
#pragma unroll
            for (int r = 0; r < ATT_RQ; ++r) {
                const int q = q0 + r;
                if (q < S) {
                    if (lane < d) ctx_out[(row0 + q) * H + h * d + lane] = o[r][0];
                    if (lane + 32 < d) ctx_out[(row0 + q) * H + h * d + lane + 32] = o[r][1];
                }
            }
            __syncwarp();
        } else {
            // D_r = dO_r . ctx_r
            float D[ATT_RQ];
#pragma unroll
            for (int r = 0; r < ATT_RQ; ++r) {
                const int q = q0 + r;
                float acc = 0.f;
                if (q < S) {
                    for (int c = lane; c < d; c += 32)
                        acc += __ldg(d_ctx + (row0 + q) * H + h * d + c) * __ldg(ctx_in + (row0 + q) * H + h * d + c);
                }
                D[r] = warp_sum(acc);
                if (lane == 0 && q < S) Dbuf[(long long)bh * S + q] = D[r];
            }
            float dp[ATT_RQ][KPL];
            dot_rows<KPL>(Vs, stride, nk, Yw + 64 * 4, d, lane, dp);
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int j = lane + 32 * i;
                if (j < nk) {
                    float ds[ATT_RQ];
#pragma unroll
                    for (int r = 0; r < ATT_RQ; ++r) {
                        const int q = q0 + r;
                        float g = dp[r][i];
                        if (p.p_drop > 0.f)
                            g *= dropout_scale(asme_seed(p.seed), p.site, ((uint64_t)bh * S + q) * S + j, p.p_drop, p.inv_keep);
                        const bool ok = q < S && key_ok(p, b, q, j);
                        ds[r] = ok ? s[r][i] * (g - D[r]) * p.scale : 0.f;
                    }
                    *reinterpret_cast<float4*>(Pw + j * 4) = make_float4(ds[0], ds[1], ds[2], ds[3]);
                }
            }
            __syncwarp();
            float o[ATT_RQ][2];
            accum_rows(Pw, Ks, stride, nk, d, lane, o);
#pragma unroll
            for (int r = 0; r < ATT_RQ; ++r) {
                const int q = q0 + r;
                if (q < S) {
                    if (lane < d) d_qkv[(row0 + q) * 3 * H + h * d + lane] = o[r][0];
                    if (lane + 32 < d) d_qkv[(row0 + q) * 3 * H + h * d + lane + 32] = o[r][1];
                }
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// dK / dV backward: key-owned; lanes own queries
// ---------------------------------------------------------------------------------------------
template <int KPL>
__global__ void __launch_bounds__(ATT_WARPS * 32) attn_key_kernel(AttnParams p, const float* __restrict__ stats,
                                                                  const float* __restrict__ d_ctx, const float* __restrict__ Dbuf,
                                                                  float* __restrict__ d_qkv) {
    extern __shared__ __align__(16) float smem[];
    const int S = p.S, d = p.d, H = p.H, stride = d + 1;
    const int bh = blockIdx.x, b = bh / p.heads, h = bh % p.heads;
    const int k_tile0 = blockIdx.y * ATT_ROWS;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

    float* Qs = smem;                           // [S][d+1]
    float* Gs = Qs + (size_t)S * stride;        // dO [S][d+1]
    float* tail = smem + staged_floats(S, d);   // 16-byte aligned
    float* Yw = tail + warp * (2 * 64 * 4);     // k vectors [d][4], v vectors [d][4]
    float* Pw = tail + ATT_WARPS * (2 * 64 * 4) + warp * (2 * ATT_RQ * 256);  // pd [S][4], ds [S][4]
    float* Rs = tail + ATT_WARPS * (2 * 64 * 4) + ATT_WARPS * (2 * ATT_RQ * 256);  // m, 1/sum, D [3][S]
    const long long row0 = (long long)b * S;
    stage_rows(p.qkv, row0, 3 * H, h * d, S, d, Qs, stride);
    stage_rows(d_ctx, row0, H, h * d, S, d, Gs, stride);
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        Rs[i] = stats[(long long)bh * S + i];
        Rs[S + i] = 1.0f / stats[(long long)p.B * p.heads * S + (long long)bh * S + i];
        Rs[2 * S + i] = Dbuf[(long long)bh * S + i];
    }
    __syncthreads();

    for (int pass = 0; pass < 2; ++pass) {
        const int j0 = k_tile0 + (pass * ATT_WARPS + warp) * ATT_RQ;
        if (j0 >= S) continue;
        stage_warp_vectors(p.qkv, row0, 3 * H, H + h * d, j0, S, d, Yw, lane);
        stage_warp_vectors(p.qkv, row0, 3 * H, 2 * H + h * d, j0, S, d, Yw + 64 * 4, lane);
        __syncwarp();
        float s[ATT_RQ][KPL], dp[ATT_RQ][KPL];
        dot_rows<KPL>(Qs, stride, S, Yw, d, lane, s);           // s[r][i] = q_i . k_{j0+r}
        dot_rows<KPL>(Gs, stride, S, Yw + 64 * 4, d, lane, dp); // dp[r][i] = dO_i . v_{j0+r}
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            const int q = lane + 32 * i;
            if (q < S) {
                float pd[ATT_RQ], ds[ATT_RQ];
                const float m = Rs[q], inv = Rs[S + q], D = Rs[2 * S + q];
#pragma unroll
                for (int r = 0; r < ATT_RQ; ++r) {
                    const int j = j0 + r;
                    const bool in = j < S;
                    const bool ok = in && key_ok(p, b, q, j);
                    const float sc = ok ? s[r][i] * p.scale : MASK_FILL;
                    const float prob = in ? expf(sc - m) * inv : 0.f;
                    float dscale = 1.f;
                    if (p.p_drop > 0.f && in)
                        dscale = dropout_scale(asme_seed(p.seed), p.site, ((uint64_t)bh * S + q) * S + j, p.p_drop, p.inv_keep);
                    pd[r] = prob * dscale;
                    ds[r] = ok ? prob * (dp[r][i] * dscale - D) * p.scale : 0.f;
                }
                *reinterpret_cast<float4*>(Pw + q * 4) = make_float4(pd[0], pd[1], pd[2], pd[3]);
                *reinterpret_cast<float4*>(Pw + ATT_RQ * 256 + q * 4) = make_float4(ds[0], ds[1], ds[2], ds[3]);
            }
        }
        __syncwarp();
        float dv[ATT_RQ][2], dk[ATT_RQ][2];
        accum_rows(Pw, Gs, stride, S, d, lane, dv);
        accum_rows(Pw + ATT_RQ * 256, Qs, stride, S, d, lane, dk);
#pragma unroll
        for (int r = 0; r < ATT_RQ; ++r) {
            const int j = j0 + r;
            if (j < S) {
                float* base = d_qkv + (row0 + j) * 3 * H + h * d;
                if (lane < d) { base[H + lane] = dk[r][0]; base[2 * H + lane] = dv[r][0]; }
                if (lane + 32 < d) { base[H + lane + 32] = dk[r][1]; base[2 * H + lane + 32] = dv[r][1]; }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static size_t attn_smem_bytes(int S, int d, int key_owned) {
    size_t f = staged_floats(S, d) + ATT_WARPS * (2 * 64 * 4);
    f += key_owned ? (size_t)ATT_WARPS * 2 * ATT_RQ * 256 + 3 * (size_t)S : (size_t)ATT_WARPS * ATT_RQ * 256;
    return f * sizeof(float);
}

static int check_attn(int B, int S, int heads, int d, float p_drop) {
    ASME_REQUIRE(B >= 0 && S >= 1 && S <= 256, "attention: S=%d unsupported (1..256)", S);
    ASME_REQUIRE(heads >= 1 && d >= 4 && d <= 64 && d % 4 == 0, "attention: head dim d=%d unsupported (4..64, multiple of 4)", d);
    ASME_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "attention: dropout p=%f out of range", p_drop);
    return ASME_OK;
}

static AttnParams make_params(const float* qkv, const uint8_t* key_valid, int B, int S, int heads, int d, int causal,
                              float p_drop, uint64_t seed, uint32_t site) {
    AttnParams p;
    p.qkv = qkv; p.key_valid = key_valid; p.B = B; p.S = S; p.heads = heads; p.d = d; p.H = heads * d; p.causal = causal;
    p.scale = 1.0f / sqrtf((float)d); p.p_drop = p_drop; p.inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    p.seed = seed; p.site = site;
    return p;
}

#define DISPATCH_KPL(S, CALL)                 \
    if ((S) <= 32) { CALL(1); }               \
    else if ((S) <= 64) { CALL(2); }          \
    else if ((S) <= 128) { CALL(4); }         \
    else { CALL(8); }

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    { const int _rc = asme_ensure_max_smem((const void*)kernel); if (_rc) return _rc; }
    return ASME_OK;
}

extern "C" int asme_b200_attn_fwd(const float* qkv, const uint8_t* key_valid, int B, int S, int heads, int d, int causal,
                                  float p_drop, uint64_t seed, uint32_t site, float* ctx, float* stats,
                                  asme_stream_t stream) {
    ASME_REQUIRE(qkv && ctx, "attn_fwd: null argument");
    int rc = check_attn(B, S, heads, d, p_drop);
    if (rc) return rc;
    if (B == 0) return ASME_OK;
    AttnParams p = make_params(qkv, key_valid, B, S, heads, d, causal, p_drop, seed, site);
    const size_t smem = attn_smem_bytes(S, d, 0);
    dim3 grid(B * heads, ceil_div(S, ATT_ROWS));
#define CALL(KPL)                                                                                              \
    {                                                                                                          \
        rc = set_smem(attn_query_kernel<KPL, 0>, smem);                                                        \
        if (rc) return rc;                                                                                     \
        attn_query_kernel<KPL, 0><<<grid, ATT_WARPS * 32, smem, (cudaStream_t)stream>>>(p, ctx, stats, nullptr, \
                                                                                       nullptr, nullptr, nullptr); \
    }
    DISPATCH_KPL(S, CALL)
#undef CALL
    ASME_LAUNCH_OK();
    return ASME_OK;
}

extern "C" size_t asme_b200_attn_bwd_workspace_bytes(int B, int S, int heads) {
    return (size_t)B * heads * S * sizeof(float);
}

extern "C" int asme_b200_attn_bwd(const float* qkv, const uint8_t* key_valid, int B, int S, int heads, int d, int causal,
                                  float p_drop, uint64_t seed, uint32_t site, const float* ctx, const float* d_ctx,
                                  const float* stats, float* d_qkv, void* ws, size_t ws_bytes, asme_stream_t stream) {
    ASME_REQUIRE(qkv && ctx && d_ctx && stats && d_qkv, "attn_bwd: null argument");
    int rc = check_attn(B, S, heads, d, p_drop);
    if (rc) return rc;
    if (B == 0) return ASME_OK;
    if (ws_bytes < asme_b200_attn_bwd_workspace_bytes(B, S, heads)) {
        asme_set_error("attn_bwd: workspace too small");
        return ASME_ERR_WORKSPACE;
    }
    float* Dbuf = (float*)ws;
    AttnParams p = make_params(qkv, key_valid, B, S, heads, d, causal, p_drop, seed, site);
    const size_t smem_q = attn_smem_bytes(S, d, 0), smem_k = attn_smem_bytes(S, d, 1);
    dim3 grid(B * heads, ceil_div(S, ATT_ROWS));
#define CALL(KPL)                                                                                                      \
    {                                                                                                                  \
        rc = set_smem(attn_query_kernel<KPL, 1>, smem_q);                                                              \
        if (rc) return rc;                                                                                             \
        attn_query_kernel<KPL, 1><<<grid, ATT_WARPS * 32, smem_q, (cudaStream_t)stream>>>(p, nullptr, (float*)stats, ctx, \
                                                                                         d_ctx, d_qkv, Dbuf);          \
        rc = set_smem(attn_key_kernel<KPL>, smem_k);                                                                   \
        if (rc) return rc;                                                                                             \
        attn_key_kernel<KPL><<<grid, ATT_WARPS * 32, smem_k, (cudaStream_t)stream>>>(p, stats, d_ctx, Dbuf, d_qkv);    \
    }
    DISPATCH_KPL(S, CALL)
#undef CALL
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ---------------------------------------------------------------------------------------------
// One query position per sequence (evaluation, last encoder layer): the attention of row only_row[b] is two matrix-vector
// products over that sequence's K and V rows -- HBM-bound (every K/V byte is read once, 128 bits per thread, coalesced), no
// tensor-core tile to fill.  Same semantics as Attention.forward (transformer_layers.py:145-155): -1e9 fill, softmax over all
// S keys (a row without any valid key attends uniformly), fp32 arithmetic on the bf16 q/k/v, context rounded to bf16.
// Thread layout: thread = (8-column slice of the H-wide row, row group); both phases walk rows group, group + G, ...
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8_bf16(const uint4 u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

__global__ void __launch_bounds__(256) attn_row_kernel(const __nv_bfloat16* __restrict__ qkv, const uint8_t* __restrict__ key_valid,
                                                       int S, int heads, int d, int causal, const int64_t* __restrict__ only_row,
                                                       __nv_bfloat16* __restrict__ ctx_rows) {
    extern __shared__ float row_smem[];
    const int H = heads * d;
    float* score = row_smem;                       // heads * S (scores, then unnormalised probabilities)
    float* red = score + heads * S;                // G * H partial contexts
    float* inv_l = red + 256 * 8;                  // heads
    const int b = blockIdx.x;
    const int slices = H / 8;                      // threads per row
    const int G = 256 / slices;                    // row groups
    const int c8 = threadIdx.x % slices, g = threadIdx.x / slices;
    const int per_head = d / 8;                    // consecutive threads that share a head (power of two <= 32)
    const int h = c8 / per_head;
    long long sq = only_row[b] - (long long)b * S;
    sq = sq < 0 ? 0 : (sq >= S ? S - 1 : sq);
    const __nv_bfloat16* base = qkv + (size_t)b * S * 3 * H;
    float q[8];
    unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(base + (size_t)sq * 3 * H + c8 * 8)), q);
    const float scale = rsqrtf((float)d);
    // ---- scores: partial dot over this thread's 8 columns, reduced over the head's per_head threads
    for (int j0 = 0; j0 < S; j0 += 4 * G) {        // four independent 128-bit loads in flight per thread
        uint4 kr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * G + g;
            kr[u] = j < S ? __ldg(reinterpret_cast<const uint4*>(base + (size_t)j * 3 * H + H + c8 * 8)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * G + g;
            float k[8];
            unpack8_bf16(kr[u], k);
            float part = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) part = fmaf(q[i], k[i], part);
            for (int o = per_head >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (j < S && (c8 % per_head) == 0) {
                const bool ok = (!key_valid || key_valid[(size_t)b * S + j]) && (!causal || j <= sq);
                score[h * S + j] = ok ? part * scale : -1e9f;
            }
        }
    }
    __syncthreads();
    // ---- softmax statistics: one warp per head
    for (int hh = threadIdx.x / 32; hh < heads; hh += 8) {
        const int lane = threadIdx.x % 32;
        float m = -INFINITY;
        for (int j = lane; j < S; j += 32) m = fmaxf(m, score[hh * S + j]);
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float l = 0.f;
        for (int j = lane; j < S; j += 32) {
            const float p = expf(score[hh * S + j] - m);
            score[hh * S + j] = p;
            l += p;
        }
        for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
        if (lane == 0) inv_l[hh] = 1.0f / l;
    }
    __syncthreads();
    // ---- context: sum_j p_j v_j over this thread's rows, then across the row groups
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int j0 = g; j0 < S; j0 += 4 * G) {
        uint4 vr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * G;
            vr[u] = j < S ? __ldg(reinterpret_cast<const uint4*>(base + (size_t)j * 3 * H + 2 * H + c8 * 8)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * G;
            float v[8];
            unpack8_bf16(vr[u], v);
            const float p = j < S ? score[h * S + j] : 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(p, v[i], acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[(size_t)g * H + c8 * 8 + i] = acc[i];
    __syncthreads();
    if (g == 0) {
        const float il = inv_l[h];
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            float a0 = 0.f, a1 = 0.f;
            for (int gg = 0; gg < G; ++gg) {           // fixed order: deterministic
                a0 += red[(size_t)gg * H + c8 * 8 + i];
                a1 += red[(size_t)gg * H + c8 * 8 + i + 1];
            }
            __nv_bfloat162 pk = __floats2bfloat162_rn(a0 * il, a1 * il);
            w[i / 2] = *reinterpret_cast<uint32_t*>(&pk);
        }
        *reinterpret_cast<uint4*>(ctx_rows + (size_t)b * H + c8 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

extern "C" int asme_b200_attn_row_fwd(const void* qkv, const uint8_t* key_valid, int B, int S, int heads, int d, int causal,
                                      const int64_t* only_row, void* ctx_rows, asme_stream_t stream) {
    ASME_REQUIRE(qkv && only_row && ctx_rows, "attn_row_fwd: null argument");
    ASME_REQUIRE(B >= 0 && S > 0 && heads > 0, "attn_row_fwd: bad shape B=%d S=%d heads=%d", B, S, heads);
    ASME_REQUIRE(d % 8 == 0 && d >= 8 && d <= 256 && (d & (d - 1)) == 0, "attn_row_fwd: head size d=%d (supported: 8..256, power of two)", d);
    const int H = heads * d;
    ASME_REQUIRE(H <= 2048 && 256 % (H / 8) == 0, "attn_row_fwd: hidden size H=%d (H/8 must divide 256)", H);
    const size_t smem = ((size_t)heads * S + 256 * 8 + heads) * sizeof(float);
    ASME_REQUIRE(smem <= 200 * 1024, "attn_row_fwd: heads*S=%d does not fit shared memory", heads * S);
    if (B == 0) return ASME_OK;
    int rc = set_smem(attn_row_kernel, smem);
    if (rc) return rc;
    attn_row_kernel<<<B, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv, key_valid, S, heads, d, causal, only_row,
                                                           (__nv_bfloat16*)ctx_rows);
    ASME_LAUNCH_OK();
    return ASME_OK;
}
