// Row-wise HBM-bound kernels: fused embedding gather-and-sum (+LayerNorm +dropout), LayerNorm
// forward/backward, column sums, positional-gradient reduce, elementwise helpers.
//
// Thread mapping (all kernels): a row of H fp32 is owned by an aligned group of LANES lanes
// (LANES = min(32, H/4)); lane l of the group owns the 128-bit column chunks (c*LANES + l)*4..+3,
// c < CH = H/(4*LANES).  Every global access is a coalesced float4 (128-bit) transaction and the
// LayerNorm reductions are xor-shuffles inside the group -- no shared memory on the forward path.
#include "common.cuh"
#include <cuda_bf16.h>

#include <stdarg.h>
#include <string.h>

// ---------------------------------------------------------------------------------------------
// error string (thread local) -- shared by all translation units
// ---------------------------------------------------------------------------------------------
static thread_local char g_last_error[512] = "";
void asme_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}
extern "C" const char* asme_b200_last_error(void) { return g_last_error; }
#include <atomic>
static std::atomic<long long> g_launches{0};
void asme_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" long long asme_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" int asme_b200_abi_version(void) { return 3; }

#include <mutex>
int asme_ensure_max_smem(const void* kernel) {
    static const void* done[256];
    static int n_done = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < n_done; ++i)
        if (done[i] == kernel) return ASME_OK;
    ASME_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ASME_MAX_DYN_SMEM));
    if (n_done < 256) done[n_done++] = kernel;
    return ASME_OK;
}

#define LN_EPS 1e-5f

// dropout scales for 4 consecutive elements starting at idx (idx % 4 == 0): one generator call
__device__ __forceinline__ float4 dropout_scale4(uint64_t seed, uint32_t site, uint64_t idx, float p, float inv_keep) {
    float s[4];
    dropout_scales4(seed, site, idx >> 2, p, inv_keep, s);
    return make_float4(s[0], s[1], s[2], s[3]);
}

template <int LANES, int CH>
struct Row {
    float4 v[CH];
    __device__ __forceinline__ void load(const float* base, int lane) {
#pragma unroll
        for (int c = 0; c < CH; ++c) v[c] = ldg4(base + (c * LANES + lane) * 4);
    }
    __device__ __forceinline__ void add(const float* base, int lane) {
#pragma unroll
        for (int c = 0; c < CH; ++c) add4(v[c], ldg4(base + (c * LANES + lane) * 4));
    }
    __device__ __forceinline__ void store(float* base, int lane) const {
#pragma unroll
        for (int c = 0; c < CH; ++c) *reinterpret_cast<float4*>(base + (c * LANES + lane) * 4) = v[c];
    }
    __device__ __forceinline__ void store_bf16(__nv_bfloat16* base, int lane) const {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v[c].x, v[c].y), hi = __floats2bfloat162_rn(v[c].z, v[c].w);
            uint2 o;
            o.x = *reinterpret_cast<const uint32_t*>(&lo);
            o.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(base + (c * LANES + lane) * 4) = o;
        }
    }
    __device__ __forceinline__ float sum() const {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) s += (v[c].x + v[c].y) + (v[c].z + v[c].w);
        return group_sum<LANES>(s);
    }
};

// y = (x - mean) * rstd * gamma + beta ; returns xhat in `x`, y in `y`
template <int LANES, int CH>
__device__ __forceinline__ void ln_forward(Row<LANES, CH>& x, Row<LANES, CH>& y, const float* gamma, const float* beta,
                                           int lane, int H, float& mean, float& rstd) {
    const float inv_h = 1.0f / (float)H;      // H is a power of two (DISPATCH_H): v * (1/H) == v / H bit for bit, without the division sequence per row
    mean = x.sum() * inv_h;
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        x.v[c].x -= mean; x.v[c].y -= mean; x.v[c].z -= mean; x.v[c].w -= mean;
        sq += (x.v[c].x * x.v[c].x + x.v[c].y * x.v[c].y) + (x.v[c].z * x.v[c].z + x.v[c].w * x.v[c].w);
    }
    sq = group_sum<LANES>(sq);
    rstd = rsqrtf(sq * inv_h + LN_EPS);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const float4 g = ldg4(gamma + (c * LANES + lane) * 4);
        const float4 b = ldg4(beta + (c * LANES + lane) * 4);
        x.v[c].x *= rstd; x.v[c].y *= rstd; x.v[c].z *= rstd; x.v[c].w *= rstd;
        y.v[c].x = x.v[c].x * g.x + b.x; y.v[c].y = x.v[c].y * g.y + b.y;
        y.v[c].z = x.v[c].z * g.z + b.z; y.v[c].w = x.v[c].w * g.w + b.w;
    }
}

// g (in: dL/dy, out: dL/dx) ; xhat given; accumulates dgamma/dbeta into per-thread registers
template <int LANES, int CH>
__device__ __forceinline__ void ln_backward(Row<LANES, CH>& g, const Row<LANES, CH>& xhat, const float* gamma, int lane,
                                            int H, float rstd, Row<LANES, CH>& dgamma, Row<LANES, CH>& dbeta) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const float4 gm = ldg4(gamma + (c * LANES + lane) * 4);
        float4& gv = g.v[c];
        const float4& xh = xhat.v[c];
        dgamma.v[c].x += gv.x * xh.x; dgamma.v[c].y += gv.y * xh.y; dgamma.v[c].z += gv.z * xh.z; dgamma.v[c].w += gv.w * xh.w;
        add4(dbeta.v[c], gv);
        gv.x *= gm.x; gv.y *= gm.y; gv.z *= gm.z; gv.w *= gm.w;
        s1 += (gv.x + gv.y) + (gv.z + gv.w);
        s2 += (gv.x * xh.x + gv.y * xh.y) + (gv.z * xh.z + gv.w * xh.w);
    }
    s1 = group_sum<LANES>(s1) / (float)H;
    s2 = group_sum<LANES>(s2) / (float)H;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        float4& gv = g.v[c];
        const float4& xh = xhat.v[c];
        gv.x = rstd * (gv.x - s1 - xh.x * s2); gv.y = rstd * (gv.y - s1 - xh.y * s2);
        gv.z = rstd * (gv.z - s1 - xh.z * s2); gv.w = rstd * (gv.w - s1 - xh.w * s2);
    }
}

template <int LANES, int CH>
__device__ __forceinline__ void apply_dropout(Row<LANES, CH>& x, uint64_t seed, uint32_t site, long long row, int H,
                                              int lane, float p, float inv_keep) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const float4 s = dropout_scale4(seed, site, (uint64_t)row * H + (c * LANES + lane) * 4, p, inv_keep);
        x.v[c].x *= s.x; x.v[c].y *= s.y; x.v[c].z *= s.z; x.v[c].w *= s.w;
    }
}

template <int LANES, int CH>
__device__ __forceinline__ void zero_row(Row<LANES, CH>& r) {
#pragma unroll
    for (int c = 0; c < CH; ++c) r.v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// deterministic block-level reduction of per-group register partials into out[block][slot][H]
template <int LANES, int CH>
__device__ __forceinline__ void block_reduce_rows(const Row<LANES, CH>& r, float* smem /* groups*H */, float* out, int H,
                                                  int group_in_block, int groups_per_block, int lane) {
    __syncthreads();
    r.store(smem + (size_t)group_in_block * H, lane);
    __syncthreads();
    for (int col = threadIdx.x; col < H; col += blockDim.x) {
        float s = 0.f;
        for (int g = 0; g < groups_per_block; ++g) s += smem[(size_t)g * H + col];
        out[col] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// fused embedding forward
// ---------------------------------------------------------------------------------------------
template <int LANES, int CH>
__device__ __forceinline__ void embed_gather_attrs(Row<LANES, CH>& x, const asme_embed_desc& d, long long t, int H, int lane) {
    for (int a = 0; a < d.n_attr; ++a) {
        const long long id = __ldg(d.attr_ids[a] + t);
        x.add(d.attr_table[a] + id * H, lane);
    }
    for (int b = 0; b < d.n_bag; ++b) {
        x.add(d.bag_bias[b], lane);
        const int w = d.bag_width[b];
        for (int j = 0; j < w; ++j) {
            const long long id = __ldg(d.bag_ids[b] + t * w + j);
            if (id != 0) x.add(d.bag_table_t[b] + id * H, lane);   // group-uniform branch
        }
    }
}

// TOK tokens per lane group: all ids, then all item rows are requested before anything is consumed, so every group keeps
// TOK independent 128-bit gathers in flight (a random 256-512 B row per token is latency-bound otherwise).  A block covers
// TOK * groups consecutive tokens; token j of group g is base + j * groups + g, so stores stay coalesced across groups.
template <int LANES, int CH, int TOK, bool PRE>
__global__ void __launch_bounds__(256, (!PRE && CH == 1) ? 5 : ((!PRE && TOK == 1) ? 4 : 1)) embed_fwd_kernel(const asme_embed_desc d, int T, int S, int H, float* __restrict__ out,
                                                        float* __restrict__ stats) {
    const int lane = threadIdx.x % LANES;
    const int groups = blockDim.x / LANES;
    const long long base = (long long)blockIdx.x * (groups * TOK) + threadIdx.x / LANES;
    const float inv_keep = d.p_drop > 0.f ? 1.0f / (1.0f - d.p_drop) : 1.0f;
    // user prefix (UBERT4Rec / UserSASRec): position 0 of every sequence is the sum of the user-attribute embeddings and the
    // items follow at positions 1..S-1, i.e. output row (b, s) reads item token b*(S-1) + s-1.  pre = 0: plain layout.
    // PRE is a template parameter: the plain layout keeps the short id -> row address chain (and its register count)
    constexpr int pre = PRE ? 1 : 0;
    const float* first[TOK];
    int bs[TOK], sps[TOK];
    Row<LANES, CH> xs[TOK];
#pragma unroll
    for (int j = 0; j < TOK; ++j) {
        const long long tj = base + (long long)j * groups;
        const long long t = tj < T ? tj : 0;
        if (PRE) {
            // token index -> (sequence, position) in 32-bit arithmetic (T is an int; a 64-bit division costs more than the gather)
            const unsigned b = (unsigned)t / (unsigned)S;
            const int sp = (int)((unsigned)t - b * (unsigned)S);
            bs[j] = (int)b; sps[j] = sp;
            if (sp == 0) first[j] = d.user_table[0] + __ldg(d.user_ids[0] + b) * H;
            else first[j] = d.item_table + __ldg(d.item_ids + (long long)b * (S - 1) + sp - 1) * H;
        } else {
            first[j] = d.item_table + __ldg(d.item_ids + t) * H;
        }
    }
#pragma unroll
    for (int j = 0; j < TOK; ++j) xs[j].load(first[j], lane);
    // no early exit below: the TOK iterations stay independent straight-line code, so their shuffle / LayerNorm chains interleave
#pragma unroll
    for (int j = 0; j < TOK; ++j) {
        const long long tj = base + (long long)j * groups;
        const bool ok = tj < T;
        const long long t = ok ? tj : 0;
        const long long b = PRE ? bs[j] : 0;
        const int sp = PRE ? sps[j] : (d.pos_table ? (int)((unsigned)t % (unsigned)S) : 0);
        const bool is_user = PRE && sp == 0;                  // uniform inside the lane group
        const long long src = PRE ? b * (S - 1) + sp - 1 : t;  // item token feeding this row (unused for the user row)
        Row<LANES, CH> x = xs[j], y;
        if (is_user) {
            for (int u = 1; u < d.n_user; ++u) x.add(d.user_table[u] + __ldg(d.user_ids[u] + b) * H, lane);
        } else {
            if (d.pos_table) x.add(d.pos_table + (long long)(sp - pre) * H, lane);
            if (d.ln1_gamma) {
                float mean, rstd;
                ln_forward<LANES, CH>(x, y, d.ln1_gamma, d.ln1_beta, lane, H, mean, rstd);
                if (stats && lane == 0 && ok) { stats[t] = mean; stats[(size_t)T + t] = rstd; }
                x = y;
                if (d.p_drop > 0.f) apply_dropout<LANES, CH>(x, asme_seed(d.seed), d.site_a, t, H, lane, d.p_drop, inv_keep);
            }
            if (d.n_attr | d.n_bag) embed_gather_attrs<LANES, CH>(x, d, src, H, lane);
        }
        if (d.seg_table) x.add(d.seg_table + (is_user ? 0 : H), lane);     // segment 0 = user token, 1 = items
        if (d.ln2_gamma) {
            float mean, rstd;
            ln_forward<LANES, CH>(x, y, d.ln2_gamma, d.ln2_beta, lane, H, mean, rstd);
            if (stats && lane == 0 && ok) { stats[(size_t)2 * T + t] = mean; stats[(size_t)3 * T + t] = rstd; }
            x = y;
            if (d.p_drop > 0.f) apply_dropout<LANES, CH>(x, asme_seed(d.seed), d.site_b, t, H, lane, d.p_drop, inv_keep);
        }
        if (ok) x.store(out + t * H, lane);
        if (d.next_gamma) {      // the first encoder block's input LayerNorm rides along: y16 = LN(x) as bf16 (+ row statistics)
            float mean, rstd;
            ln_forward<LANES, CH>(x, y, d.next_gamma, d.next_beta, lane, H, mean, rstd);
            if (ok) {
                if (d.next_stats && lane == 0) { d.next_stats[t] = mean; d.next_stats[(size_t)T + t] = rstd; }
                y.store_bf16(reinterpret_cast<__nv_bfloat16*>(d.next_out) + t * H, lane);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// basket inputs (N,S,BS): item embeddings of a step pooled over the basket (models/common/layers/sequence_embedding.py:9-45,
// :83-93: max / sum / mean over dim -2, PAD slots take part with E[pad] like any other id).  The dense fallback of SURVEY.md 2.1 #3:
// the pooled rows (T,H) then enter the fused embedding kernel as a table indexed by the token number.
//   mode 0 = sum, 1 = mean, 2 = max (arg (T,H) uint8 = basket slot of the maximum, first slot on ties as torch.max reports it)
// ---------------------------------------------------------------------------------------------
template <int LANES, int CH>
__global__ void __launch_bounds__(256) embed_pool_fwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ table, int T,
                                                             int BS, int H, int mode, float* __restrict__ out,
                                                             uint8_t* __restrict__ arg) {
    const int lane = threadIdx.x % LANES;
    const long long t = (long long)blockIdx.x * (blockDim.x / LANES) + threadIdx.x / LANES;
    if (t >= T) return;
    Row<LANES, CH> acc;
    int am[CH * 4];
    acc.load(table + __ldg(ids + t * BS) * H, lane);
#pragma unroll
    for (int e = 0; e < CH * 4; ++e) am[e] = 0;
    for (int j = 1; j < BS; ++j) {
        Row<LANES, CH> r;
        r.load(table + __ldg(ids + t * BS + j) * H, lane);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (mode == 2) {
                if (r.v[c].x > acc.v[c].x) { acc.v[c].x = r.v[c].x; am[4 * c] = j; }
                if (r.v[c].y > acc.v[c].y) { acc.v[c].y = r.v[c].y; am[4 * c + 1] = j; }
                if (r.v[c].z > acc.v[c].z) { acc.v[c].z = r.v[c].z; am[4 * c + 2] = j; }
                if (r.v[c].w > acc.v[c].w) { acc.v[c].w = r.v[c].w; am[4 * c + 3] = j; }
            } else {
                add4(acc.v[c], r.v[c]);
            }
        }
    }
    if (mode == 1) {
        const float inv = 1.0f / (float)BS;           // torch.mean: sum / count
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            acc.v[c].x = acc.v[c].x / (float)BS; acc.v[c].y = acc.v[c].y / (float)BS;
            acc.v[c].z = acc.v[c].z / (float)BS; acc.v[c].w = acc.v[c].w / (float)BS;
        }
        (void)inv;
    }
    acc.store(out + t * H, lane);
    if (mode == 2 && arg != nullptr) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            uchar4 a = make_uchar4((unsigned char)am[4 * c], (unsigned char)am[4 * c + 1], (unsigned char)am[4 * c + 2], (unsigned char)am[4 * c + 3]);
            *reinterpret_cast<uchar4*>(arg + t * H + (c * LANES + lane) * 4) = a;
        }
    }
}
// gradient rows of every (token, slot): d_rows[t*BS + j] = d_out[t] (sum), d_out[t] / BS (mean), d_out[t] where slot j held the maximum
template <int LANES, int CH>
__global__ void __launch_bounds__(256) embed_pool_bwd_kernel(const float* __restrict__ d_out, const uint8_t* __restrict__ arg, int T,
                                                             int BS, int H, int mode, float* __restrict__ d_rows) {
    const int lane = threadIdx.x % LANES;
    const long long t = (long long)blockIdx.x * (blockDim.x / LANES) + threadIdx.x / LANES;
    if (t >= T) return;
    Row<LANES, CH> d;
    d.load(d_out + t * H, lane);
    uchar4 am[CH];
    if (mode == 2) {
#pragma unroll
        for (int c = 0; c < CH; ++c) am[c] = *reinterpret_cast<const uchar4*>(arg + t * H + (c * LANES + lane) * 4);
    }
    for (int j = 0; j < BS; ++j) {
        Row<LANES, CH> r;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (mode == 0) r.v[c] = d.v[c];
            else if (mode == 1) r.v[c] = make_float4(d.v[c].x / (float)BS, d.v[c].y / (float)BS, d.v[c].z / (float)BS, d.v[c].w / (float)BS);
            else r.v[c] = make_float4(am[c].x == j ? d.v[c].x : 0.f, am[c].y == j ? d.v[c].y : 0.f, am[c].z == j ? d.v[c].z : 0.f,
                                      am[c].w == j ? d.v[c].w : 0.f);
        }
        r.store(d_rows + (t * BS + j) * H, lane);
    }
}

// ---------------------------------------------------------------------------------------------
// fused embedding backward (recomputes the forward from ids + saved LayerNorm statistics)
// partials: [gridDim.x][4][H] (dgamma1, dbeta1, dgamma2, dbeta2)
// ---------------------------------------------------------------------------------------------
template <int LANES, int CH>
__global__ void __launch_bounds__(256) embed_bwd_kernel(const asme_embed_desc d, int T, int S, int H,
                                                        const float* __restrict__ d_out, const float* __restrict__ stats,
                                                        float* __restrict__ d_item_rows, float* __restrict__ d_attr_rows,
                                                        float* __restrict__ partials) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x % LANES;
    const int groups_per_block = blockDim.x / LANES;
    const int group_in_block = threadIdx.x / LANES;
    const float inv_keep = d.p_drop > 0.f ? 1.0f / (1.0f - d.p_drop) : 1.0f;
    Row<LANES, CH> dg1, db1, dg2, db2;
    zero_row(dg1); zero_row(db1); zero_row(dg2); zero_row(db2);
    const int pre = d.n_user > 0 ? 1 : 0;
    for (long long t = (long long)blockIdx.x * groups_per_block + group_in_block; t < T;
         t += (long long)gridDim.x * groups_per_block) {
        Row<LANES, CH> x, xhat1, xhat2, g;
        const long long b = (unsigned)t / (unsigned)S;       // 32-bit: T is an int
        const int sp = (int)(t - b * S);
        const bool is_user = pre && sp == 0;
        const long long src = pre ? b * (S - 1) + sp - 1 : t;
        if (is_user) {
            x.load(d.user_table[0] + __ldg(d.user_ids[0] + b) * H, lane);
            for (int u = 1; u < d.n_user; ++u) x.add(d.user_table[u] + __ldg(d.user_ids[u] + b) * H, lane);
        } else {
            x.load(d.item_table + __ldg(d.item_ids + src) * H, lane);
            if (d.pos_table) x.add(d.pos_table + (long long)(sp - pre) * H, lane);
        }
        float rstd1 = 0.f, rstd2 = 0.f;
        if (d.ln1_gamma && !is_user) {
            const float mean = stats[t];
            rstd1 = stats[(size_t)T + t];
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const float4 gm = ldg4(d.ln1_gamma + (c * LANES + lane) * 4);
                const float4 bt = ldg4(d.ln1_beta + (c * LANES + lane) * 4);
                float4 h;
                h.x = (x.v[c].x - mean) * rstd1; h.y = (x.v[c].y - mean) * rstd1;
                h.z = (x.v[c].z - mean) * rstd1; h.w = (x.v[c].w - mean) * rstd1;
                xhat1.v[c] = h;
                x.v[c].x = h.x * gm.x + bt.x; x.v[c].y = h.y * gm.y + bt.y;
                x.v[c].z = h.z * gm.z + bt.z; x.v[c].w = h.w * gm.w + bt.w;
            }
            if (d.p_drop > 0.f) apply_dropout<LANES, CH>(x, asme_seed(d.seed), d.site_a, t, H, lane, d.p_drop, inv_keep);
        }
        g.load(d_out + t * H, lane);
        if (d.ln2_gamma) {
            if (!is_user) embed_gather_attrs<LANES, CH>(x, d, src, H, lane);
            if (d.seg_table) x.add(d.seg_table + (is_user ? 0 : H), lane);
            const float mean = stats[(size_t)2 * T + t];
            rstd2 = stats[(size_t)3 * T + t];
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                xhat2.v[c].x = (x.v[c].x - mean) * rstd2; xhat2.v[c].y = (x.v[c].y - mean) * rstd2;
                xhat2.v[c].z = (x.v[c].z - mean) * rstd2; xhat2.v[c].w = (x.v[c].w - mean) * rstd2;
            }
            if (d.p_drop > 0.f) apply_dropout<LANES, CH>(g, asme_seed(d.seed), d.site_b, t, H, lane, d.p_drop, inv_keep);
            ln_backward<LANES, CH>(g, xhat2, d.ln2_gamma, lane, H, rstd2, dg2, db2);
        }
        if (d_attr_rows && d_attr_rows != d_item_rows) g.store(d_attr_rows + t * H, lane);
        if (d.ln1_gamma && !is_user) {
            if (d.p_drop > 0.f) apply_dropout<LANES, CH>(g, asme_seed(d.seed), d.site_a, t, H, lane, d.p_drop, inv_keep);
            ln_backward<LANES, CH>(g, xhat1, d.ln1_gamma, lane, H, rstd1, dg1, db1);
        }
        g.store(d_item_rows + t * H, lane);
    }
    float* out = partials + (size_t)blockIdx.x * 4 * H;
    block_reduce_rows<LANES, CH>(dg1, smem, out, H, group_in_block, groups_per_block, lane);
    block_reduce_rows<LANES, CH>(db1, smem, out + H, H, group_in_block, groups_per_block, lane);
    block_reduce_rows<LANES, CH>(dg2, smem, out + 2 * H, H, group_in_block, groups_per_block, lane);
    block_reduce_rows<LANES, CH>(db2, smem, out + 3 * H, H, group_in_block, groups_per_block, lane);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm forward / backward
// ---------------------------------------------------------------------------------------------
template <int LANES, int CH>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int M, int H,
                                                            float* __restrict__ y, float* __restrict__ stats,
                                                            __nv_bfloat16* __restrict__ y16, const int32_t* __restrict__ n_live) {
    const int lane = threadIdx.x % LANES;
    const long long r = (long long)blockIdx.x * (blockDim.x / LANES) + threadIdx.x / LANES;
    if (r >= asme_live_rows(M, n_live)) return;
    Row<LANES, CH> xr, yr;
    xr.load(x + r * H, lane);
    float mean, rstd;
    ln_forward<LANES, CH>(xr, yr, gamma, beta, lane, H, mean, rstd);
    if (stats && lane == 0) { stats[r] = mean; stats[(size_t)M + r] = rstd; }
    if (y) yr.store(y + r * H, lane);
    if (y16) yr.store_bf16(y16 + r * H, lane);
}

template <int LANES, int CH>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ gamma, const float* __restrict__ stats,
                                                            int M, int H, const float* __restrict__ d_residual,
                                                            float* __restrict__ dx, float* __restrict__ partials,
                                                            float p_drop, uint64_t seed, uint32_t site_a, uint32_t site_b,
                                                            __nv_bfloat16* __restrict__ dx16, const int32_t* __restrict__ n_live) {
    extern __shared__ float smem[];
    const float inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    const int M_stats = M;                       // the statistics were saved with the capacity as their row stride
    M = asme_live_rows(M, n_live);
    const int lane = threadIdx.x % LANES;
    const int groups_per_block = blockDim.x / LANES;
    const int group_in_block = threadIdx.x / LANES;
    Row<LANES, CH> dg, db;
    zero_row(dg); zero_row(db);
    for (long long r = (long long)blockIdx.x * groups_per_block + group_in_block; r < M;
         r += (long long)gridDim.x * groups_per_block) {
        Row<LANES, CH> xhat, g;
        xhat.load(x + r * H, lane);
        const float mean = stats[r], rstd = stats[(size_t)M_stats + r];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            xhat.v[c].x = (xhat.v[c].x - mean) * rstd; xhat.v[c].y = (xhat.v[c].y - mean) * rstd;
            xhat.v[c].z = (xhat.v[c].z - mean) * rstd; xhat.v[c].w = (xhat.v[c].w - mean) * rstd;
        }
        g.load(dy + r * H, lane);
        ln_backward<LANES, CH>(g, xhat, gamma, lane, H, rstd, dg, db);
        if (d_residual) g.add(d_residual + r * H, lane);
        // what the next stage of the backward pass consumes, produced here instead of by a separate dropout_cast launch:
        // dx (fp32) = gradient * mask_a, dx16 = bf16(dx * mask_b)   (site 0 = no mask)
        if (p_drop > 0.f && site_a) apply_dropout<LANES, CH>(g, asme_seed(seed), site_a, r, H, lane, p_drop, inv_keep);
        g.store(dx + r * H, lane);
        if (dx16) {
            if (p_drop > 0.f && site_b) apply_dropout<LANES, CH>(g, asme_seed(seed), site_b, r, H, lane, p_drop, inv_keep);
            g.store_bf16(dx16 + r * H, lane);
        }
    }
    float* out = partials + (size_t)blockIdx.x * 2 * H;
    block_reduce_rows<LANES, CH>(dg, smem, out, H, group_in_block, groups_per_block, lane);
    block_reduce_rows<LANES, CH>(db, smem, out + H, H, group_in_block, groups_per_block, lane);
}

// ---------------------------------------------------------------------------------------------
// column sums: out[n] (+)= sum_m x[m,n]. stage 1: one block per chunk of rows; stage 2: over chunks.
// ---------------------------------------------------------------------------------------------
#define COLSUM_ROWS 128
__global__ void colsum_stage1_kernel(const float* __restrict__ x, int M, int N, float* __restrict__ partial,
                                     const int32_t* __restrict__ m_live) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    M = asme_live_rows(M, m_live);               // chunks past the live rows write zeros
    const int m0 = blockIdx.y * COLSUM_ROWS;
    const int m1 = min(M, m0 + COLSUM_ROWS);
    float s = 0.f;
    for (int m = m0; m < m1; ++m) s += x[(size_t)m * N + n];
    partial[(size_t)blockIdx.y * N + n] = s;
}
// out[n] (+)= sum_c partial[c][n].  1024 threads per block: 32 columns x 32 chunk groups, four independent accumulators per
// thread, so a reduction over several hundred partial rows is ~5 dependent load rounds instead of hundreds (it used to take
// 25-30 us per call, ~10 % of a training step); fixed summation order -> deterministic.
__global__ void __launch_bounds__(1024) colsum_stage2_kernel(const float* __restrict__ partial, int chunks, int N,
                                                             float* __restrict__ out, int accumulate) {
    __shared__ float red[32][33];
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
    const int n = blockIdx.x * 32 + tx;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (n < N) {
        int c = ty;
        for (; c + 96 < chunks; c += 128) {
            s0 += partial[(size_t)c * N + n];
            s1 += partial[(size_t)(c + 32) * N + n];
            s2 += partial[(size_t)(c + 64) * N + n];
            s3 += partial[(size_t)(c + 96) * N + n];
        }
        for (; c < chunks; c += 32) s0 += partial[(size_t)c * N + n];
    }
    red[ty][tx] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (ty == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < 32; ++g) s += red[g][tx];
        out[n] = accumulate ? out[n] + s : s;
    }
}

static int launch_colsum(const float* x, int M, int N, float* out, int accumulate, void* ws, size_t ws_bytes,
                         cudaStream_t stream, const int32_t* m_live = nullptr) {
    if (M <= COLSUM_ROWS && m_live == nullptr) {   // small: single stage
        colsum_stage2_kernel<<<ceil_div(N, 32), 1024, 0, stream>>>(x, M, N, out, accumulate);
        ASME_LAUNCH_OK();
        return ASME_OK;
    }
    const int chunks = ceil_div(M, COLSUM_ROWS);
    if (ws_bytes < (size_t)chunks * N * sizeof(float)) {
        asme_set_error("colsum: workspace too small (%zu < %zu)", ws_bytes, (size_t)chunks * N * sizeof(float));
        return ASME_ERR_WORKSPACE;
    }
    float* partial = (float*)ws;
    colsum_stage1_kernel<<<dim3(ceil_div(N, 128), chunks), 128, 0, stream>>>(x, M, N, partial, m_live);
    ASME_LAUNCH_OK();
    colsum_stage2_kernel<<<ceil_div(N, 32), 1024, 0, stream>>>(partial, chunks, N, out, accumulate);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

extern "C" size_t asme_b200_colsum_workspace_bytes(int M, int N) {
    return (size_t)ceil_div(M, COLSUM_ROWS) * N * sizeof(float);
}
extern "C" int asme_b200_colsum_accumulate(const float* x, int M, int N, float* out, void* ws, size_t ws_bytes,
                                           asme_stream_t stream) {
    ASME_REQUIRE(M >= 0 && N > 0, "colsum: bad shape M=%d N=%d", M, N);
    if (M == 0) return ASME_OK;
    return launch_colsum(x, M, N, out, 1, ws, ws_bytes, (cudaStream_t)stream);
}
// the same over the live rows of a capacity-sized buffer (device count, may be NULL)
extern "C" int asme_b200_colsum_accumulate_live(const float* x, int M, int N, float* out, void* ws, size_t ws_bytes,
                                                const int32_t* m_live, asme_stream_t stream) {
    ASME_REQUIRE(M >= 0 && N > 0, "colsum: bad shape M=%d N=%d", M, N);
    if (M == 0) return ASME_OK;
    return launch_colsum(x, M, N, out, 1, ws, ws_bytes, (cudaStream_t)stream, m_live);
}

// d_pos[s,:] += sum_b d_rows[b*S+s,:]
__global__ void posgrad_kernel(const float* __restrict__ d_rows, int B, int S, int H, float* __restrict__ d_pos) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over S*H/4
    if (i >= (long long)S * H / 4) return;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) add4(s, ldg4(d_rows + ((long long)b * S * H) + i * 4));
    float4* o = reinterpret_cast<float4*>(d_pos) + i;
    float4 cur = *o;
    add4(cur, s);
    *o = cur;
}
__global__ void posgrad_strided_kernel(const float* __restrict__ d_rows, int B, int S, long long seq_stride, int H, float* __restrict__ d_pos) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over S*H/4
    if (i >= (long long)S * H / 4) return;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) add4(s, ldg4(d_rows + (long long)b * seq_stride + i * 4));
    float4* o = reinterpret_cast<float4*>(d_pos) + i;
    float4 cur = *o;
    add4(cur, s);
    *o = cur;
}
extern "C" int asme_b200_posgrad_reduce_strided(const float* d_rows, int B, int S, int seq_stride_rows, int H, float* d_pos,
                                                asme_stream_t stream) {
    ASME_REQUIRE(H % 4 == 0, "posgrad: H=%d must be a multiple of 4", H);
    ASME_REQUIRE(seq_stride_rows >= S, "posgrad: sequence stride %d < S=%d", seq_stride_rows, S);
    const long long n = (long long)S * H / 4;
    if (n == 0 || B == 0) return ASME_OK;
    posgrad_strided_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(d_rows, B, S, (long long)seq_stride_rows * H, H, d_pos);
    ASME_LAUNCH_OK();
    return ASME_OK;
}
extern "C" int asme_b200_posgrad_reduce(const float* d_rows, int B, int S, int H, float* d_pos, asme_stream_t stream) {
    ASME_REQUIRE(H % 4 == 0, "posgrad: H=%d must be a multiple of 4", H);
    const long long n = (long long)S * H / 4;
    posgrad_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(d_rows, B, S, H, d_pos);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ---------------------------------------------------------------------------------------------
// dispatch on H
// ---------------------------------------------------------------------------------------------
#define DISPATCH_H(H, CALL)                                       \
    switch (H) {                                                  \
        case 16: { CALL(4, 1); break; }                           \
        case 32: { CALL(8, 1); break; }                           \
        case 64: { CALL(16, 1); break; }                          \
        case 128: { CALL(32, 1); break; }                         \
        case 256: { CALL(32, 2); break; }                         \
        case 512: { CALL(32, 4); break; }                         \
        default:                                                  \
            asme_set_error("unsupported hidden size H=%d (supported: 16,32,64,128,256,512)", H); \
            return ASME_ERR_INVALID;                              \
    }

static int lanes_for(int H) { return H / 4 < 32 ? H / 4 : 32; }

// Forward kernels (embedding gather, LayerNorm): WIDE rows -- a row is owned by H/16 lanes with four 128-bit chunks each instead of
// H/4 lanes with one.  The per-row work that does not scale with the row (address arithmetic, the two shuffle reductions of the
// LayerNorm statistics, rsqrt) is paid by a quarter of the threads: the narrow layout issued ~265 warp instructions per token of
// H = 128 and was ALU-bound at 78 % issue utilisation (ncu, C5 shape), not memory-bound.  Four chunks per lane keep the same four
// independent 128-bit loads in flight per thread that the narrow kernel gets from four tokens per lane group.
static int g_row_wide = 1;
extern "C" int asme_b200_rowwise_tune(int knob, int value) {
    ASME_REQUIRE(knob == 0 && (value == 0 || value == 1), "rowwise_tune: knob 0 (wide forward rows) takes 0 or 1");
    g_row_wide = value;
    return ASME_OK;
}
static bool wide_rows(int H) { return g_row_wide && H >= 64; }
static int lanes_wide(int H) { return H / 16; }
#define DISPATCH_H_WIDE(H, CALL)                                  \
    switch (H) {                                                  \
        case 64: { CALL(4, 4); break; }                           \
        case 128: { CALL(8, 4); break; }                          \
        case 256: { CALL(16, 4); break; }                         \
        case 512: { CALL(32, 4); break; }                         \
        default:                                                  \
            asme_set_error("unsupported hidden size H=%d (supported: 16,32,64,128,256,512)", H); \
            return ASME_ERR_INVALID;                              \
    }

extern "C" int asme_b200_embed_pool_fwd(const int64_t* ids, const float* table, int T, int BS, int H, int mode, float* out,
                                        uint8_t* arg, asme_stream_t stream) {
    ASME_REQUIRE(ids && table && out, "embed_pool_fwd: null argument");
    ASME_REQUIRE(mode >= 0 && mode <= 2 && BS >= 1 && BS <= 255, "embed_pool_fwd: mode=%d BS=%d unsupported", mode, BS);
    ASME_REQUIRE(mode != 2 || arg, "embed_pool_fwd: max pooling needs the arg output");
    if (T == 0) return ASME_OK;
    const int lanes = lanes_for(H);
    const int groups = 256 / lanes;
#define CALL(L, C) embed_pool_fwd_kernel<L, C><<<ceil_div(T, groups), 256, 0, (cudaStream_t)stream>>>(ids, table, T, BS, H, mode, out, arg)
    DISPATCH_H(H, CALL)
#undef CALL
    ASME_LAUNCH_OK();
    return ASME_OK;
}
extern "C" int asme_b200_embed_pool_bwd(const float* d_out, const uint8_t* arg, int T, int BS, int H, int mode, float* d_rows,
                                        asme_stream_t stream) {
    ASME_REQUIRE(d_out && d_rows, "embed_pool_bwd: null argument");
    ASME_REQUIRE(mode >= 0 && mode <= 2 && BS >= 1 && BS <= 255, "embed_pool_bwd: mode=%d BS=%d unsupported", mode, BS);
    ASME_REQUIRE(mode != 2 || arg, "embed_pool_bwd: max pooling needs arg");
    if (T == 0) return ASME_OK;
    const int lanes = lanes_for(H);
    const int groups = 256 / lanes;
#define CALL(L, C) embed_pool_bwd_kernel<L, C><<<ceil_div(T, groups), 256, 0, (cudaStream_t)stream>>>(d_out, arg, T, BS, H, mode, d_rows)
    DISPATCH_H(H, CALL)
#undef CALL
    ASME_LAUNCH_OK();
    return ASME_OK;
}

extern "C" int asme_b200_embed_fwd(const asme_embed_desc* d, int T, int S, int H, float* out, float* stats,
                                   asme_stream_t stream) {
    ASME_REQUIRE(d && out, "embed_fwd: null argument");
    ASME_REQUIRE(T >= 0 && S > 0, "embed_fwd: bad shape T=%d S=%d", T, S);
    ASME_REQUIRE(d->n_attr >= 0 && d->n_attr <= ASME_MAX_ATTR && d->n_bag >= 0 && d->n_bag <= ASME_MAX_ATTR,
                 "embed_fwd: too many attribute tables");
    ASME_REQUIRE(d->p_drop >= 0.f && d->p_drop < 1.f, "embed_fwd: dropout p=%f out of range", d->p_drop);
    ASME_REQUIRE(d->n_user >= 0 && d->n_user <= ASME_MAX_ATTR, "embed_fwd: too many user-attribute tables");
    ASME_REQUIRE(d->n_user == 0 || (S >= 2 && T % S == 0), "embed_fwd: user prefix needs S >= 2 (S counts the user position) and T = B*S");
    if (T == 0) return ASME_OK;
    const bool wide = wide_rows(H);
    const int lanes = wide ? lanes_wide(H) : lanes_for(H);
    const int groups = 256 / lanes;
    // tokens per lane group: as many as still leave >= 4 blocks per SM (wide rows: one -- the four chunks are the loads in flight)
    const int tok = wide ? 1 : (ceil_div(T, groups * 4) >= ASME_NUM_SMS * 4 ? 4 : (ceil_div(T, groups * 2) >= ASME_NUM_SMS * 4 ? 2 : 1));
#define LAUNCH_TOK(L, C, K, P) embed_fwd_kernel<L, C, K, P><<<ceil_div(T, groups * K), 256, 0, (cudaStream_t)stream>>>(*d, T, S, H, out, stats)
#define CALL(L, C)                                                                       \
    {                                                                                    \
        if (d->n_user > 0) {                                                             \
            if (tok == 4) LAUNCH_TOK(L, C, 4, true);                                     \
            else if (tok == 2) LAUNCH_TOK(L, C, 2, true);                                \
            else LAUNCH_TOK(L, C, 1, true);                                              \
        } else {                                                                         \
            if (tok == 4) LAUNCH_TOK(L, C, 4, false);                                    \
            else if (tok == 2) LAUNCH_TOK(L, C, 2, false);                               \
            else LAUNCH_TOK(L, C, 1, false);                                             \
        }                                                                                \
    }
    if (wide) {
#define CALLW(L, C)                                                                      \
    {                                                                                    \
        if (d->n_user > 0) LAUNCH_TOK(L, C, 1, true);                                    \
        else LAUNCH_TOK(L, C, 1, false);                                                 \
    }
        DISPATCH_H_WIDE(H, CALLW)
#undef CALLW
    } else {
        DISPATCH_H(H, CALL)
    }
#undef CALL
#undef LAUNCH_TOK
    ASME_LAUNCH_OK();
    return ASME_OK;
}

static int rowwise_bwd_grid(int M, int groups) {
    const int need = ceil_div(M, groups);
    const int cap = ASME_NUM_SMS * 4;
    return need < cap ? need : cap;
}

extern "C" size_t asme_b200_embed_bwd_workspace_bytes(int T, int H) {
    (void)T;
    return (size_t)ASME_NUM_SMS * 4 * 4 * H * sizeof(float);
}

extern "C" int asme_b200_embed_bwd(const asme_embed_desc* d, int T, int S, int H, const float* d_out, const float* stats,
                                   float* d_item_rows, float* d_attr_rows, float* dln, void* ws, size_t ws_bytes,
                                   asme_stream_t stream) {
    ASME_REQUIRE(d && d_out && d_item_rows, "embed_bwd: null argument");
    ASME_REQUIRE(!(d->ln1_gamma || d->ln2_gamma) || (stats && dln), "embed_bwd: LayerNorm needs stats and dln");
    if (T == 0) return ASME_OK;
    const int lanes = lanes_for(H);
    const int groups = 256 / lanes;
    const int grid = rowwise_bwd_grid(T, groups);
    if (ws_bytes < (size_t)grid * 4 * H * sizeof(float)) {
        asme_set_error("embed_bwd: workspace too small");
        return ASME_ERR_WORKSPACE;
    }
    float* partials = (float*)ws;
    const size_t smem = (size_t)groups * H * sizeof(float);
#define CALL(L, C)                                                                                              \
    embed_bwd_kernel<L, C><<<grid, 256, smem, (cudaStream_t)stream>>>(*d, T, S, H, d_out, stats, d_item_rows, \
                                                                      d_attr_rows, partials)
    DISPATCH_H(H, CALL)
#undef CALL
    ASME_LAUNCH_OK();
    if (dln) {
        colsum_stage2_kernel<<<ceil_div(4 * H, 32), 1024, 0, (cudaStream_t)stream>>>(partials, grid, 4 * H, dln, 1);
        ASME_LAUNCH_OK();
    }
    return ASME_OK;
}

extern "C" int asme_b200_layernorm_fwd(const float* x, const float* gamma, const float* beta, int M, int H, float* y,
                                       float* stats, const int32_t* n_live, asme_stream_t stream) {
    ASME_REQUIRE(x && gamma && beta && y, "layernorm_fwd: null argument");
    if (M == 0) return ASME_OK;
    const bool wide = wide_rows(H);
    const int lanes = wide ? lanes_wide(H) : lanes_for(H);
    const int groups = 256 / lanes;
#define CALL(L, C) layernorm_fwd_kernel<L, C><<<ceil_div(M, groups), 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, M, H, y, stats, nullptr, n_live)
    if (wide) { DISPATCH_H_WIDE(H, CALL) } else { DISPATCH_H(H, CALL) }
#undef CALL
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// LayerNorm whose output feeds a tensor-core GEMM: y as bf16 (and optionally fp32 too)
extern "C" int asme_b200_layernorm_fwd_bf16(const float* x, const float* gamma, const float* beta, int M, int H, float* y_f32,
                                            void* y_bf16, float* stats, asme_stream_t stream) {
    ASME_REQUIRE(x && gamma && beta && y_bf16, "layernorm_fwd_bf16: null argument");
    if (M == 0) return ASME_OK;
    const bool wide = wide_rows(H);
    const int lanes = wide ? lanes_wide(H) : lanes_for(H);
    const int groups = 256 / lanes;
#define CALL(L, C)                                                                                                     \
    layernorm_fwd_kernel<L, C><<<ceil_div(M, groups), 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, M, H, y_f32, stats, \
                                                                                      (__nv_bfloat16*)y_bf16, nullptr)
    if (wide) { DISPATCH_H_WIDE(H, CALL) } else { DISPATCH_H(H, CALL) }
#undef CALL
    ASME_LAUNCH_OK();
    return ASME_OK;
}

extern "C" size_t asme_b200_layernorm_bwd_workspace_bytes(int M, int H) {
    (void)M;
    return (size_t)ASME_NUM_SMS * 4 * 2 * H * sizeof(float);
}

static int layernorm_bwd_impl(const float* dy, const float* x, const float* gamma, const float* stats, int M, int H,
                              const float* d_residual, float* dx, float* dgb, void* ws, size_t ws_bytes, float p_drop, uint64_t seed,
                              uint32_t site_a, uint32_t site_b, void* dx_bf16, const int32_t* n_live, asme_stream_t stream);
extern "C" int asme_b200_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* stats, int M,
                                       int H, const float* d_residual, float* dx, float* dgb, void* ws, size_t ws_bytes,
                                       const int32_t* n_live, asme_stream_t stream) {
    return layernorm_bwd_impl(dy, x, gamma, stats, M, H, d_residual, dx, dgb, ws, ws_bytes, 0.f, 0ull, 0u, 0u, nullptr, n_live, stream);
}
// LayerNorm backward that also emits what the next backward stage consumes (replaces a dropout_cast launch):
// dx (fp32) = (LN gradient + d_residual) * mask(site_a),  dx_bf16 = bf16(dx * mask(site_b));  site 0 = no mask
extern "C" int asme_b200_layernorm_bwd_drop(const float* dy, const float* x, const float* gamma, const float* stats, int M,
                                            int H, const float* d_residual, float* dx, float* dgb, void* ws, size_t ws_bytes,
                                            float p_drop, uint64_t seed, uint32_t site_a, uint32_t site_b, void* dx_bf16,
                                            asme_stream_t stream) {
    ASME_REQUIRE(dx_bf16, "layernorm_bwd_drop: null bf16 output");
    ASME_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "layernorm_bwd_drop: p=%f out of range", p_drop);
    return layernorm_bwd_impl(dy, x, gamma, stats, M, H, d_residual, dx, dgb, ws, ws_bytes, p_drop, seed, site_a, site_b, dx_bf16, nullptr, stream);
}
static int layernorm_bwd_impl(const float* dy, const float* x, const float* gamma, const float* stats, int M, int H,
                              const float* d_residual, float* dx, float* dgb, void* ws, size_t ws_bytes, float p_drop, uint64_t seed,
                              uint32_t site_a, uint32_t site_b, void* dx_bf16, const int32_t* n_live, asme_stream_t stream) {
    // dgb == NULL: the (gamma, beta) partials stay in ``ws`` -- [asme_b200_layernorm_bwd_chunks(M,H)][2H] -- and the caller reduces them
    // later with asme_b200_rows_reduce (the reduction is a leaf of the backward pass: second stream / parallel graph branch)
    ASME_REQUIRE(dy && x && gamma && stats && dx, "layernorm_bwd: null argument");
    if (M == 0) return ASME_OK;
    const int lanes = lanes_for(H);
    const int groups = 256 / lanes;
    const int grid = rowwise_bwd_grid(M, groups);
    if (ws_bytes < (size_t)grid * 2 * H * sizeof(float)) {
        asme_set_error("layernorm_bwd: workspace too small");
        return ASME_ERR_WORKSPACE;
    }
    float* partials = (float*)ws;
    const size_t smem = (size_t)groups * H * sizeof(float);
#define CALL(L, C)                                                                                                 \
    layernorm_bwd_kernel<L, C><<<grid, 256, smem, (cudaStream_t)stream>>>(dy, x, gamma, stats, M, H, d_residual, dx, \
                                                                          partials, p_drop, seed, site_a, site_b,   \
                                                                          (__nv_bfloat16*)dx_bf16, n_live)
    DISPATCH_H(H, CALL)
#undef CALL
    ASME_LAUNCH_OK();
    if (dgb != nullptr) {
        colsum_stage2_kernel<<<ceil_div(2 * H, 32), 1024, 0, (cudaStream_t)stream>>>(partials, grid, 2 * H, dgb, 1);
        ASME_LAUNCH_OK();
    }
    return ASME_OK;
}
extern "C" int asme_b200_layernorm_bwd_chunks(int M, int H) {
    if (M <= 0 || H < 16) return 0;
    return rowwise_bwd_grid(M, 256 / lanes_for(H));
}
// out[n] (+)= sum over the chunks of partial[chunk][n]: the second stage of the column reductions, as its own launch
extern "C" int asme_b200_rows_reduce(const float* partial, int chunks, int N, float* out, int accumulate, asme_stream_t stream) {
    ASME_REQUIRE(partial && out && chunks >= 0 && N >= 1, "rows_reduce: bad argument");
    if (chunks == 0) return ASME_OK;
    colsum_stage2_kernel<<<ceil_div(N, 32), 1024, 0, (cudaStream_t)stream>>>(partial, chunks, N, out, accumulate);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ---------------------------------------------------------------------------------------------
// elementwise helpers
// ---------------------------------------------------------------------------------------------
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, long long n4, float p, float inv_keep,
                               uint64_t seed, uint32_t site) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    seed = asme_seed(seed);
    float4 v = ldg4(x + i * 4);
    const float4 s = dropout_scale4(seed, site, (uint64_t)i * 4, p, inv_keep);
    v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
    reinterpret_cast<float4*>(y)[i] = v;
}
extern "C" int asme_b200_dropout(const float* x, float* y, long long n, float p, uint64_t seed, uint32_t site,
                                 asme_stream_t stream) {
    ASME_REQUIRE(n % 4 == 0, "dropout: n=%lld must be a multiple of 4", n);
    ASME_REQUIRE(p >= 0.f && p < 1.f, "dropout: p=%f out of range", p);
    if (n == 0) return ASME_OK;
    dropout_kernel<<<ceil_div(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n / 4, p, 1.0f / (1.0f - p), seed, site);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// y_f32 = x * mask_a ; y_bf16 = bf16(y_f32 * mask_b): the gradient entering a residual block is needed both as the fp32
// residual gradient (after the block-end dropout, site_a) and as the bf16 GEMM operand (after the sub-layer dropout, site_b)
__global__ void dropout_cast_kernel(const float* __restrict__ x, long long n4, float p, float inv_keep, uint64_t seed,
                                    uint32_t site_a, uint32_t site_b, float* __restrict__ y32, __nv_bfloat16* __restrict__ y16) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    seed = asme_seed(seed);
    float4 v = ldg4(x + i * 4);
    if (p > 0.f && site_a) {
        const float4 s = dropout_scale4(seed, site_a, (uint64_t)i * 4, p, inv_keep);
        v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
    }
    if (y32) reinterpret_cast<float4*>(y32)[i] = v;
    if (p > 0.f && site_b) {
        const float4 s = dropout_scale4(seed, site_b, (uint64_t)i * 4, p, inv_keep);
        v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
    }
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&lo);
    o.y = *reinterpret_cast<const uint32_t*>(&hi);
    reinterpret_cast<uint2*>(y16)[i] = o;
}
extern "C" int asme_b200_dropout_cast(const float* x, long long n, float p, uint64_t seed, uint32_t site_a, uint32_t site_b,
                                      float* y_f32, void* y_bf16, asme_stream_t stream) {
    ASME_REQUIRE(x && y_bf16, "dropout_cast: null argument");
    ASME_REQUIRE(n % 4 == 0, "dropout_cast: n=%lld must be a multiple of 4", n);
    ASME_REQUIRE(p >= 0.f && p < 1.f, "dropout_cast: p=%f out of range", p);
    if (n == 0) return ASME_OK;
    dropout_cast_kernel<<<ceil_div(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, n / 4, p, 1.0f / (1.0f - p), seed, site_a, site_b,
                                                                               y_f32, (__nv_bfloat16*)y_bf16);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

__global__ void binary_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, long long n,
                              int op) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    y[i] = op == 0 ? a[i] + b[i] : a[i] * b[i];
}
extern "C" int asme_b200_binary(const float* a, const float* b, float* y, long long n, int op, asme_stream_t stream) {
    if (n == 0) return ASME_OK;
    binary_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, y, n, op);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

__global__ void gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z, float* __restrict__ dz, long long n,
                                int row_width, const int32_t* __restrict__ n_live) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n_live != nullptr) n = min(n, (long long)__ldg(n_live) * row_width);
    if (i < n) dz[i] = dy[i] * gelu_erf_grad(z[i]);
}
extern "C" int asme_b200_gelu_bwd(const float* dy, const float* z, float* dz, long long n, int row_width, const int32_t* n_live,
                                  asme_stream_t stream) {
    ASME_REQUIRE(dy && z && dz, "gelu_bwd: null argument");
    ASME_REQUIRE(!n_live || row_width > 0, "gelu_bwd: a live-row count needs the row width");
    if (n == 0) return ASME_OK;
    gelu_bwd_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, z, dz, n, row_width, n_live);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ---------------------------------------------------------------------------------------------
// row selection on the device: rows[] = ascending flat indices t with target[t] != ignore_id (the positions the cross entropy
// sees, nn.CrossEntropyLoss(ignore_index=pad): modules/masked_training_module.py:93-111), row_targets[] = their targets,
// n_rows[0] = how many.  Slots past n_rows carry -1 / ignore_id.  Two launches: per-block counts, then every block sums the
// counts of the blocks before it (a few hundred integers) and writes its own matches in order -- deterministic, no atomics,
// no host round trip (torch.nonzero synchronises to size its result).
// ---------------------------------------------------------------------------------------------
#define SEL_THREADS 256
#define SEL_PER_THREAD 8
#define SEL_BLOCK (SEL_THREADS * SEL_PER_THREAD)
__global__ void __launch_bounds__(SEL_THREADS) select_count_kernel(const int64_t* __restrict__ target, long long T, int64_t ignore_id,
                                                                   int32_t* __restrict__ block_counts) {
    __shared__ int warp_counts[SEL_THREADS / 32];
    const long long base = (long long)blockIdx.x * SEL_BLOCK + (long long)threadIdx.x * SEL_PER_THREAD;
    int c = 0;
#pragma unroll
    for (int e = 0; e < SEL_PER_THREAD; ++e)
        if (base + e < T && target[base + e] != ignore_id) ++c;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (threadIdx.x % 32 == 0) warp_counts[threadIdx.x / 32] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < SEL_THREADS / 32; ++w) s += warp_counts[w];
        block_counts[blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(SEL_THREADS) select_write_kernel(const int64_t* __restrict__ target, long long T, int64_t ignore_id,
                                                                   const int32_t* __restrict__ block_counts, int64_t* __restrict__ rows,
                                                                   int64_t* __restrict__ row_targets, int32_t* __restrict__ n_rows) {
    __shared__ int warp_sums[SEL_THREADS / 32];
    __shared__ int block_offset, total;
    // offset of this block = matches of all blocks before it; block 0 also publishes the total
    int before = 0, all = 0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += SEL_THREADS) {
        const int c = block_counts[b];
        all += c;
        if (b < (int)blockIdx.x) before += c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { before += __shfl_xor_sync(0xffffffffu, before, o); all += __shfl_xor_sync(0xffffffffu, all, o); }
    if (threadIdx.x % 32 == 0) { warp_sums[threadIdx.x / 32] = before; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < SEL_THREADS / 32; ++w) s += warp_sums[w];
        block_offset = s;
    }
    __syncthreads();
    if (threadIdx.x % 32 == 0) warp_sums[threadIdx.x / 32] = all;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < SEL_THREADS / 32; ++w) s += warp_sums[w];
        total = s;
        if (blockIdx.x == 0) n_rows[0] = s;
    }
    __syncthreads();
    // matches of this thread's SEL_PER_THREAD consecutive positions, exclusive scan over the block
    const long long base = (long long)blockIdx.x * SEL_BLOCK + (long long)threadIdx.x * SEL_PER_THREAD;
    int64_t tg[SEL_PER_THREAD];
    int c = 0;
#pragma unroll
    for (int e = 0; e < SEL_PER_THREAD; ++e) {
        tg[e] = base + e < T ? target[base + e] : ignore_id;
        if (tg[e] != ignore_id) ++c;
    }
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)(threadIdx.x % 32) >= o) incl += v;
    }
    __syncthreads();
    if (threadIdx.x % 32 == 31) warp_sums[threadIdx.x / 32] = incl;
    __syncthreads();
    int warp_before = 0;
    for (int w = 0; w < (int)(threadIdx.x / 32); ++w) warp_before += warp_sums[w];
    int slot = block_offset + warp_before + incl - c;
#pragma unroll
    for (int e = 0; e < SEL_PER_THREAD; ++e)
        if (tg[e] != ignore_id) { rows[slot] = base + e; row_targets[slot] = tg[e]; ++slot; }
    // the unused tail of the capacity: -1 / ignore_id (every block clears its share)
    for (long long i = (long long)total + (long long)blockIdx.x * SEL_THREADS + threadIdx.x; i < T; i += (long long)gridDim.x * SEL_THREADS) {
        rows[i] = -1;
        row_targets[i] = ignore_id;
    }
}
extern "C" size_t asme_b200_select_rows_workspace_bytes(long long T) { return (size_t)ceil_div(T < 1 ? 1 : T, SEL_BLOCK) * sizeof(int32_t); }
extern "C" int asme_b200_select_rows(const int64_t* target, long long T, int64_t ignore_id, int64_t* rows, int64_t* row_targets,
                                     int32_t* n_rows, void* ws, size_t ws_bytes, asme_stream_t stream) {
    ASME_REQUIRE(target && rows && row_targets && n_rows, "select_rows: null argument");
    ASME_REQUIRE(T >= 0 && T < (1ll << 31), "select_rows: T=%lld out of range", T);
    cudaStream_t st = (cudaStream_t)stream;
    if (T == 0) {
        ASME_CUDA_OK(cudaMemsetAsync(n_rows, 0, sizeof(int32_t), st));
        return ASME_OK;
    }
    const int blocks = ceil_div(T, SEL_BLOCK);
    if (ws_bytes < (size_t)blocks * sizeof(int32_t)) {
        asme_set_error("select_rows: workspace too small");
        return ASME_ERR_WORKSPACE;
    }
    select_count_kernel<<<blocks, SEL_THREADS, 0, st>>>(target, T, ignore_id, (int32_t*)ws);
    ASME_LAUNCH_OK();
    select_write_kernel<<<blocks, SEL_THREADS, 0, st>>>(target, T, ignore_id, (const int32_t*)ws, rows, row_targets, n_rows);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ---------------------------------------------------------------------------------------------
// gradient clipping by global L2 norm over the flat gradient arena (pl.Trainer(gradient_clip_val=c) = torch clip_grad_norm_:
// g *= min(1, c / (||g|| + 1e-6))).  Fixed-order two-stage sum of squares (double), then one scaling pass.
// ---------------------------------------------------------------------------------------------
#define CLIP_BLOCKS (ASME_NUM_SMS * 2)
__global__ void __launch_bounds__(256) sumsq_stage1_kernel(const float* __restrict__ g, long long n, double* __restrict__ partial) {
    __shared__ double red[256];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const double v = (double)g[i];
        acc += v * v;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__global__ void __launch_bounds__(256) clip_scale_kernel(float* __restrict__ g, long long n, const double* __restrict__ partial, int parts,
                                                         float max_norm, float* __restrict__ norm_out) {
    __shared__ float coef_s;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int p = 0; p < parts; ++p) s += partial[p];
        const float norm = (float)sqrt(s);
        const float coef = max_norm / (norm + 1e-6f);
        coef_s = coef < 1.0f ? coef : 1.0f;
        if (norm_out && blockIdx.x == 0) norm_out[0] = norm;
    }
    __syncthreads();
    const float coef = coef_s;
    if (coef >= 1.0f) return;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) g[i] *= coef;
}
extern "C" size_t asme_b200_clip_grad_norm_workspace_bytes(void) { return (size_t)CLIP_BLOCKS * sizeof(double); }
extern "C" int asme_b200_clip_grad_norm(float* grad, long long n, float max_norm, float* norm_out, void* ws, size_t ws_bytes,
                                        asme_stream_t stream) {
    ASME_REQUIRE(grad && max_norm > 0.f, "clip_grad_norm: null gradient or max_norm <= 0");
    ASME_REQUIRE(ws && ws_bytes >= asme_b200_clip_grad_norm_workspace_bytes() && ((uintptr_t)ws & 7) == 0, "clip_grad_norm: workspace too small or misaligned");
    if (n == 0) return ASME_OK;
    cudaStream_t st = (cudaStream_t)stream;
    sumsq_stage1_kernel<<<CLIP_BLOCKS, 256, 0, st>>>(grad, n, (double*)ws);
    ASME_LAUNCH_OK();
    clip_scale_kernel<<<CLIP_BLOCKS, 256, 0, st>>>(grad, n, (const double*)ws, CLIP_BLOCKS, max_norm, norm_out);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

__global__ void fill_kernel(float* x, long long n, float v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}
__global__ void fill4_kernel(float4* x, long long n4, float v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) x[i] = make_float4(v, v, v, v);
}
// ---- device-resident step state (CUDA-graph replays): {seed, adam step, learning rate} ----------------------------------------
struct asme_step_state_t { unsigned long long seed; double lr; long long adam_step; };
__global__ void step_state_advance_kernel(asme_step_state_t* s) { s->seed += 1ull; s->adam_step += 1; }
extern "C" int asme_b200_step_state_advance(void* state, asme_stream_t stream) {
    ASME_REQUIRE(state, "step_state_advance: null argument");
    step_state_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((asme_step_state_t*)state);
    ASME_LAUNCH_OK();
    return ASME_OK;
}
__device__ __forceinline__ void adam_update(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                            float* __restrict__ v, long long i, float lr_over_bc1, float beta1, float beta2, float omb1,
                                            float omb2, float eps, float wd, float inv_sqrt_bc2);
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                long long n, const asme_step_state_t* __restrict__ st, float b1, float b2, float omb1, float omb2,
                                float eps, float wd, double b1d, double b2d) {
    // the two bias-correction scalars are double-precision pow / sqrt: one thread per block computes them (they used to be
    // evaluated by every thread, which made this the slowest "memory-bound" kernel of the step)
    __shared__ float sc[2];
    if (threadIdx.x == 0) {
        const double step = (double)st->adam_step;
        sc[0] = (float)(1.0 / sqrt(1.0 - pow(b2d, step)));       // same double-precision scalars the host computes
        sc[1] = (float)(st->lr / (1.0 - pow(b1d, step)));        // for asme_b200_adam_step
    }
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    adam_update(p, g, m, v, i, sc[1], b1, b2, omb1, omb2, eps, wd, sc[0]);
}
extern "C" int asme_b200_adam_step_dev(float* param, const float* grad, float* m, float* v, long long n, const void* state,
                                       double beta1, double beta2, double eps, double weight_decay, asme_stream_t stream) {
    ASME_REQUIRE(param && grad && m && v && state, "adam_step_dev: null argument");
    if (n == 0) return ASME_OK;
    adam_dev_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, m, v, n, (const asme_step_state_t*)state,
                                                                           (float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2),
                                                                           (float)eps, (float)weight_decay, beta1, beta2);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

extern "C" int asme_b200_fill(float* x, long long n, float value, asme_stream_t stream) {
    if (n == 0) return ASME_OK;
    if (((uintptr_t)x & 15) == 0 && n % 4 == 0) {
        fill4_kernel<<<ceil_div(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(x), n / 4, value);
        ASME_LAUNCH_OK();
        return ASME_OK;
    }
    fill_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, value);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

__global__ void gather_rows_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, int R, int H4,
                                   float* __restrict__ out, int scatter, const int32_t* __restrict__ n_live) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)asme_live_rows(R, n_live) * H4) return;
    const long long r = i / H4, c = i % H4;
    const long long src = idx[r];
    if (src < 0) return;                          // slots past the live rows of a row selection carry -1
    if (!scatter) reinterpret_cast<float4*>(out)[r * H4 + c] = __ldg(reinterpret_cast<const float4*>(x) + src * H4 + c);
    else reinterpret_cast<float4*>(out)[src * H4 + c] = __ldg(reinterpret_cast<const float4*>(x) + r * H4 + c);
}
extern "C" int asme_b200_gather_rows(const float* x, const int64_t* row_index, int R, int H, float* out,
                                     const int32_t* n_live, asme_stream_t stream) {
    ASME_REQUIRE(H % 4 == 0, "gather_rows: H=%d must be a multiple of 4", H);
    if (R == 0) return ASME_OK;
    gather_rows_kernel<<<ceil_div((long long)R * H / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, row_index, R, H / 4, out, 0, n_live);
    ASME_LAUNCH_OK();
    return ASME_OK;
}
extern "C" int asme_b200_scatter_rows(const float* rows, const int64_t* row_index, int R, int H, float* out,
                                      const int32_t* n_live, asme_stream_t stream) {
    ASME_REQUIRE(H % 4 == 0, "scatter_rows: H=%d must be a multiple of 4", H);
    if (R == 0) return ASME_OK;
    gather_rows_kernel<<<ceil_div((long long)R * H / 4, 256), 256, 0, (cudaStream_t)stream>>>(rows, row_index, R, H / 4, out, 1, n_live);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ---------------------------------------------------------------------------------------------
// fused Adam over the flat arena
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void adam_update(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                            float* __restrict__ v, long long i, float lr_over_bc1, float beta1, float beta2, float omb1,
                                            float omb2, float eps, float wd, float inv_sqrt_bc2) {
    float gi = g[i];
    const float pi = p[i];
    if (wd != 0.f) gi += wd * pi;
    const float mi = beta1 * m[i] + omb1 * gi;          // omb = 1 - beta evaluated in double on the host, as torch does
    const float vi = beta2 * v[i] + omb2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - lr_over_bc1 * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
}
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr_over_bc1, float beta1, float beta2, float omb1, float omb2, float eps,
                            float wd, float inv_sqrt_bc2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    adam_update(p, g, m, v, i, lr_over_bc1, beta1, beta2, omb1, omb2, eps, wd, inv_sqrt_bc2);
}
extern "C" int asme_b200_adam_step(float* param, const float* grad, float* m, float* v, long long n, double lr, double beta1,
                                   double beta2, double eps, double weight_decay, int step, asme_stream_t stream) {
    ASME_REQUIRE(step >= 1, "adam: step must be >= 1");
    if (n == 0) return ASME_OK;
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    adam_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
        param, grad, m, v, n, (float)(lr / bc1), (float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2),
        (float)eps, (float)weight_decay, (float)(1.0 / sqrt(bc2)));
    ASME_LAUNCH_OK();
    return ASME_OK;
}
