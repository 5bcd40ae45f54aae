// Exact top-k on the tensor-core path (metrics/common.py:18-27: argsort of the fp32 logits, ties -> lowest item id).
//
// The bf16 sweep (score_tc.cu) is a CANDIDATE GENERATOR: it returns the KC best items by the bf16-operand score s~.  Here every
// candidate is re-scored from the fp32 hidden row and the fp32 table with the arithmetic of the strict fp32 path -- one
// sequential fmaf chain over the hidden dimension, then + bias (score_simt.cu: tile_dot / score_targets) -- so the scores, and the
// (score desc, id asc) order built from them, are bit-identical to what the fp32 sweep produces.
//
// Completeness is CHECKED per row, not assumed: an item outside the candidate list has s~ <= s~_min (the list's last entry), and its
// exact score differs from s~ by at most
//       E = ||h - h~|| max_v ||w_v||  +  ||h~|| max_v ||w_v - w~_v||  +  2^-15 max_v |b_v|  +  2^-16 ||h|| max_v ||w_v||
// (h~, w~: the bf16 operands; Cauchy-Schwarz on (h - h~).w and h~.(w - w~); the hi/lo-split bias; fp32 accumulation).  If s~_min + E < x_k, the exact k-th best candidate score, no outside item can enter the top k: the list is
// certified.  Otherwise the row is FLAGGED and the caller re-runs the exact fp32 sweep for the flagged rows only
// (asme_b200_score_topk_rank with row_flag) -- rare: the list carries KC - k spare entries.
#include "common.cuh"

#include <cuda_bf16.h>
#include <limits.h>

// ---- constants of the certificate (cached by the caller until the weights change): out3 = {max_v ||w_v||_2, max_v |b_v|,
//      max_v ||w_v - bf16(w_v)||_2} ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) table_norm_bound_kernel(const float* __restrict__ W, int V, int H, const float* __restrict__ bias,
                                                               unsigned int* __restrict__ out_bits) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
    const int n_warps = gridDim.x * blockDim.x / 32;
    float best = 0.f, bbest = 0.f, ebest = 0.f;
    for (int v = warp; v < V; v += n_warps) {
        const float* w = W + (size_t)v * H;
        float s = 0.f, e = 0.f;
        for (int k = lane; k < H; k += 32) {
            const float x = __ldg(w + k);
            const float d = x - __bfloat162float(__float2bfloat16_rn(x));
            s = fmaf(x, x, s);
            e = fmaf(d, d, e);
        }
        s = warp_sum(s);
        e = warp_sum(e);
        best = fmaxf(best, s);
        ebest = fmaxf(ebest, e);
        if (bias && lane == 0) bbest = fmaxf(bbest, fabsf(__ldg(bias + v)));
    }
    bbest = warp_max(bbest);
    if (lane == 0) {      // non-negative floats order like their bit patterns: an integer atomicMax is an exact, order-independent max
        atomicMax(out_bits, __float_as_uint(sqrtf(best) * (1.0f + 1e-6f)));
        atomicMax(out_bits + 1, __float_as_uint(bbest));
        atomicMax(out_bits + 2, __float_as_uint(sqrtf(ebest) * (1.0f + 1e-6f)));
    }
}
extern "C" int asme_b200_table_norm_bound(const float* W, int V, int H, const float* bias, float* out3, asme_stream_t stream) {
    ASME_REQUIRE(W && out3 && V >= 1 && H >= 1, "table_norm_bound: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    ASME_CUDA_OK(cudaMemsetAsync(out3, 0, 3 * sizeof(float), st));
    const int blocks = V / 8 + 1 < ASME_NUM_SMS * 8 ? V / 8 + 1 : ASME_NUM_SMS * 8;
    table_norm_bound_kernel<<<blocks, 256, 0, st>>>(W, V, H, bias, reinterpret_cast<unsigned int*>(out3));
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ---- re-score + certify: one warp per row, lane j owns candidates j and j + 32 -------------------------------------------------------
__device__ __forceinline__ bool better_(float v, int id, float v2, int id2) { return v > v2 || (v == v2 && id < id2); }

// the fp32 path's score of item ``id`` for hidden row ``h``: k ascending, one fmaf each, bias last (score_simt.cu: tile_dot)
__device__ __forceinline__ float exact_score(const float* __restrict__ h, const float* __restrict__ w, int H, const float* __restrict__ bias,
                                             long long local_id) {
    float acc = 0.f;
    for (int c = 0; c < H; c += 4) {
        const float4 a = *reinterpret_cast<const float4*>(h + c);
        const float4 b = __ldg(reinterpret_cast<const float4*>(w + c));
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
    if (bias) acc = acc + __ldg(bias + local_id);
    return acc;
}

__global__ void __launch_bounds__(128) topk_rescore_kernel(const float* __restrict__ Hrows, int R, int H, const float* __restrict__ W,
                                                           const float* __restrict__ bias, int v0, int V, const int32_t* __restrict__ cand_idx,
                                                           const float* __restrict__ cand_val, const float* __restrict__ cand_bound,
                                                           int KC, int k,
                                                           const float* __restrict__ norm_bound, const int64_t* __restrict__ target,
                                                           float* __restrict__ topk_val, int32_t* __restrict__ topk_idx,
                                                           float* __restrict__ target_score, int32_t* __restrict__ rank,
                                                           int32_t* __restrict__ row_flag, int32_t* __restrict__ n_flagged) {
    const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= R) return;
    const float* h = Hrows + (size_t)row * H;
    // this lane's candidates (slot 0: lane, slot 1: lane + 32); ids are global catalog ids, W / bias are the slice [v0, v0 + V)
    int id[2];
    float score[2], approx[2];
    bool valid[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int j = lane + 32 * s;
        id[s] = j < KC ? cand_idx[(size_t)row * KC + j] : -1;
        valid[s] = id[s] >= v0 && id[s] < v0 + V;
        approx[s] = valid[s] ? cand_val[(size_t)row * KC + j] : INFINITY;
        score[s] = valid[s] ? exact_score(h, W + (size_t)(id[s] - v0) * H, H, bias, id[s] - v0) : -INFINITY;
        if (!valid[s]) id[s] = INT_MAX;
    }
    // position of each candidate in the exact order (score desc, id asc)
    int pos[2] = {0, 0};
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            const float sj = __shfl_sync(0xffffffffu, score[s2], j);
            const int ij = __shfl_sync(0xffffffffu, id[s2], j);
            pos[0] += !(s2 == 0 && j == lane) && better_(sj, ij, score[0], id[0]);
            pos[1] += !(s2 == 1 && j == lane) && better_(sj, ij, score[1], id[1]);
        }
    }
    const int n_valid = __popc(__ballot_sync(0xffffffffu, valid[0])) + __popc(__ballot_sync(0xffffffffu, valid[1]));
#pragma unroll
    for (int s = 0; s < 2; ++s)
        if (valid[s] && pos[s] < k) {
            topk_val[(size_t)row * k + pos[s]] = score[s];
            topk_idx[(size_t)row * k + pos[s]] = id[s];
        }
    if (lane >= n_valid && lane < k) {            // fewer candidates than k (catalog smaller than k): empty slots as the sweeps write them
        topk_val[(size_t)row * k + lane] = -INFINITY;
        topk_idx[(size_t)row * k + lane] = -1;
    }
    // certificate: can an item OUTSIDE the list beat the exact k-th best candidate?  |exact - bf16| of ANY item is at most
    //   |(h - h~) . w| + |h~ . (w - w~)| + bias split + fp32 accumulation  <=  ||h - h~|| Wmax + ||h~|| Emax + 2^-15 bmax + 2^-16 ||h|| Wmax
    float hh = 0.f, aa = 0.f, tt = 0.f;
    for (int c = lane; c < H; c += 32) {
        const float x = h[c];
        const float xt = __bfloat162float(__float2bfloat16_rn(x));
        hh = fmaf(x, x, hh);
        tt = fmaf(xt, xt, tt);
        aa = fmaf(x - xt, x - xt, aa);
    }
    hh = warp_sum(hh); aa = warp_sum(aa); tt = warp_sum(tt);
    const float E = (sqrtf(aa) * norm_bound[0] + sqrtf(tt) * norm_bound[2] + 3.0517578125e-5f * norm_bound[1] +
                     1.52587890625e-5f * sqrtf(hh) * norm_bound[0]) * (1.0f + 1e-4f);
    const int kth = min(k, n_valid) - 1;
    const unsigned m0 = __ballot_sync(0xffffffffu, valid[0] && pos[0] == kth), m1 = __ballot_sync(0xffffffffu, valid[1] && pos[1] == kth);
    const float x_k = m0 ? __shfl_sync(0xffffffffu, score[0], __ffs(m0) - 1) : __shfl_sync(0xffffffffu, score[1], m1 ? __ffs(m1) - 1 : 0);
    float approx_min = fminf(approx[0], approx[1]);      // the candidates are sorted by bf16 score: the minimum is the last valid entry
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) approx_min = fminf(approx_min, __shfl_xor_sync(0xffffffffu, approx_min, o));
    // what is NOT in the list: items the KC limit cut (bf16 score <= the list's last entry; only when the list is full) and items the
    // sweep's own lists dropped (<= cand_bound, asme_b200_tc_score_candidates; -inf when the list is the true bf16 top KC)
    float outside = n_valid == KC ? approx_min : -INFINITY;
    if (cand_bound != nullptr) outside = fmaxf(outside, cand_bound[row]);
    const bool complete = outside == -INFINITY || (n_valid >= k && !(outside + E >= x_k));
    // the target: its exact score, and its 1-based position among the exact top k (k + 1: not among them)
    if (target != nullptr) {
        const long long t = target[row] - v0;
        const bool owner = t >= 0 && t < V;      // only the shard that owns the target's row writes its score (callers zero-fill)
        const float ts = owner ? exact_score(h, W + (size_t)t * H, H, bias, t) : 0.f;
        const unsigned h0 = __ballot_sync(0xffffffffu, valid[0] && (long long)id[0] - v0 == t && pos[0] < k);
        const unsigned h1 = __ballot_sync(0xffffffffu, valid[1] && (long long)id[1] - v0 == t && pos[1] < k);
        const int hit_pos = h0 ? __shfl_sync(0xffffffffu, pos[0], __ffs(h0) - 1) : __shfl_sync(0xffffffffu, pos[1], h1 ? __ffs(h1) - 1 : 0);
        if (lane == 0) {
            if (target_score && owner) target_score[row] = ts;
            if (rank) rank[row] = (h0 | h1) ? hit_pos + 1 : k + 1;
        }
    }
    if (lane == 0) {
        row_flag[row] = complete ? 0 : 1;
        if (!complete) atomicAdd(n_flagged, 1);
    }
}

extern "C" int asme_b200_topk_rescore(const float* Hrows, int R, int H, const float* W, const float* bias, int v0, int V,
                                      const int32_t* cand_idx, const float* cand_val, const float* cand_bound, int KC, int k,
                                      const float* norm_bound,
                                      const int64_t* target, float* topk_val, int32_t* topk_idx, float* target_score, int32_t* rank,
                                      int32_t* row_flag, int32_t* n_flagged, asme_stream_t stream) {
    ASME_REQUIRE(Hrows && W && cand_idx && cand_val && norm_bound && topk_val && topk_idx && row_flag && n_flagged,
                 "topk_rescore: null argument");
    ASME_REQUIRE(KC >= 1 && KC <= 64 && k >= 1 && k <= KC && k <= 32, "topk_rescore: k=%d KC=%d unsupported (1 <= k <= 32, k <= KC <= 64)", k, KC);
    ASME_REQUIRE(H % 4 == 0 && ((uintptr_t)Hrows & 15) == 0 && ((uintptr_t)W & 15) == 0, "topk_rescore: H=%d / alignment unsupported", H);
    cudaStream_t st = (cudaStream_t)stream;
    ASME_CUDA_OK(cudaMemsetAsync(n_flagged, 0, sizeof(int32_t), st));
    if (R == 0) return ASME_OK;
    topk_rescore_kernel<<<ceil_div(R, 4), 128, 0, st>>>(Hrows, R, H, W, bias, v0, V, cand_idx, cand_val, cand_bound, KC, k, norm_bound, target, topk_val,
                                                        topk_idx, target_score, rank, row_flag, n_flagged);
    ASME_LAUNCH_OK();
    return ASME_OK;
}
