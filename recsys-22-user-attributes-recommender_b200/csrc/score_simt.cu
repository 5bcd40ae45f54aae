// Full-catalog scoring fused with its consumers, fp32 SIMT path (strict-parity mode, H <= 256):
//   * top-k + exact target rank (evaluation)          -- K12 + K15 + K18-K20
//   * log-softmax cross-entropy forward / backward    -- K12 + K16
//   * SASRec positive/negative dots + BCE             -- K13 + K17
// The (rows x V) logits only ever exist as 64x64 tiles in shared memory.
//
// Tile engine: a CTA stages a 64-row block of hidden states and a 64-item block of the table in
// shared memory (row stride H+4 floats -> conflict-free 128-bit reads), every thread owns rows
// ty*4..+3 and the interleaved columns tx+16j, and accumulates with a strictly sequential fmaf
// chain over k -- the same order asme_b200_score_targets uses, so target scores are bit-identical
// to the tile values they are compared with.
#include "common.cuh"

#include <limits.h>

#define TS 64             // tile side (rows and items)
#define TILE_LD (TS + 1)  // logits tile stride
#define SC_THREADS 256
#define ROWS_PER_WARP 8   // 8 warps x 8 rows = 64

__device__ __forceinline__ void stage_tile(const float* __restrict__ src, int nrows_total, int row0, int H, float* __restrict__ dst,
                                           int ld) {
    const int h4 = H / 4;
    for (int i = threadIdx.x; i < TS * h4; i += SC_THREADS) {
        const int r = i / h4, c = (i % h4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < nrows_total) v = ldg4(src + (size_t)(row0 + r) * H + c);
        *reinterpret_cast<float4*>(dst + r * ld + c) = v;
    }
}

// acc[i][j] = sum_k hs[ty*4+i][k] * Ws[tx+16j][k]
__device__ __forceinline__ void tile_dot(const float* __restrict__ hs, const float* __restrict__ Ws, int ld, int H, int ty, int tx,
                                         float (&acc)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k = 0; k < H; k += 4) {
        float4 a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(hs + (ty * 4 + i) * ld + k);
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Ws + (tx + 16 * j) * ld + k);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
            }
    }
}

// logits tile (+bias) -> shared memory
__device__ __forceinline__ void store_logits(const float (&acc)[4][4], const float* __restrict__ bias, int n0, int Vloc, int ty,
                                             int tx, float* __restrict__ tile) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx + 16 * j;
        const float b = (bias && n < Vloc) ? __ldg(bias + n) : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) tile[(ty * 4 + i) * TILE_LD + tx + 16 * j] = acc[i][j] + b;
    }
}

static size_t score_smem_bytes(int H) { return ((size_t)2 * TS * (H + 4) + TS * TILE_LD) * sizeof(float); }

static int check_score(int R, int H, int Vloc) {
    ASME_REQUIRE(R >= 0 && Vloc >= 1, "score: bad shape R=%d Vloc=%d", R, Vloc);
    ASME_REQUIRE(H >= 4 && H <= 256 && H % 4 == 0, "score (fp32 SIMT path): H=%d unsupported (4..256, multiple of 4)", H);
    return ASME_OK;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    { const int _rc = asme_ensure_max_smem((const void*)kernel); if (_rc) return _rc; }
    return ASME_OK;
}

// number of item-splits so that row_tiles * splits fills the machine about twice
static int item_splits(int R, int Vloc) {
    const int row_tiles = ceil_div(R, TS), item_tiles = ceil_div(Vloc, TS);
    int s = ceil_div(2 * ASME_NUM_SMS, row_tiles);
    if (s > item_tiles) s = item_tiles;
    return s < 1 ? 1 : s;
}

// ---------------------------------------------------------------------------------------------
// target scores: sequential fmaf chain, bit-identical to tile_dot
// ---------------------------------------------------------------------------------------------
__global__ void score_targets_kernel(const float* __restrict__ Hrows, int R, int H, const float* __restrict__ W,
                                     const float* __restrict__ bias, int v0, int Vloc, const int64_t* __restrict__ target,
                                     float* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const long long t = target[r] - v0;
    if (t < 0 || t >= Vloc) return;
    const float* h = Hrows + (size_t)r * H;
    const float* w = W + (size_t)t * H;
    float acc = 0.f;
    for (int k = 0; k < H; ++k) acc = fmaf(__ldg(h + k), __ldg(w + k), acc);
    if (bias) acc = acc + __ldg(bias + t);
    out[r] += acc;
}

extern "C" int asme_b200_score_targets(const float* Hrows, int R, int H, const float* W, const float* bias, int v0, int Vloc,
                                       const int64_t* target, float* target_score, asme_stream_t stream) {
    ASME_REQUIRE(Hrows && W && target && target_score, "score_targets: null argument");
    if (R == 0) return ASME_OK;
    score_targets_kernel<<<ceil_div(R, 128), 128, 0, (cudaStream_t)stream>>>(Hrows, R, H, W, bias, v0, Vloc, target, target_score);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ---------------------------------------------------------------------------------------------
// warp-distributed sorted list: lane i holds the i-th best (score desc, id asc)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool better(float v, int id, float v2, int id2) { return v > v2 || (v == v2 && id < id2); }

__device__ __forceinline__ void topk_insert(float& lv, int& li, float cv, int ci, int lane) {
    const unsigned keep = __ballot_sync(0xffffffffu, better(lv, li, cv, ci));
    const int pos = __popc(keep);   // the list is sorted, so the better entries form a prefix
    const float up_v = __shfl_up_sync(0xffffffffu, lv, 1);
    const int up_i = __shfl_up_sync(0xffffffffu, li, 1);
    if (lane == pos) { lv = cv; li = ci; }
    else if (lane > pos) { lv = up_v; li = up_i; }
}

__device__ __forceinline__ void topk_offer(float& lv, int& li, float& thr_v, int& thr_i, float v, int id, bool valid, int lane,
                                           int k) {
    unsigned mask = __ballot_sync(0xffffffffu, valid && better(v, id, thr_v, thr_i));
    while (mask) {
        const int src = __ffs(mask) - 1;
        const float cv = __shfl_sync(0xffffffffu, v, src);
        const int ci = __shfl_sync(0xffffffffu, id, src);
        topk_insert(lv, li, cv, ci, lane);
        mask &= mask - 1;
    }
    thr_v = __shfl_sync(0xffffffffu, lv, k - 1);
    thr_i = __shfl_sync(0xffffffffu, li, k - 1);
}

// ---------------------------------------------------------------------------------------------
// scoring + top-k + rank counts, partial over an item split
// ws layout: pv [splits][R][32] float, pi [splits][R][32] int, pg [splits][R] int, pt [splits][R] int
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SC_THREADS) score_topk_kernel(const float* __restrict__ Hrows, int R, int H,
                                                                const float* __restrict__ W, const float* __restrict__ bias,
                                                                int v0, int Vloc, const int64_t* __restrict__ target,
                                                                const float* __restrict__ target_score, int k,
                                                                int tiles_per_split, float* __restrict__ pv,
                                                                int* __restrict__ pi, int* __restrict__ pg,
                                                                int* __restrict__ pt, const int32_t* __restrict__ tile_list) {
    extern __shared__ __align__(16) float smem[];
    const int ld = H + 4;
    float* hs = smem;
    float* Ws = hs + TS * ld;
    float* tile = Ws + TS * ld;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16, warp = tid / 32, lane = tid % 32;
    const int split = blockIdx.y;
    // tile_list (flagged variant): {count, row tile, row tile, ...} written by flagged_tiles_kernel -- the grid's x dimension is a
    // small constant and CTA x sweeps the listed row tiles x, x + gridDim.x, ...: with nothing flagged every CTA reads the count
    // and leaves (the first version launched one CTA per (row tile, split) just to test the flags: 19 000 empty CTAs at 8192 rows)
    const int n_list = tile_list ? tile_list[0] : (int)gridDim.x;
  for (int ti = blockIdx.x; ti < n_list; ti += gridDim.x) {
    const int m0 = (tile_list ? tile_list[1 + ti] : ti) * TS;
    __syncthreads();                 // the previous row tile's shared memory is no longer read
    const int tile_begin = split * tiles_per_split;
    const int tile_end = min(ceil_div(Vloc, TS), tile_begin + tiles_per_split);

    float lv[ROWS_PER_WARP], thr_v[ROWS_PER_WARP], st[ROWS_PER_WARP];
    int li[ROWS_PER_WARP], thr_i[ROWS_PER_WARP], cg[ROWS_PER_WARP], ct[ROWS_PER_WARP];
    long long tg[ROWS_PER_WARP];
#pragma unroll
    for (int r = 0; r < ROWS_PER_WARP; ++r) {
        lv[r] = -INFINITY; li[r] = INT_MAX; thr_v[r] = -INFINITY; thr_i[r] = INT_MAX; cg[r] = 0; ct[r] = 0;
        const int row = m0 + warp * ROWS_PER_WARP + r;
        st[r] = (row < R && target_score) ? target_score[row] : INFINITY;
        tg[r] = (row < R && target) ? target[row] : -1;
    }
    stage_tile(Hrows, R, m0, H, hs, ld);
    for (int t = tile_begin; t < tile_end; ++t) {
        const int n0 = t * TS;
        __syncthreads();   // previous tile fully consumed
        stage_tile(W, Vloc, n0, H, Ws, ld);
        __syncthreads();
        float acc[4][4];
        tile_dot(hs, Ws, ld, H, ty, tx, acc);
        store_logits(acc, bias, n0, Vloc, ty, tx, tile);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < ROWS_PER_WARP; ++r) {
            const int lr = warp * ROWS_PER_WARP + r;
            if (m0 + lr >= R) break;   // warp-uniform
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int col = lane + 32 * half;
                const int n = n0 + col;
                const bool valid = n < Vloc;
                const float v = tile[lr * TILE_LD + col];
                const int id = v0 + n;
                if (valid && (long long)id != tg[r]) {
                    cg[r] += v > st[r];
                    ct[r] += (v == st[r]) && ((long long)id < tg[r]);
                }
                topk_offer(lv[r], li[r], thr_v[r], thr_i[r], v, id, valid, lane, k);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS_PER_WARP; ++r) {
        const int row = m0 + warp * ROWS_PER_WARP + r;
        if (row >= R) break;
        const size_t o = ((size_t)split * R + row);
        pv[o * 32 + lane] = lv[r];
        pi[o * 32 + lane] = li[r];
        int g = cg[r], tt = ct[r];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            g += __shfl_xor_sync(0xffffffffu, g, s);
            tt += __shfl_xor_sync(0xffffffffu, tt, s);
        }
        if (lane == 0) { pg[o] = g; pt[o] = tt; }
    }
  }
}

// row tiles (of TS rows) that hold a flagged row -> list[0] = count, list[1..] = tile indices in ascending order (one block)
__global__ void __launch_bounds__(1024) flagged_tiles_kernel(const int32_t* __restrict__ row_flag, int R, int32_t* __restrict__ list) {
    __shared__ int warp_count[32];
    __shared__ int base;
    const int n_tiles = ceil_div(R, TS);
    const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    if (tid == 0) base = 0;
    __syncthreads();
    for (int t0 = 0; t0 < n_tiles; t0 += blockDim.x) {
        const int t = t0 + tid;
        bool any = false;
        if (t < n_tiles) {
            const int r1 = min(R, (t + 1) * TS);
            for (int r = t * TS; r < r1; ++r) any |= row_flag[r] != 0;
        }
        const unsigned b = __ballot_sync(0xffffffffu, any);
        if (lane == 0) warp_count[warp] = __popc(b);
        __syncthreads();
        int before = base;
        for (int w = 0; w < warp; ++w) before += warp_count[w];
        if (any) list[1 + before + __popc(b & ((1u << lane) - 1u))] = t;
        __syncthreads();
        if (tid == 0) {
            int total = 0;
            for (int w = 0; w < (int)blockDim.x / 32; ++w) total += warp_count[w];
            base += total;
        }
        __syncthreads();
    }
    if (tid == 0) list[0] = base;
}

// merge `parts` sorted 32-wide (or k-wide) lists per row; one warp per row
__global__ void topk_merge_kernel(const float* __restrict__ pv, const int* __restrict__ pi, const int* __restrict__ pg,
                                  const int* __restrict__ pt, int parts, int R, int width, int k, float* __restrict__ out_v,
                                  int32_t* __restrict__ out_i, int32_t* __restrict__ out_g, int32_t* __restrict__ out_t,
                                  const int32_t* __restrict__ row_flag = nullptr, int32_t* __restrict__ rank_out = nullptr) {
    const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= R) return;
    if (row_flag != nullptr && row_flag[row] == 0) return;      // rows that were not re-run keep what they hold
    float lv = -INFINITY, thr_v = -INFINITY;
    int li = INT_MAX, thr_i = INT_MAX;
    int g = 0, t = 0;
    for (int p = 0; p < parts; ++p) {
        const size_t o = (size_t)p * R + row;
        float v = -INFINITY;
        int id = INT_MAX;
        if (lane < width) { v = pv[o * width + lane]; id = pi[o * width + lane]; }
        const bool valid = lane < width && id != INT_MAX && id >= 0;
        topk_offer(lv, li, thr_v, thr_i, v, id, valid, lane, k);
        if (pg && lane == 0) { g += pg[o]; t += pt[o]; }
    }
    if (lane < k) {
        out_v[(size_t)row * k + lane] = lv;
        out_i[(size_t)row * k + lane] = li == INT_MAX ? -1 : li;
    }
    if (pg && lane == 0) {
        if (out_g) { out_g[row] = g; out_t[row] = t; }
        if (rank_out) rank_out[row] = g + t + 1;
    }
}

extern "C" size_t asme_b200_score_topk_workspace_bytes(int R, int Vloc, int k) {
    (void)k;
    const size_t splits = item_splits(R, Vloc);
    return splits * R * (32 * (sizeof(float) + sizeof(int)) + 2 * sizeof(int));
}

static int score_topk_impl(const float* Hrows, int R, int H, const float* W, const float* bias, int v0, int Vloc,
                           const int64_t* target, const float* target_score, int k, float* topk_val,
                           int32_t* topk_idx, int32_t* n_greater, int32_t* n_tie_lower, void* ws,
                           size_t ws_bytes, const int32_t* row_flag, int32_t* rank_out, asme_stream_t stream);
extern "C" int asme_b200_score_topk_rank(const float* Hrows, int R, int H, const float* W, const float* bias, int v0, int Vloc,
                                         const int64_t* target, const float* target_score, int k, float* topk_val,
                                         int32_t* topk_idx, int32_t* n_greater, int32_t* n_tie_lower, void* ws,
                                         size_t ws_bytes, asme_stream_t stream) {
    ASME_REQUIRE(target == nullptr || (n_greater && n_tie_lower), "score_topk_rank: rank outputs missing");
    return score_topk_impl(Hrows, R, H, W, bias, v0, Vloc, target, target_score, k, topk_val, topk_idx, n_greater, n_tie_lower, ws, ws_bytes,
                           nullptr, nullptr, stream);
}
// The exact sweep for the rows a certificate could not cover (asme_b200_topk_rescore: row_flag != 0): only row tiles holding a
// flagged row run, only flagged rows are written -- topk_val / topk_idx (and rank = exact full rank, when given) of the others
// keep what they hold.  Everything is decided on the device: with no flagged row the launch is a few microseconds of empty CTAs.
#define FLAGGED_GRID_X 8      // row tiles swept at a time by the flagged variant (each over all SMs' worth of catalog splits)
extern "C" size_t asme_b200_score_topk_flagged_workspace_bytes(int R, int Vloc) {
    const size_t splits = ceil_div(Vloc, TS) < ASME_NUM_SMS ? ceil_div(Vloc, TS) : ASME_NUM_SMS;
    return splits * R * (32 * (sizeof(float) + sizeof(int)) + 2 * sizeof(int)) + ((size_t)ceil_div(R, TS) + 4) * sizeof(int32_t);
}
extern "C" int asme_b200_score_topk_flagged(const float* Hrows, int R, int H, const float* W, const float* bias, int v0, int Vloc,
                                            const int64_t* target, const float* target_score, int k, const int32_t* row_flag,
                                            float* topk_val, int32_t* topk_idx, int32_t* rank, void* ws, size_t ws_bytes,
                                            asme_stream_t stream) {
    ASME_REQUIRE(row_flag, "score_topk_flagged: null row_flag");
    ASME_REQUIRE(!rank || target, "score_topk_flagged: a rank needs the target");
    return score_topk_impl(Hrows, R, H, W, bias, v0, Vloc, target, target_score, k, topk_val, topk_idx, nullptr, nullptr, ws, ws_bytes,
                           row_flag, rank, stream);
}
static int score_topk_impl(const float* Hrows, int R, int H, const float* W, const float* bias, int v0, int Vloc,
                           const int64_t* target, const float* target_score, int k, float* topk_val,
                           int32_t* topk_idx, int32_t* n_greater, int32_t* n_tie_lower, void* ws,
                           size_t ws_bytes, const int32_t* row_flag, int32_t* rank_out, asme_stream_t stream) {
    ASME_REQUIRE(Hrows && W && topk_val && topk_idx, "score_topk_rank: null argument");
    ASME_REQUIRE(k >= 1 && k <= 32, "score_topk_rank: k=%d unsupported (1..32)", k);
    ASME_REQUIRE((target == nullptr) == (target_score == nullptr), "score_topk_rank: target and target_score go together");
    int rc = check_score(R, H, Vloc);
    if (rc) return rc;
    if (R == 0) return ASME_OK;
    // few row tiles are expected to be live in the flagged variant: split the catalog over the whole machine for each of them
    const int splits = row_flag ? (ceil_div(Vloc, TS) < ASME_NUM_SMS ? ceil_div(Vloc, TS) : ASME_NUM_SMS) : item_splits(R, Vloc);
    const size_t parts_bytes = (size_t)splits * R * (32 * (sizeof(float) + sizeof(int)) + 2 * sizeof(int));
    if (ws_bytes < parts_bytes + (row_flag ? ((size_t)ceil_div(R, TS) + 4) * sizeof(int32_t) : 0)) {
        asme_set_error("score_topk_rank: workspace too small");
        return ASME_ERR_WORKSPACE;
    }
    const int tiles_per_split = ceil_div(ceil_div(Vloc, TS), splits);
    float* pv = (float*)ws;
    int* pi = (int*)(pv + (size_t)splits * R * 32);
    int* pg = pi + (size_t)splits * R * 32;
    int* pt = pg + (size_t)splits * R;
    const size_t smem = score_smem_bytes(H);
    rc = set_smem(score_topk_kernel, smem);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* tile_list = nullptr;
    int grid_x = ceil_div(R, TS);
    if (row_flag) {          // compact the row tiles that hold a flagged row; a small fixed grid loops over them
        tile_list = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(ws) + parts_bytes);
        flagged_tiles_kernel<<<1, 1024, 0, st>>>(row_flag, R, tile_list);
        ASME_LAUNCH_OK();
        if (grid_x > FLAGGED_GRID_X) grid_x = FLAGGED_GRID_X;
    }
    score_topk_kernel<<<dim3(grid_x, splits), SC_THREADS, smem, st>>>(Hrows, R, H, W, bias, v0, Vloc, target,
                                                                      target_score, k, tiles_per_split, pv, pi, pg, pt, tile_list);
    ASME_LAUNCH_OK();
    topk_merge_kernel<<<ceil_div(R, 4), 128, 0, st>>>(pv, pi, target ? pg : nullptr, pt, splits, R, 32, k, topk_val, topk_idx,
                                                      n_greater, n_tie_lower, row_flag, rank_out);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

extern "C" int asme_b200_topk_merge(const float* vals, const int32_t* idx, int G, int R, int k, float* out_val, int32_t* out_idx,
                                    asme_stream_t stream) {
    ASME_REQUIRE(vals && idx && out_val && out_idx, "topk_merge: null argument");
    ASME_REQUIRE(k >= 1 && k <= 32 && G >= 1, "topk_merge: k=%d G=%d unsupported", k, G);
    if (R == 0) return ASME_OK;
    topk_merge_kernel<<<ceil_div(R, 4), 128, 0, (cudaStream_t)stream>>>(vals, idx, nullptr, nullptr, G, R, k, k, out_val, out_idx,
                                                                        nullptr, nullptr);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ---------------------------------------------------------------------------------------------
// ranking metrics from the 1-based rank: out[4][n_k] += sums of recall, NDCG, MRR, precision
// ---------------------------------------------------------------------------------------------
#define MAX_KS 8
__global__ void ranking_metrics_kernel(const int32_t* __restrict__ rank, int R, const int32_t* __restrict__ ks, int n_k,
                                       float* __restrict__ out) {
    __shared__ float red[4 * MAX_KS][256];
    float acc[4 * MAX_KS];
#pragma unroll
    for (int i = 0; i < 4 * MAX_KS; ++i) acc[i] = 0.f;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const int rk = rank[r];
        const float frk = (float)rk;
#pragma unroll
        for (int i = 0; i < MAX_KS; ++i) {
            if (i < n_k && rk <= ks[i]) {
                acc[0 * MAX_KS + i] += 1.0f;
                acc[1 * MAX_KS + i] += 1.0f / log2f(frk + 1.0f);
                acc[2 * MAX_KS + i] += 1.0f / frk;
                acc[3 * MAX_KS + i] += 1.0f / (float)ks[i];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4 * MAX_KS; ++i) red[i][threadIdx.x] = acc[i];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            for (int i = 0; i < 4 * MAX_KS; ++i) red[i][threadIdx.x] += red[i][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x < 4 * n_k) {
        const int m = threadIdx.x / n_k, i = threadIdx.x % n_k;
        out[m * n_k + i] += red[m * MAX_KS + i][0];
    }
}
extern "C" int asme_b200_ranking_metrics(const int32_t* rank, int R, const int32_t* ks, int n_k, float* out,
                                         asme_stream_t stream) {
    ASME_REQUIRE(rank && ks && out, "ranking_metrics: null argument");
    ASME_REQUIRE(n_k >= 1 && n_k <= MAX_KS, "ranking_metrics: n_k=%d unsupported (1..%d)", n_k, MAX_KS);
    if (R == 0) return ASME_OK;
    ranking_metrics_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(rank, R, ks, n_k, out);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ---------------------------------------------------------------------------------------------
// scoring + cross-entropy forward: per row (max, sumexp, target logit), partial over item splits
// ws: pm [splits][R], ps [splits][R], ptl [splits][R]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SC_THREADS) score_ce_kernel(const float* __restrict__ Hrows, int R, int H,
                                                              const float* __restrict__ W, const float* __restrict__ bias, int v0,
                                                              int Vloc, const int64_t* __restrict__ target, int tiles_per_split,
                                                              float* __restrict__ pm, float* __restrict__ ps,
                                                              float* __restrict__ ptl, const int32_t* __restrict__ n_live) {
    extern __shared__ __align__(16) float smem[];
    const int ld = H + 4;
    float* hs = smem;
    float* Ws = hs + TS * ld;
    float* tile = Ws + TS * ld;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16, warp = tid / 32, lane = tid % 32;
    const int m0 = blockIdx.x * TS;
    const int R_cap = R;                          // partial buffers are laid out for the capacity
    R = asme_live_rows(R, n_live);
    if (m0 >= R) return;
    const int split = blockIdx.y;
    const int tile_begin = split * tiles_per_split;
    const int tile_end = min(ceil_div(Vloc, TS), tile_begin + tiles_per_split);

    float mx[ROWS_PER_WARP], sm[ROWS_PER_WARP], tl[ROWS_PER_WARP];
    long long tg[ROWS_PER_WARP];
#pragma unroll
    for (int r = 0; r < ROWS_PER_WARP; ++r) {
        mx[r] = -INFINITY; sm[r] = 0.f; tl[r] = 0.f;
        const int row = m0 + warp * ROWS_PER_WARP + r;
        tg[r] = row < R ? target[row] - v0 : -1;
    }
    stage_tile(Hrows, R, m0, H, hs, ld);
    for (int t = tile_begin; t < tile_end; ++t) {
        const int n0 = t * TS;
        __syncthreads();
        stage_tile(W, Vloc, n0, H, Ws, ld);
        __syncthreads();
        float acc[4][4];
        tile_dot(hs, Ws, ld, H, ty, tx, acc);
        store_logits(acc, bias, n0, Vloc, ty, tx, tile);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < ROWS_PER_WARP; ++r) {
            const int lr = warp * ROWS_PER_WARP + r;
            if (m0 + lr >= R) break;
            const float a = (n0 + lane < Vloc) ? tile[lr * TILE_LD + lane] : -INFINITY;
            const float b = (n0 + lane + 32 < Vloc) ? tile[lr * TILE_LD + lane + 32] : -INFINITY;
            const float m_new = fmaxf(mx[r], warp_max(fmaxf(a, b)));
            const float e = warp_sum(expf(a - m_new) + expf(b - m_new));
            sm[r] = sm[r] * expf(mx[r] - m_new) + e;
            mx[r] = m_new;
            if (tg[r] >= n0 && tg[r] < n0 + TS) tl[r] = tile[lr * TILE_LD + (int)(tg[r] - n0)];
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < ROWS_PER_WARP; ++r) {
            const int row = m0 + warp * ROWS_PER_WARP + r;
            if (row >= R) break;
            const size_t o = (size_t)split * R_cap + row;
            pm[o] = mx[r]; ps[o] = sm[r]; ptl[o] = tl[r];
        }
    }
}

__global__ void ce_combine_kernel(const float* __restrict__ pm, const float* __restrict__ ps, const float* __restrict__ ptl,
                                  int splits, int R, float* __restrict__ row_max, float* __restrict__ row_sumexp,
                                  float* __restrict__ target_logit, const int32_t* __restrict__ n_live) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= asme_live_rows(R, n_live)) return;
    float m = -INFINITY;
    for (int s = 0; s < splits; ++s) m = fmaxf(m, pm[(size_t)s * R + r]);
    float sum = 0.f, tl = 0.f;
    for (int s = 0; s < splits; ++s) {
        const float ms = pm[(size_t)s * R + r];
        if (ms > -INFINITY) sum += ps[(size_t)s * R + r] * expf(ms - m);
        if (ptl) tl += ptl[(size_t)s * R + r];
    }
    row_max[r] = m; row_sumexp[r] = sum;
    if (target_logit) target_logit[r] = tl;
}

// merge of per-shard softmax statistics (vocab-sharded scoring): pm / ps (G, R) row maxima and sum-exps over each shard's catalog slice
// -> the row maximum and sum-exp over the whole catalog.  The same arithmetic that merges the item splits of one sweep.
extern "C" int asme_b200_ce_combine(const float* pm, const float* ps, int G, int R, float* row_max, float* row_sumexp,
                                    asme_stream_t stream) {
    ASME_REQUIRE(pm && ps && row_max && row_sumexp && G >= 1, "ce_combine: bad argument");
    if (R == 0) return ASME_OK;
    ce_combine_kernel<<<ceil_div(R, 128), 128, 0, (cudaStream_t)stream>>>(pm, ps, nullptr, G, R, row_max, row_sumexp, nullptr, nullptr);
    ASME_LAUNCH_OK();
    return ASME_OK;
}
// out[r] = row_sumexp[r] * exp(row_max[r] - global_max[r]): a shard's sum-exp re-expressed against the all-reduced row maximum
__global__ void ce_rescale_kernel(const float* __restrict__ rsum, const float* __restrict__ rmax, const float* __restrict__ gmax, int R,
                                  float* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float m = rmax[r];
    out[r] = m > -INFINITY ? rsum[r] * expf(m - gmax[r]) : 0.f;
}
extern "C" int asme_b200_ce_rescale(const float* row_sumexp, const float* row_max, const float* global_max, int R, float* out,
                                    asme_stream_t stream) {
    ASME_REQUIRE(row_sumexp && row_max && global_max && out, "ce_rescale: null argument");
    if (R == 0) return ASME_OK;
    ce_rescale_kernel<<<ceil_div(R, 256), 256, 0, (cudaStream_t)stream>>>(row_sumexp, row_max, global_max, R, out);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

extern "C" size_t asme_b200_score_ce_workspace_bytes(int R, int Vloc) {
    return (size_t)item_splits(R, Vloc) * R * 3 * sizeof(float);
}

extern "C" int asme_b200_score_ce_partial(const float* Hrows, int R, int H, const float* W, const float* bias, int v0, int Vloc,
                                          const int64_t* target, float* row_max, float* row_sumexp, float* target_logit,
                                          void* ws, size_t ws_bytes, const int32_t* n_live, asme_stream_t stream) {
    ASME_REQUIRE(Hrows && W && target && row_max && row_sumexp && target_logit, "score_ce_partial: null argument");
    int rc = check_score(R, H, Vloc);
    if (rc) return rc;
    if (R == 0) return ASME_OK;
    if (ws_bytes < asme_b200_score_ce_workspace_bytes(R, Vloc)) {
        asme_set_error("score_ce_partial: workspace too small");
        return ASME_ERR_WORKSPACE;
    }
    const int splits = item_splits(R, Vloc);
    const int tiles_per_split = ceil_div(ceil_div(Vloc, TS), splits);
    float* pm = (float*)ws;
    float* ps = pm + (size_t)splits * R;
    float* ptl = ps + (size_t)splits * R;
    const size_t smem = score_smem_bytes(H);
    rc = set_smem(score_ce_kernel, smem);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    score_ce_kernel<<<dim3(ceil_div(R, TS), splits), SC_THREADS, smem, st>>>(Hrows, R, H, W, bias, v0, Vloc, target,
                                                                             tiles_per_split, pm, ps, ptl, n_live);
    ASME_LAUNCH_OK();
    ce_combine_kernel<<<ceil_div(R, 128), 128, 0, st>>>(pm, ps, ptl, splits, R, row_max, row_sumexp, target_logit, n_live);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// lse[r] = max + log(sumexp); loss_sum += sum_r (lse[r] - target_logit[r])   (single block: deterministic)
// with a device row count: only the live rows, and loss_mean[0] = loss_sum / n_live (0 live rows -> NaN, as the mean over an
// empty set is in nn.CrossEntropyLoss)
__global__ void ce_loss_kernel(const float* __restrict__ row_max, const float* __restrict__ row_sumexp,
                               const float* __restrict__ target_logit, int R, float* __restrict__ lse,
                               float* __restrict__ loss_sum, const int32_t* __restrict__ n_live, float* __restrict__ loss_mean) {
    __shared__ float red[1024];
    float acc = 0.f;
    R = asme_live_rows(R, n_live);
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const float l = row_max[r] + logf(row_sumexp[r]);
        lse[r] = l;
        acc += l - target_logit[r];
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float total = loss_sum[0] + red[0];
        loss_sum[0] = total;
        if (loss_mean) loss_mean[0] = total / (float)R;
    }
}
extern "C" int asme_b200_ce_loss_from_partials(const float* row_max, const float* row_sumexp, const float* target_logit, int R,
                                               float* lse, float* loss_sum, const int32_t* n_live, float* loss_mean,
                                               asme_stream_t stream) {
    ASME_REQUIRE(row_max && row_sumexp && target_logit && lse && loss_sum, "ce_loss: null argument");
    if (R == 0 && !loss_mean) return ASME_OK;
    ce_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(row_max, row_sumexp, target_logit, R, lse, loss_sum, n_live, loss_mean);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ---------------------------------------------------------------------------------------------
// scoring + cross-entropy backward.  dlogit[r,n] = (exp(logit - lse[r]) - [n == target[r]]) * scale
// MODE 0: CTA owns 64 rows, loops over an item split:  dH_partial[split] = dlogit W
// MODE 1: CTA owns 64 items, loops over a row split:   dW_partial[split] = dlogit^T H, dbias_partial
// ---------------------------------------------------------------------------------------------
template <int NC>   // NC = ceil(H / 64) float4 column chunks per thread
__global__ void __launch_bounds__(SC_THREADS) score_ce_bwd_kernel(const float* __restrict__ Hrows, int R, int H,
                                                                  const float* __restrict__ W, const float* __restrict__ bias,
                                                                  int v0, int Vloc, const int64_t* __restrict__ target,
                                                                  const float* __restrict__ lse, float scale, int mode,
                                                                  int tiles_per_split, float* __restrict__ out,
                                                                  float* __restrict__ out_bias, const int32_t* __restrict__ n_live) {
    extern __shared__ __align__(16) float smem[];
    const int ld = H + 4;
    float* hs = smem;
    float* Ws = hs + TS * ld;
    float* tile = Ws + TS * ld;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int split = blockIdx.y;
    const int own0 = blockIdx.x * TS;    // first owned row (mode 0) / item (mode 1)
    const int R_cap = R;                 // partial buffers are laid out for the capacity
    if (n_live != nullptr) {             // row selection: the live count and the mean's 1/n come from the device
        R = asme_live_rows(R, n_live);
        scale = scale / (float)R;
        if (mode == 0 && own0 >= R) return;
    }
    const int n_other = mode == 0 ? Vloc : R;
    // mode 0 streams a contiguous range of item tiles; mode 1 streams hidden-row tiles round robin over the splits, so that the
    // live rows of a capacity-sized selection are spread over all of them
    const int tile_begin = mode == 0 ? split * tiles_per_split : split;
    const int tile_end = mode == 0 ? min(ceil_div(n_other, TS), tile_begin + tiles_per_split) : ceil_div(n_other, TS);
    const int tile_step = mode == 0 ? 1 : (int)gridDim.y;

    float4 acc2[4][NC];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < NC; ++c) acc2[i][c] = make_float4(0.f, 0.f, 0.f, 0.f);
    float bias_acc = 0.f;

    if (mode == 0) stage_tile(Hrows, R, own0, H, hs, ld);
    else stage_tile(W, Vloc, own0, H, Ws, ld);
    for (int t = tile_begin; t < tile_end; t += tile_step) {
        const int oth0 = t * TS;
        const int m0 = mode == 0 ? own0 : oth0;
        const int n0 = mode == 0 ? oth0 : own0;
        __syncthreads();
        if (mode == 0) stage_tile(W, Vloc, oth0, H, Ws, ld);
        else stage_tile(Hrows, R, oth0, H, hs, ld);
        __syncthreads();
        float acc[4][4];
        tile_dot(hs, Ws, ld, H, ty, tx, acc);
        // dlogit -> tile[row][item]
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = m0 + ty * 4 + i;
            const float l = row < R ? lse[row] : 0.f;
            const long long tgt = row < R ? target[row] - v0 : -1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tx + 16 * j;
                float dl = 0.f;
                if (row < R && n < Vloc) {
                    const float logit = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
                    dl = (expf(logit - l) - ((long long)n == tgt ? 1.f : 0.f)) * scale;
                }
                tile[(ty * 4 + i) * TILE_LD + tx + 16 * j] = dl;
            }
        }
        __syncthreads();
        if (mode == 0) {
            // dH[ty*4+i][c] += sum_j tile[ty*4+i][j] * Ws[j][c]
            for (int j = 0; j < TS; ++j) {
                float dl[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) dl[i] = tile[(ty * 4 + i) * TILE_LD + j];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const int col = tx * 4 + 64 * c;
                    if (col < H) {
                        const float4 w = *reinterpret_cast<const float4*>(Ws + j * ld + col);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            acc2[i][c].x = fmaf(dl[i], w.x, acc2[i][c].x); acc2[i][c].y = fmaf(dl[i], w.y, acc2[i][c].y);
                            acc2[i][c].z = fmaf(dl[i], w.z, acc2[i][c].z); acc2[i][c].w = fmaf(dl[i], w.w, acc2[i][c].w);
                        }
                    }
                }
            }
        } else {
            // dW[ty*4+i][c] += sum_r tile[r][ty*4+i] * hs[r][c]
            for (int r = 0; r < TS; ++r) {
                float dl[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) dl[i] = tile[r * TILE_LD + ty * 4 + i];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const int col = tx * 4 + 64 * c;
                    if (col < H) {
                        const float4 h = *reinterpret_cast<const float4*>(hs + r * ld + col);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            acc2[i][c].x = fmaf(dl[i], h.x, acc2[i][c].x); acc2[i][c].y = fmaf(dl[i], h.y, acc2[i][c].y);
                            acc2[i][c].z = fmaf(dl[i], h.z, acc2[i][c].z); acc2[i][c].w = fmaf(dl[i], h.w, acc2[i][c].w);
                        }
                    }
                }
            }
            if (tid < TS) {
                float s = 0.f;
                for (int r = 0; r < TS; ++r) s += tile[r * TILE_LD + tid];
                bias_acc += s;
            }
        }
    }
    const int n_own = mode == 0 ? R : Vloc;
    float* o = out + (size_t)split * (mode == 0 ? R_cap : Vloc) * H;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = own0 + ty * 4 + i;
        if (row >= n_own) continue;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int col = tx * 4 + 64 * c;
            if (col < H) *reinterpret_cast<float4*>(o + (size_t)row * H + col) = acc2[i][c];
        }
    }
    if (mode == 1 && out_bias && tid < TS && own0 + tid < Vloc) out_bias[(size_t)split * Vloc + own0 + tid] = bias_acc;
}

__global__ void split_reduce2_kernel(const float* __restrict__ partial, int splits, long long n, float* __restrict__ out,
                                     int accumulate) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int p = 0; p < splits; ++p) s += partial[(size_t)p * n + i];
    out[i] = accumulate ? out[i] + s : s;
}

static int own_splits(int n_own, int n_other) {
    int s = ceil_div(2 * ASME_NUM_SMS, ceil_div(n_own, TS));
    const int other_tiles = ceil_div(n_other, TS);
    if (s > other_tiles) s = other_tiles;
    return s < 1 ? 1 : s;
}

extern "C" size_t asme_b200_score_ce_bwd_workspace_bytes(int R, int H, int Vloc) {
    const size_t a = (size_t)own_splits(R, Vloc) * R * H;
    const size_t b = (size_t)own_splits(Vloc, R) * ((size_t)Vloc * H + Vloc);
    return (a > b ? a : b) * sizeof(float);
}

extern "C" int asme_b200_score_ce_bwd(const float* Hrows, int R, int H, const float* W, const float* bias, int v0, int Vloc,
                                      const int64_t* target, const float* lse, float scale, float* dH, float* dW, float* dbias,
                                      void* ws, size_t ws_bytes, const int32_t* n_live, asme_stream_t stream) {
    ASME_REQUIRE(Hrows && W && target && lse, "score_ce_bwd: null argument");
    int rc = check_score(R, H, Vloc);
    if (rc) return rc;
    if (R == 0) return ASME_OK;
    if (ws_bytes < asme_b200_score_ce_bwd_workspace_bytes(R, H, Vloc)) {
        asme_set_error("score_ce_bwd: workspace too small");
        return ASME_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = score_smem_bytes(H);
    const int nc = ceil_div(H, 64);
    float* partial = (float*)ws;
#define LAUNCH(NC, MODE, GRID, TPS, OUT, OUTB)                                                                              \
    {                                                                                                                       \
        rc = set_smem(score_ce_bwd_kernel<NC>, smem);                                                                       \
        if (rc) return rc;                                                                                                  \
        score_ce_bwd_kernel<NC><<<GRID, SC_THREADS, smem, st>>>(Hrows, R, H, W, bias, v0, Vloc, target, lse, scale, MODE, TPS, \
                                                                OUT, OUTB, n_live);                                         \
    }
#define LAUNCH_NC(MODE, GRID, TPS, OUT, OUTB)                          \
    switch (nc) {                                                      \
        case 1: LAUNCH(1, MODE, GRID, TPS, OUT, OUTB); break;          \
        case 2: LAUNCH(2, MODE, GRID, TPS, OUT, OUTB); break;          \
        default: LAUNCH(4, MODE, GRID, TPS, OUT, OUTB); break;         \
    }
    if (dH) {
        const int splits = own_splits(R, Vloc);
        const int tps = ceil_div(ceil_div(Vloc, TS), splits);
        dim3 grid(ceil_div(R, TS), splits);
        LAUNCH_NC(0, grid, tps, partial, nullptr);
        ASME_LAUNCH_OK();
        const long long n = (long long)R * H;
        split_reduce2_kernel<<<ceil_div(n, 256), 256, 0, st>>>(partial, splits, n, dH, 0);
        ASME_LAUNCH_OK();
    }
    if (dW) {
        const int splits = own_splits(Vloc, R);
        const int tps = ceil_div(ceil_div(R, TS), splits);
        dim3 grid(ceil_div(Vloc, TS), splits);
        float* pb = partial + (size_t)splits * Vloc * H;
        LAUNCH_NC(1, grid, tps, partial, dbias ? pb : nullptr);
        ASME_LAUNCH_OK();
        const long long n = (long long)Vloc * H;
        split_reduce2_kernel<<<ceil_div(n, 256), 256, 0, st>>>(partial, splits, n, dW, 1);
        ASME_LAUNCH_OK();
        if (dbias) {
            split_reduce2_kernel<<<ceil_div(Vloc, 256), 256, 0, st>>>(pb, splits, Vloc, dbias, 1);
            ASME_LAUNCH_OK();
        }
    }
#undef LAUNCH
#undef LAUNCH_NC
    return ASME_OK;
}

// ---------------------------------------------------------------------------------------------
// SASRec positive / negative dots + BCE
// ---------------------------------------------------------------------------------------------
#define BCE_EPS 1e-24f
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void posneg_fwd_kernel(const float* __restrict__ Hseq, const float* __restrict__ E, const int64_t* __restrict__ pos,
                                  const int64_t* __restrict__ neg, int T, int H, float* __restrict__ pos_logit,
                                  float* __restrict__ neg_logit) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
    if (warp >= T) return;
    const float* h = Hseq + (size_t)warp * H;
    const float* ep = E + (size_t)pos[warp] * H;
    const float* en = E + (size_t)neg[warp] * H;
    float p = 0.f, n = 0.f;
    for (int c = lane * 4; c < H; c += 128) {
        const float4 hv = ldg4(h + c), pv = ldg4(ep + c), nv = ldg4(en + c);
        p += (hv.x * pv.x + hv.y * pv.y) + (hv.z * pv.z + hv.w * pv.w);
        n += (hv.x * nv.x + hv.y * nv.y) + (hv.z * nv.z + hv.w * nv.w);
    }
    p = warp_sum(p);
    n = warp_sum(n);
    if (lane == 0) { pos_logit[warp] = p; neg_logit[warp] = n; }
}

__global__ void bce_sum_kernel(const float* __restrict__ pos_logit, const float* __restrict__ neg_logit,
                               const uint8_t* __restrict__ mask, int T, float* __restrict__ sums) {
    __shared__ float red0[1024], red1[1024];
    float a = 0.f, m = 0.f;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const float mk = mask[t] ? 1.f : 0.f;
        const float lp = logf(sigmoidf_(pos_logit[t]) + BCE_EPS) * mk;
        const float ln = logf(1.0f - sigmoidf_(neg_logit[t]) + BCE_EPS) * mk;
        a += -lp - ln;
        m += mk;
    }
    red0[threadIdx.x] = a; red1[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) { red0[threadIdx.x] += red0[threadIdx.x + s]; red1[threadIdx.x] += red1[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { sums[0] += red0[0]; sums[1] += red1[0]; }
}

extern "C" int asme_b200_posneg_bce_fwd(const float* Hseq, const float* E, const int64_t* pos, const int64_t* neg,
                                        const uint8_t* mask, int T, int H, float* pos_logit, float* neg_logit, float* sums,
                                        asme_stream_t stream) {
    ASME_REQUIRE(Hseq && E && pos && neg && pos_logit && neg_logit, "posneg_bce_fwd: null argument");
    ASME_REQUIRE(H % 4 == 0, "posneg_bce_fwd: H=%d must be a multiple of 4", H);
    if (T == 0) return ASME_OK;
    cudaStream_t st = (cudaStream_t)stream;
    posneg_fwd_kernel<<<ceil_div((long long)T * 32, 256), 256, 0, st>>>(Hseq, E, pos, neg, T, H, pos_logit, neg_logit);
    ASME_LAUNCH_OK();
    if (sums) {
        ASME_REQUIRE(mask, "posneg_bce_fwd: mask required for the loss");
        bce_sum_kernel<<<1, 1024, 0, st>>>(pos_logit, neg_logit, mask, T, sums);
        ASME_LAUNCH_OK();
    }
    return ASME_OK;
}

__global__ void posneg_bwd_kernel(const float* __restrict__ Hseq, const float* __restrict__ E, const int64_t* __restrict__ pos,
                                  const int64_t* __restrict__ neg, const uint8_t* __restrict__ mask, int T, int H,
                                  const float* __restrict__ pos_logit, const float* __restrict__ neg_logit,
                                  const float* __restrict__ sums, float dloss, float* __restrict__ dH,
                                  float* __restrict__ d_pos_rows, float* __restrict__ d_neg_rows) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
    if (warp >= T) return;
    float dp = 0.f, dn = 0.f;
    if (mask[warp]) {
        const float g = dloss / sums[1];
        const float sp = sigmoidf_(pos_logit[warp]), sn = sigmoidf_(neg_logit[warp]);
        dp = -g * sp * (1.f - sp) / (sp + BCE_EPS);
        dn = g * sn * (1.f - sn) / (1.f - sn + BCE_EPS);
    }
    const float* h = Hseq + (size_t)warp * H;
    const float* ep = E + (size_t)pos[warp] * H;
    const float* en = E + (size_t)neg[warp] * H;
    for (int c = lane * 4; c < H; c += 128) {
        const float4 hv = ldg4(h + c), pv = ldg4(ep + c), nv = ldg4(en + c);
        float4 o;
        o.x = dp * pv.x + dn * nv.x; o.y = dp * pv.y + dn * nv.y; o.z = dp * pv.z + dn * nv.z; o.w = dp * pv.w + dn * nv.w;
        *reinterpret_cast<float4*>(dH + (size_t)warp * H + c) = o;
        *reinterpret_cast<float4*>(d_pos_rows + (size_t)warp * H + c) = make_float4(dp * hv.x, dp * hv.y, dp * hv.z, dp * hv.w);
        *reinterpret_cast<float4*>(d_neg_rows + (size_t)warp * H + c) = make_float4(dn * hv.x, dn * hv.y, dn * hv.z, dn * hv.w);
    }
}

extern "C" int asme_b200_posneg_bce_bwd(const float* Hseq, const float* E, const int64_t* pos, const int64_t* neg,
                                        const uint8_t* mask, int T, int H, const float* pos_logit, const float* neg_logit,
                                        const float* sums, float dloss, float* dH, float* d_pos_rows, float* d_neg_rows,
                                        asme_stream_t stream) {
    ASME_REQUIRE(Hseq && E && pos && neg && mask && pos_logit && neg_logit && sums && dH && d_pos_rows && d_neg_rows,
                 "posneg_bce_bwd: null argument");
    ASME_REQUIRE(H % 4 == 0, "posneg_bce_bwd: H=%d must be a multiple of 4", H);
    if (T == 0) return ASME_OK;
    posneg_bwd_kernel<<<ceil_div((long long)T * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        Hseq, E, pos, neg, mask, T, H, pos_logit, neg_logit, sums, dloss, dH, d_pos_rows, d_neg_rows);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ---------------------------------------------------------------------------------------------
// dense-signature ranking metrics (compatibility path of RankingMetric.update(predictions, positive_item_mask,
// metric_mask), metrics/metric.py:57-83): one warp per row, O(I) scan instead of the reference's full argsort.
// out (7, N): recall, precision, dcg, ndcg, mrr, f1 @k and the full rank of the worst relevant item.
// ---------------------------------------------------------------------------------------------
__global__ void dense_ranking_kernel(const float* __restrict__ pred, const int64_t* __restrict__ pos_mask,
                                     const int64_t* __restrict__ metric_mask, int N, int I, int k, float* __restrict__ out) {
    const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= N) return;
    const float* p = pred + (size_t)row * I;
    const int64_t* pm = pos_mask + (size_t)row * I;
    const int64_t* mm = metric_mask ? metric_mask + (size_t)row * I : nullptr;
    const float FMIN = -3.402823466e+38f;
    // pass 1: top-k list, number of relevant items, worst relevant item
    float lv = -INFINITY, thr_v = -INFINITY, worst_v = INFINITY;
    int li = INT_MAX, thr_i = INT_MAX, worst_i = -1, n_rel = 0;
    for (int base = 0; base < I; base += 32) {
        const int j = base + lane;
        const bool valid = j < I;
        float v = valid ? p[j] : -INFINITY;
        if (valid && mm && mm[j] == 0) v = FMIN;
        topk_offer(lv, li, thr_v, thr_i, v, j, valid, lane, k < 32 ? k : 32);
        if (valid && pm[j] != 0) {
            ++n_rel;
            if (worst_i < 0 || v < worst_v || (v == worst_v && j > worst_i)) { worst_v = v; worst_i = j; }   // lowest score, ties -> largest id
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        n_rel += __shfl_xor_sync(0xffffffffu, n_rel, s);
        const float ov = __shfl_xor_sync(0xffffffffu, worst_v, s);
        const int oi = __shfl_xor_sync(0xffffffffu, worst_i, s);
        if (oi >= 0 && (worst_i < 0 || ov < worst_v || (ov == worst_v && oi > worst_i))) { worst_v = ov; worst_i = oi; }
    }
    // pass 2: rank of the worst relevant item = 1 + #items ranked before it
    int before = 0;
    if (worst_i >= 0) {
        for (int j = lane; j < I; j += 32) {
            float v = p[j];
            if (mm && mm[j] == 0) v = FMIN;
            before += (j != worst_i) && better(v, j, worst_v, worst_i);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) before += __shfl_xor_sync(0xffffffffu, before, s);
    }
    // metrics from the true-positive flags of the top-k list
    const int kk = min(k, I);
    const bool hit = lane < kk && li != INT_MAX && pm[li] != 0;
    const unsigned hits = __ballot_sync(0xffffffffu, hit);
    const float tp = (float)__popc(hits);
    float dcg = hit ? 1.0f / log2f((float)lane + 2.0f) : 0.f;
    float idcg = (lane < min(n_rel, k)) ? 1.0f / log2f((float)lane + 2.0f) : 0.f;
    dcg = warp_sum(dcg);
    idcg = warp_sum(idcg);
    if (lane == 0) {
        const float recall = n_rel > 0 ? tp / (float)n_rel : 0.f;
        const float precision = tp / (float)k;
        const int last_hit = hits ? 32 - __clz(hits) : 0;     // mrr.py:31: max(rank * tp)
        const float f1 = (recall + precision) > 0.f ? 2.f * recall * precision / (recall + precision) : 0.f;
        out[0 * (size_t)N + row] = recall;
        out[1 * (size_t)N + row] = precision;
        out[2 * (size_t)N + row] = dcg;
        out[3 * (size_t)N + row] = idcg > 0.f ? dcg / idcg : 0.f;
        out[4 * (size_t)N + row] = last_hit ? 1.0f / (float)last_hit : 0.f;
        out[5 * (size_t)N + row] = f1;
        out[6 * (size_t)N + row] = worst_i >= 0 ? (float)(before + 1) : 0.f;
    }
}

extern "C" int asme_b200_dense_ranking(const float* pred, const int64_t* pos_mask, const int64_t* metric_mask, int N, int I, int k,
                                       float* out, asme_stream_t stream) {
    ASME_REQUIRE(pred && pos_mask && out, "dense_ranking: null argument");
    ASME_REQUIRE(k >= 1 && k <= 32, "dense_ranking: k=%d unsupported (1..32)", k);
    ASME_REQUIRE(I >= 1, "dense_ranking: I=%d", I);
    if (N == 0) return ASME_OK;
    dense_ranking_kernel<<<ceil_div(N, 4), 128, 0, (cudaStream_t)stream>>>(pred, pos_mask, metric_mask, N, I, k, out);
    ASME_LAUNCH_OK();
    return ASME_OK;
}
