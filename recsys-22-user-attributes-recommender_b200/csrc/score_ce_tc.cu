// Backward of the fused scoring + cross-entropy layer on the tensor cores (tcgen05 + TMEM, TMA-fed):
//     dlogit[r, n] = (exp(h_r . w_n + b_n - lse_r) - [n == target_r]) * scale          -- never written to HBM
//     dH = dlogit W        dW += dlogit^T H        dbias += column sums of dlogit
// replaces the autograd of layers.py:105-143 + nn.CrossEntropyLoss (modules/masked_training_module.py:107-111).
//
// Two launches of one kernel, each recomputing the logit tiles it needs in tensor memory (flops are free here):
//   MODE_DH  rows = hidden rows (one 128-row tile per CTA), columns = items streamed in 128-item tiles
//            T = Hrows W_tile^T  ->  G = dlogit (bf16, shared memory)  ->  dH_acc += G W_tile        (W tile consumed MN-major)
//   MODE_DW  rows = items (one 128-item tile per CTA), columns = hidden rows streamed in 128-row tiles
//            T = W_tile Hrows^T  ->  G = dlogit^T                      ->  dW_acc += G Hrows_tile    (Hrows consumed MN-major)
//            dbias = row sums of G (fp32, before the bf16 rounding)
// The transposed product comes from a transposed MMA, not from a data transpose.  Work is split over the streamed dimension
// (<= 148 CTAs); fp32 partials are reduced in a fixed order (deterministic).
//   warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11: two epilogue warpgroups (64 columns each)
#include "common.cuh"
#include "tc_common.cuh"

#include <limits.h>

using namespace tc;

#define CE_THREADS 384
#define CE_TILE 128
#define CE_CHUNK_BYTES (CE_TILE * 128)     // one 64-wide K chunk of a 128-row tile: 16 KB
#define CE_MAX_STAGES 3
#define CE_LOG2E 1.4426950408889634f
enum { MODE_DH = 0, MODE_DW = 1 };

__host__ __device__ constexpr uint32_t ce_idesc(int M, int N, int a_mn, int b_mn) {
    return idesc_bf16_f32(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}
__device__ __forceinline__ uint64_t ce_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t ce_pack(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float ce_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct CeBwdArgs {
    int R, Vloc, v0, H, Kp, kch, stages;
    int n_row_tiles, n_col_tiles, col_tiles_per_split;
    const float* bias;          // (Vloc) or NULL
    const int64_t* target;      // (R) global ids
    const float* lse;           // (R)
    float scale;
    float* partial;             // MODE_DH: [splits][R][H]   MODE_DW: [splits][Vloc][H]
    float* partial_bias;        // MODE_DW: [splits][Vloc] or NULL
    const int32_t* n_live;      // device count of live hidden rows (row selections of capacity R) or NULL; then scale /= n_live
};
struct __align__(16) CeShared {
    float col_lse[CE_TILE];     // MODE_DW: lse of the 128 hidden rows of the current column tile (log2 units)
    int col_tgt[CE_TILE];       //          local target column of those rows (-1: none in this slice)
    uint64_t row_full, t_full, x_full, acc_done;
    uint64_t full[CE_MAX_STAGES], empty[CE_MAX_STAGES];
    uint32_t tmem_base;
};

template <int MODE>
__global__ void __launch_bounds__(CE_THREADS, 1) ce_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmRow,
                                                                  const __grid_constant__ CUtensorMap tmCol, const CeBwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u)   /* pointer arithmetic on the __shared__ array keeps the address space: LDS/STS, not generic LD/ST */;
    const int kch = a.kch, stages = a.stages;
    uint8_t* sRow = smem;                                        // [kch][128 rows][128 B]
    uint8_t* sCol = sRow + (size_t)kch * CE_CHUNK_BYTES;         // [stages][kch][128 rows][128 B]
    uint8_t* sX = sCol + (size_t)stages * kch * CE_CHUNK_BYTES;  // [2][128 rows][128 B]: G tile, 128 columns
    CeShared* sh = reinterpret_cast<CeShared*>(sX + 2 * CE_CHUNK_BYTES);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int row0 = blockIdx.x * CE_TILE;
    const int split = blockIdx.y;
    // hidden rows that are live: the capacity a.R, or the device count of a row selection (the mean's 1/n then comes from it too)
    int R_live = a.R;
    float scale = a.scale;
    if (a.n_live != nullptr) {
        R_live = asme_live_rows(a.R, a.n_live);
        scale = scale / (float)R_live;
        if (MODE == MODE_DH && row0 >= R_live) return;          // whole CTA, before any barrier / TMEM allocation
    }
    // MODE_DH streams a contiguous range of item tiles.  MODE_DW streams the hidden-row tiles ROUND ROBIN over the splits: the
    // live rows of a capacity-sized selection are then spread over all of them, whatever their number
    const int n_stream_tiles = MODE == MODE_DH ? a.n_col_tiles : ceil_div(R_live, CE_TILE);
    const int ct0 = MODE == MODE_DH ? split * a.col_tiles_per_split : split;
    const int ct_step = MODE == MODE_DH ? 1 : (int)gridDim.y;
    const int n_mine = MODE == MODE_DH ? max(0, min(n_stream_tiles, ct0 + a.col_tiles_per_split) - ct0)
                                       : (split < n_stream_tiles ? (n_stream_tiles - split + ct_step - 1) / ct_step : 0);
    const int n_rows_total = MODE == MODE_DH ? R_live : a.Vloc;  // extent of the row dimension
    const int n_rows_layout = MODE == MODE_DH ? a.R : a.Vloc;    // row stride of the partial buffers (capacity)
    uint32_t tmem_cols = 256;
    while ((int)tmem_cols < CE_TILE + a.Kp) tmem_cols <<= 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmRow);
        tma_prefetch_desc(&tmCol);
        mbar_init(&sh->row_full, 1);
        mbar_init(&sh->t_full, 1);
        mbar_init(&sh->x_full, 8);                     // one arrival per epilogue warp
        mbar_init(&sh->acc_done, 1);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&sh->full[s], 1);
            mbar_init(&sh->empty[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&sh->tmem_base, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm_t = sh->tmem_base, tm_acc = sh->tmem_base + CE_TILE;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&sh->row_full, (uint32_t)kch * CE_CHUNK_BYTES);
            for (int c = 0; c < kch; ++c) tma_load_2d(sRow + (size_t)c * CE_CHUNK_BYTES, &tmRow, &sh->row_full, c * CHUNK_K, row0);
            for (int i = 0; i < n_mine; ++i) {
                const int ct = ct0 + i * ct_step;
                const int s = i % stages;
                const uint32_t ph = (uint32_t)(i / stages) & 1u;
                mbar_wait_lean(&sh->empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&sh->full[s], (uint32_t)kch * CE_CHUNK_BYTES);
                for (int c = 0; c < kch; ++c)
                    tma_load_2d(sCol + ((size_t)s * kch + c) * CE_CHUNK_BYTES, &tmCol, &sh->full[s], c * CHUNK_K, ct * CE_TILE);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_t = ce_idesc(CE_TILE, CE_TILE, 0, 0);
            const uint32_t idesc_acc = ce_idesc(CE_TILE, a.Kp, 0, 1);
            mbar_wait_lean(&sh->row_full, 0);
            tc_fence_after();
            auto issue_t = [&](int i) {
                const int s = i % stages;
                const uint32_t ph = (uint32_t)(i / stages) & 1u;
                mbar_wait_lean(&sh->full[s], ph);
                tc_fence_after();
                for (int c = 0; c < kch; ++c) {
                    const uint64_t ad = smem_desc_sw128(smem_u32(sRow + (size_t)c * CE_CHUNK_BYTES));
                    const uint64_t bd = smem_desc_sw128(smem_u32(sCol + ((size_t)s * kch + c) * CE_CHUNK_BYTES));
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        umma_bf16(tm_t, smem_desc_advance(ad, k4 * 32), smem_desc_advance(bd, k4 * 32), idesc_t, (uint32_t)((c | k4) != 0));
                }
                umma_commit(&sh->t_full);
            };
            const int n = n_mine;
            if (n > 0) issue_t(0);
            for (int i = 0; i < n; ++i) {
                const int s = i % stages;
                mbar_wait_lean(&sh->x_full, (uint32_t)i & 1u);
                tc_fence_after();
                // acc[128 x Kp] += G[128 x 128 columns] . ColTile[128 columns x Kp]   (column tile consumed MN-major)
#pragma unroll
                for (int ks = 0; ks < CE_TILE / UMMA_K; ++ks) {
                    const uint64_t xd = smem_desc_advance(smem_desc_sw128(smem_u32(sX + (size_t)(ks / 4) * CE_CHUNK_BYTES)), (ks % 4) * 32);
                    const uint64_t bd = ce_desc_mn(smem_u32(sCol + (size_t)s * kch * CE_CHUNK_BYTES + (size_t)ks * 2048), CE_CHUNK_BYTES);
                    umma_bf16(tm_acc, xd, bd, idesc_acc, (uint32_t)((i | ks) != 0));
                }
                umma_commit(&sh->empty[s]);                  // the column tile may be refilled once these MMAs have read it
                if (i + 1 < n) issue_t(i + 1);               // its commit also covers the accumulate MMAs above
            }
            umma_commit(&sh->acc_done);
        }
    } else if (warp >= 4) {
        const int wg = (warp - 4) / 4;
        const int q4 = warp % 4;
        const int r = q4 * 32 + lane;
        const int et = threadIdx.x - 128;
        const uint32_t lane_addr = ((uint32_t)(q4 * 32) << 16);
        const int row = row0 + r;
        const bool row_ok = row < n_rows_total;
        // per-row constants: MODE_DH: lse and local target of the hidden row;  MODE_DW: bias of the item row
        float row_lse2 = 0.f, row_bias = 0.f;
        int row_tgt = -2;
        if (MODE == MODE_DH) {
            if (row_ok) {
                row_lse2 = a.lse[row] * CE_LOG2E;
                const long long t = a.target[row] - (long long)a.v0;
                row_tgt = (t >= 0 && t < (long long)a.Vloc) ? (int)t : -1;
            }
        } else {
            if (row_ok && a.bias) row_bias = a.bias[row];
        }
        const float sc2 = scale;
        float bias_sum = 0.f;
        const int n = n_mine;
        for (int i = 0; i < n; ++i) {
            const int col_base = (ct0 + i * ct_step) * CE_TILE;
            if (MODE == MODE_DW) {
                // lse / target of the 128 hidden rows forming this tile's columns
                asm volatile("bar.sync 1, 256;" ::: "memory");       // previous tile's constants no longer in use
                if (et < CE_TILE) {
                    const int hr = col_base + et;
                    float l2 = 0.f;
                    int tg = -1;
                    if (hr < R_live) {
                        l2 = a.lse[hr] * CE_LOG2E;
                        const long long t = a.target[hr] - (long long)a.v0;
                        tg = (t >= 0 && t < (long long)a.Vloc) ? (int)t : -1;
                    }
                    sh->col_lse[et] = l2;
                    sh->col_tgt[et] = tg;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            mbar_wait_lean(&sh->t_full, (uint32_t)i & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int c16 = wg * 4; c16 < wg * 4 + 4; ++c16) {
                float t[16], g[16];
                tmem_ld16(tm_t + lane_addr + (uint32_t)(c16 * 16), t);
                tmem_ld_wait();
                const int col0 = col_base + c16 * 16;
                if (MODE == MODE_DH) {
#pragma unroll
                    for (int c = 0; c < 16; c += 4) {
                        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (a.bias) {
                            if (col0 + c + 3 < a.Vloc) bv = __ldg(reinterpret_cast<const float4*>(a.bias + col0 + c));
                            else {
                                if (col0 + c + 0 < a.Vloc) bv.x = __ldg(a.bias + col0 + c + 0);
                                if (col0 + c + 1 < a.Vloc) bv.y = __ldg(a.bias + col0 + c + 1);
                                if (col0 + c + 2 < a.Vloc) bv.z = __ldg(a.bias + col0 + c + 2);
                            }
                        }
                        const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int col = col0 + c + e;
                            const float p = ce_exp2(fmaf(t[c + e] + b4[e], CE_LOG2E, -row_lse2));
                            const float gg = (p - (col == row_tgt ? 1.f : 0.f)) * sc2;
                            g[c + e] = (row_ok && col < a.Vloc) ? gg : 0.f;
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const int cc = c16 * 16 + c;                       // column inside the tile
                        const float p = ce_exp2(fmaf(t[c] + row_bias, CE_LOG2E, -sh->col_lse[cc]));
                        const float gg = (p - (sh->col_tgt[cc] == row ? 1.f : 0.f)) * sc2;
                        g[c] = (row_ok && col0 + c < R_live) ? gg : 0.f;
                        bias_sum += g[c];
                    }
                }
                uint8_t* xchunk = sX + (size_t)(c16 / 4) * CE_CHUNK_BYTES;
#pragma unroll
                for (int u16 = 0; u16 < 2; ++u16) {
                    uint4 w;
                    w.x = ce_pack(g[u16 * 8 + 0], g[u16 * 8 + 1]); w.y = ce_pack(g[u16 * 8 + 2], g[u16 * 8 + 3]);
                    w.z = ce_pack(g[u16 * 8 + 4], g[u16 * 8 + 5]); w.w = ce_pack(g[u16 * 8 + 6], g[u16 * 8 + 7]);
                    *reinterpret_cast<uint4*>(xchunk + sw128_offset(r, (c16 & 3) * 2 + u16)) = w;
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive_warp(&sh->x_full);
        }
        // ---- accumulators -> partial outputs (each warpgroup half of the Kp columns)
        float* out = a.partial + (size_t)split * n_rows_layout * a.H;
        if (n > 0) {
            mbar_wait_lean(&sh->acc_done, 0);
            tc_fence_after();
        }
        const int half = a.Kp / 2;
        for (int c0 = wg * half; c0 < wg * half + half; c0 += 16) {
            float o[16];
            if (n > 0) {
                tmem_ld16(tm_acc + lane_addr + (uint32_t)c0, o);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int c = 0; c < 16; ++c) o[c] = 0.f;
            }
            if (row_ok) {
#pragma unroll
                for (int c = 0; c < 16; c += 4) {
                    if (c0 + c + 3 < a.H) *reinterpret_cast<float4*>(out + (size_t)row * a.H + c0 + c) = make_float4(o[c], o[c + 1], o[c + 2], o[c + 3]);
                    else {
                        for (int e = 0; e < 4; ++e)
                            if (c0 + c + e < a.H) out[(size_t)row * a.H + c0 + c + e] = o[c + e];
                    }
                }
            }
        }
        if (MODE == MODE_DW && a.partial_bias) {
            // the two warpgroups hold the sums of their column halves: combine through shared memory (fixed order)
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (wg == 1) sh->col_lse[r] = bias_sum;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (wg == 0 && row_ok) a.partial_bias[(size_t)split * a.Vloc + row] = bias_sum + sh->col_lse[r];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(sh->tmem_base, tmem_cols);
    }
}

// out[i] (=|+=) sum over splits of partial[split][i]
__global__ void __launch_bounds__(256) ce_reduce_kernel(const float* __restrict__ partial, int splits, long long n, float* __restrict__ out,
                                                        int accumulate) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < splits; ++p) add4(s, ldg4(partial + (size_t)p * n + i));
    float4* o = reinterpret_cast<float4*>(out + i);
    if (accumulate) {
        const float4 c = *o;
        s.x += c.x; s.y += c.y; s.z += c.z; s.w += c.w;
    }
    *o = s;
}
__global__ void ce_reduce_scalar_kernel(const float* __restrict__ partial, int splits, int n, float* __restrict__ out, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int p = 0; p < splits; ++p) s += partial[(size_t)p * n + i];
    out[i] = accumulate ? out[i] + s : s;
}

struct CePlan {
    int row_tiles, col_tiles, splits, tiles_per_split, kch, stages;
    size_t smem;
};
// ``plan_rows`` (0 = n_rows): how many of the rows are expected to be live (row selections pass their capacity and the count on the
// device); only the split of the streamed dimension over CTAs is chosen from it
static int ce_plan(int n_rows, int n_cols, int Kp, CePlan* p, int plan_rows = 0) {
    p->kch = Kp / CHUNK_K;
    p->row_tiles = ceil_div(n_rows, CE_TILE);
    p->col_tiles = ceil_div(n_cols, CE_TILE);
    const int fill_tiles = (plan_rows > 0 && plan_rows < n_rows) ? ceil_div(plan_rows, CE_TILE) : p->row_tiles;
    int splits = ASME_NUM_SMS / fill_tiles;
    if (splits < 1) splits = 1;
    if (splits > p->col_tiles) splits = p->col_tiles;
    p->tiles_per_split = ceil_div(p->col_tiles, splits);
    p->splits = ceil_div(p->col_tiles, p->tiles_per_split);
    const size_t fixed = 1024 + (size_t)p->kch * CE_CHUNK_BYTES + 2 * CE_CHUNK_BYTES + sizeof(CeShared);
    const size_t stage = (size_t)p->kch * CE_CHUNK_BYTES;
    int stages = (int)((227 * 1024 - fixed) / stage);
    if (stages > CE_MAX_STAGES) stages = CE_MAX_STAGES;
    ASME_REQUIRE(stages >= 1, "tc_score_ce_bwd: shared memory budget exceeded (Kp=%d)", Kp);
    p->stages = stages;
    p->smem = fixed + stages * stage;
    return ASME_OK;
}

extern "C" size_t asme_b200_tc_score_ce_bwd_workspace_bytes(int R, int H, int Kp, int Vloc, int plan_rows) {
    CePlan a, b;
    if (R < 1) R = 1;
    if (ce_plan(R, Vloc, Kp, &a, plan_rows) || ce_plan(Vloc, R, Kp, &b)) return 0;
    const size_t wa = (size_t)a.splits * R * H;
    const size_t wb = (size_t)b.splits * ((size_t)Vloc * H + Vloc);
    return (wa > wb ? wa : wb) * sizeof(float);
}

extern "C" int asme_b200_tc_score_ce_bwd(const void* Hb, int R, int H, int Kp, const void* Wb, const float* bias, int v0, int Vloc,
                                         const int64_t* target, const float* lse, float scale, float* dH, float* dW, float* dbias,
                                         void* ws, size_t ws_bytes, const int32_t* n_live, int plan_rows, asme_stream_t stream) {
    ASME_REQUIRE(Hb && Wb && target && lse, "tc_score_ce_bwd: null argument");
    if (!n_live) plan_rows = 0;
    ASME_REQUIRE(Kp >= 64 && Kp <= 256 && Kp % 64 == 0 && H <= Kp && H % 4 == 0, "tc_score_ce_bwd: H=%d Kp=%d unsupported", H, Kp);
    ASME_REQUIRE(!bias || ((uintptr_t)bias & 15) == 0, "tc_score_ce_bwd: bias must be 16-byte aligned");
    if (R == 0) return ASME_OK;
    ASME_REQUIRE(ws_bytes >= asme_b200_tc_score_ce_bwd_workspace_bytes(R, H, Kp, Vloc, plan_rows), "tc_score_ce_bwd: workspace too small");
    CUtensorMap tmH, tmW;
    int rc = asme_tc_make_tmap_bf16(&tmH, Hb, R, Kp, Kp, CE_TILE);
    if (rc) return rc;
    rc = asme_tc_make_tmap_bf16(&tmW, Wb, Vloc, Kp, Kp, CE_TILE);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    CeBwdArgs a{};
    a.R = R; a.Vloc = Vloc; a.v0 = v0; a.H = H; a.Kp = Kp; a.bias = bias; a.target = target; a.lse = lse; a.scale = scale; a.n_live = n_live;
    a.partial = (float*)ws;
    if (dH) {
        CePlan p;
        rc = ce_plan(R, Vloc, Kp, &p, plan_rows);
        if (rc) return rc;
        a.kch = p.kch; a.stages = p.stages; a.n_row_tiles = p.row_tiles; a.n_col_tiles = p.col_tiles; a.col_tiles_per_split = p.tiles_per_split;
        a.partial_bias = nullptr;
        { const int _rc = asme_ensure_max_smem((const void*)ce_bwd_tc_kernel<MODE_DH>); if (_rc) return _rc; }
        ce_bwd_tc_kernel<MODE_DH><<<dim3(p.row_tiles, p.splits), CE_THREADS, p.smem, st>>>(tmH, tmW, a);
        ASME_LAUNCH_OK();
        const long long n = (long long)R * H;
        ce_reduce_kernel<<<ceil_div(n / 4, 256), 256, 0, st>>>(a.partial, p.splits, n, dH, 0);
        ASME_LAUNCH_OK();
    }
    if (dW) {
        CePlan p;
        rc = ce_plan(Vloc, R, Kp, &p);
        if (rc) return rc;
        a.kch = p.kch; a.stages = p.stages; a.n_row_tiles = p.row_tiles; a.n_col_tiles = p.col_tiles; a.col_tiles_per_split = p.tiles_per_split;
        a.partial_bias = dbias ? a.partial + (size_t)p.splits * Vloc * H : nullptr;
        { const int _rc = asme_ensure_max_smem((const void*)ce_bwd_tc_kernel<MODE_DW>); if (_rc) return _rc; }
        ce_bwd_tc_kernel<MODE_DW><<<dim3(p.row_tiles, p.splits), CE_THREADS, p.smem, st>>>(tmW, tmH, a);
        ASME_LAUNCH_OK();
        const long long n = (long long)Vloc * H;
        ce_reduce_kernel<<<ceil_div(n / 4, 256), 256, 0, st>>>(a.partial, p.splits, n, dW, 1);
        ASME_LAUNCH_OK();
        if (dbias) {
            ce_reduce_scalar_kernel<<<ceil_div(Vloc, 256), 256, 0, st>>>(a.partial_bias, p.splits, Vloc, dbias, 1);
            ASME_LAUNCH_OK();
        }
    }
    return ASME_OK;
}
