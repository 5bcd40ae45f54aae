// Position-wise feed-forward block of the encoder as ONE kernel (inference): out = x + W2 gelu(W1 y + b1) + b2, optionally followed
// by the next sublayer's LayerNorm -- the (tokens x 4H) intermediate never leaves the SM.
//
// replaces: PositionwiseFeedForward.forward + the residual of SublayerConnection (models/common/layers/transformer_layers.py:217-220,
//           :120-130) on the evaluation path, i.e. two asme_b200_tc_gemm launches that wrote and re-read a (T, 4H) bf16 tensor
//           (2 x 210 MB per layer at 204 800 tokens x 512) plus the stand-alone LayerNorm of the next block's input.
//
// Arithmetic is the unfused path's, operation for operation: z = fp32 accumulation of bf16 operands, a = bf16(gelu_erf_fast(z + b1)),
// out = (acc2 + b2) + x with acc2 accumulated over the intermediate columns in ascending order -> the fp32 output is bit-identical
// to asme_b200_tc_gemm(act = GELU, bf16 out) followed by asme_b200_tc_gemm(residual).
//
// Persistent CTA (one per SM, 512 threads), tile = 128 tokens, the intermediate dimension in chunks of 64 columns:
//   warp 0      TMA producer: y tiles (double-buffered), per chunk j the rows j*64.. of W1 and the columns j*64.. of W2 (ring of 3)
//   warp 1      MMA issuer:   G1(j): D1[j&1] (128 x 64)  = y (128 x H) . W1_j^T          (H/16 MMAs, N = 64)
//                             G2(j): D2[tile&1] (128 x H) += a_j (128 x 64) . W2_j^T     (4 MMAs, N = H)
//               issued as G1(0) G1(1) G2(0) G1(2) G2(1) ... across tile boundaries, so the tensor core works on G1(j+1) while the
//               GELU warps turn D1(j) into the bf16 operand a_j in shared memory
//   warp 2      TMEM allocator (2 x 64 columns D1 + 2 x 128 columns D2)
//   warp 3      TMA producer of the residual rows: 16-column fp32 blocks of the tile (64-byte swizzle) through a ring of 3 slots --
//               the output warps never wait for a global load (their first version re-read the residual through staging tiles,
//               eight serialised L2 round trips per tile, and was the bottleneck of the whole kernel)
//   warps 4-11  GELU epilogue: two warpgroups, each 32 of a chunk's 64 columns; thread = token row (TMEM lane)
//   warps 12-15 output epilogue of the PREVIOUS tile (D2 is double-buffered): + b2 + residual -> fp32 rows; for the LayerNorm the
//               finished fp32 row is parked in its own accumulator columns (tcgen05.st) and read back once the row statistics
//               are known -> bf16 rows
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_epilogue.cuh"

using namespace tc;

#define FF_THREADS(GW) (256 + 128 * (GW))    // 4 service warps + GW GELU warpgroups + 1 output warpgroup
#define FF_BM 128
#define FF_CH 64                 // intermediate columns per chunk
#define FF_WSLOTS 3
#define FF_WSLOT_BYTES 32768     // W1_j: (H/64) x [64 rows][128 B] at 0, W2_j: [H rows][128 B] at 16 KB
#define FF_A2_BYTES 16384        // [128 rows][128 B]
#define FF_D2_COL0 128           // TMEM: D1 stages at columns 0 / 64, D2 stages at 128 / 256
#define FF_RSLOTS 3
#define FF_RSLOT_BYTES 8192      // [128 rows][16 fp32 = 64 B], 64-byte swizzle

struct FfnFusedArgs {
    int M, H, FF;
    const float* b1;             // (FF)
    const float* b2;             // (H)
    const float* residual;       // (M,H) fp32
    float* out_f32;              // (M,H) or NULL
    const float* ln_gamma;       // (H) or NULL: ln_out = LayerNorm(out) as bf16
    const float* ln_beta;
    __nv_bfloat16* ln_out;       // (M,H) or NULL
    // PRO (block tail): Y is the attention output ctx, x2 = residual + ctx Wo^T + bo and y = LayerNorm(x2; pro_gamma, pro_beta) are
    // computed on chip, x2 is the residual of the feed-forward block
    const float* bo;             // (H)
    const float* pro_gamma;      // (H)
    const float* pro_beta;
};

struct __align__(8) FfnBars {
    uint64_t a1_full[2], a1_empty[2];
    uint64_t w_full[FF_WSLOTS], w_empty[FF_WSLOTS];
    uint64_t d1_full[2], d1_empty[2];
    uint64_t a2_full[2], a2_empty[2];
    uint64_t d2_full[2], d2_empty[2];
    uint64_t r_full[FF_RSLOTS], r_empty[FF_RSLOTS];
    uint64_t d0_full[2], d0_empty[2], y_full[2];       // PRO
    uint32_t tmem_base;
};

// GW = GELU warpgroups (2 or 4), each 64 / GW columns of a chunk.  The GELU stage is latency-bound (TMEM load -> dependent MUFU chains
// -> shared-memory stores, two barrier waits per chunk): with two warpgroups the kernel issued on 40 % of the cycles (ncu) with
// long-scoreboard stalls on top; four warpgroups give every scheduler four GELU warps to switch between.
//
// PRO = the whole tail of an encoder block (transformer_layers.py:181-199 output_linear, :120-130, :217-220, :251-258): the tile that
// arrives by TMA is the attention output ctx; G0 = ctx Wo^T (Wo rides in the weight ring once per tile) lands in one of two D0
// accumulator stages, the output warpgroup turns it into x2 = (G0 + bo) + x -- x by the TMA residual ring -- parks x2 in the D0
// columns, LayerNorms it and writes y as the bf16 A operand OVER the ctx tile; the feed-forward chain runs as before and its output
// epilogue takes the residual x2 from tensor memory.  x2 and y never reach HBM.  G0 of tile i+1 is issued in the middle of tile i, so
// the prologue of the next tile overlaps the chain of the current one; the output accumulator is single-buffered in this mode
// (512 TMEM columns: 2 x 64 D1 + 128 D2 + 2 x 128 D0).  The output warpgroup does the prologue of tile i+1 and then the output of
// tile i; a prologue warpgroup of its own was measured SLOWER (0.230 vs 0.2125 ms for the C5 tail: 640 threads leave 96 registers
// per thread and the GELU warps pay for it).
template <int GW, bool PRO>
__global__ void __launch_bounds__(FF_THREADS(GW), 1) ffn_fused_kernel(const __grid_constant__ CUtensorMap tmY,
                                                                 const __grid_constant__ CUtensorMap tmW1,
                                                                 const __grid_constant__ CUtensorMap tmW2,
                                                                 const __grid_constant__ CUtensorMap tmR,
                                                                 const __grid_constant__ CUtensorMap tmWo, const FfnFusedArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int H = a.H, HC = H / 64, NJ = a.FF / FF_CH;
    const uint32_t a1_bytes = (uint32_t)HC * 16384u;
    uint8_t* sA1 = smem;                                   // [2][HC][128 rows][128 B]
    uint8_t* sW = sA1 + 2 * (size_t)a1_bytes;              // [FF_WSLOTS][FF_WSLOT_BYTES]
    uint8_t* sA2 = sW + (size_t)FF_WSLOTS * FF_WSLOT_BYTES;   // [2][FF_A2_BYTES]
    uint8_t* sRes = sA2 + 2 * FF_A2_BYTES;                 // [FF_RSLOTS][FF_RSLOT_BYTES]
    uint8_t* stage_base = sRes + FF_RSLOTS * FF_RSLOT_BYTES;   // [4 output warps][2 KB]
    FfnBars* bars = reinterpret_cast<FfnBars*>(stage_base + 4 * 2048);
    const int HQ = H / 16;                                 // 16-column blocks of a row
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int m_tiles = (a.M + FF_BM - 1) / FF_BM;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmY);
        tma_prefetch_desc(&tmW1);
        tma_prefetch_desc(&tmW2);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->a1_full[s], 1);
            mbar_init(&bars->a1_empty[s], 1);
            mbar_init(&bars->d1_full[s], 1);
            mbar_init(&bars->d1_empty[s], 4 * GW);
            mbar_init(&bars->a2_full[s], 4 * GW);
            mbar_init(&bars->a2_empty[s], 1);
            mbar_init(&bars->d2_full[s], 1);
            mbar_init(&bars->d2_empty[s], 4);
        }
        for (int s = 0; s < FF_WSLOTS; ++s) {
            mbar_init(&bars->w_full[s], 1);
            mbar_init(&bars->w_empty[s], 1);
        }
        for (int s = 0; s < FF_RSLOTS; ++s) {
            mbar_init(&bars->r_full[s], 1);
            mbar_init(&bars->r_empty[s], 4);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->d0_full[s], 1);
            mbar_init(&bars->d0_empty[s], 4);
            mbar_init(&bars->y_full[s], 4);
        }
        if (PRO) tma_prefetch_desc(&tmWo);
        tma_prefetch_desc(&tmR);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    // TMEM columns: D1 stages at 0 / 64; D2 at 128 (+ 128 for the second stage without PRO); PRO: D0 stages at 256 / 384
    auto d2_col = [&](int t) { return (uint32_t)(FF_D2_COL0 + (PRO ? 0 : t * 128)); };
    auto d0_col = [&](int t) { return (uint32_t)(256 + t * 128); };
    const int jmid = NJ / 2;                               // PRO: G0 of the next tile is issued after G1(jmid) of the current one

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            auto load_a1 = [&](int i, int mt) {
                const int t = i & 1;
                mbar_wait_lean(&bars->a1_empty[t], (((uint32_t)i >> 1) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&bars->a1_full[t], a1_bytes);
                for (int c = 0; c < HC; ++c)
                    tma_load_2d(sA1 + (size_t)t * a1_bytes + (size_t)c * 16384, &tmY, &bars->a1_full[t], c * CHUNK_K, mt * FF_BM);
            };
            int slot = 0;
            uint32_t wph = 0;
            int i = 0;
            auto load_wo = [&]() {       // PRO: the output projection's weight, HC K-chunks of [H rows][128 B], one ring slot
                mbar_wait_lean(&bars->w_empty[slot], wph ^ 1u);
                mbar_arrive_expect_tx(&bars->w_full[slot], (uint32_t)HC * (uint32_t)H * 128u);
                uint8_t* dst = sW + (size_t)slot * FF_WSLOT_BYTES;
                for (int c = 0; c < HC; ++c) tma_load_2d(dst + (size_t)c * H * 128, &tmWo, &bars->w_full[slot], c * CHUNK_K, 0);
                if (++slot == FF_WSLOTS) { slot = 0; wph ^= 1u; }
            };
            if ((int)blockIdx.x < m_tiles) {
                load_a1(0, blockIdx.x);
                if (PRO) load_wo();
            }
            for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++i) {
                const bool has_next = mt + (int)gridDim.x < m_tiles;
                if (has_next) load_a1(i + 1, mt + gridDim.x);      // one tile ahead
                for (int j = 0; j < NJ; ++j) {
                    mbar_wait_lean(&bars->w_empty[slot], wph ^ 1u);
                    mbar_arrive_expect_tx(&bars->w_full[slot], (uint32_t)HC * 8192u + (uint32_t)H * 128u);
                    uint8_t* dst = sW + (size_t)slot * FF_WSLOT_BYTES;
                    for (int c = 0; c < HC; ++c) tma_load_2d(dst + (size_t)c * 8192, &tmW1, &bars->w_full[slot], c * CHUNK_K, j * FF_CH);
                    tma_load_2d(dst + 16384, &tmW2, &bars->w_full[slot], j * FF_CH, 0);
                    if (++slot == FF_WSLOTS) { slot = 0; wph ^= 1u; }
                    if (PRO && j == jmid && has_next) load_wo();
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =====================================
        if (lane == 0) {
            const uint32_t idesc1 = idesc_bf16_f32(FF_BM, FF_CH);
            const uint32_t idesc2 = idesc_bf16_f32(FF_BM, H);
            // second GEMM of global chunk gp (tile index ip, chunk jp of the tile, weight ring slot sp)
            auto issue_g2 = [&](int gp, int ip, int jp, int sp) {
                const int tp = ip & 1, s2 = gp & 1;
                if (jp == 0) {       // the output accumulator: double-buffered without PRO, one stage with it
                    if (PRO) mbar_wait_lean(&bars->d2_empty[0], ((uint32_t)ip & 1u) ^ 1u);
                    else mbar_wait_lean(&bars->d2_empty[tp], (((uint32_t)ip >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                }
                mbar_wait_lean(&bars->a2_full[s2], ((uint32_t)gp >> 1) & 1u);
                tc_fence_after();
                const uint64_t ad = smem_desc_sw128(smem_u32(sA2 + (size_t)s2 * FF_A2_BYTES));
                const uint64_t bd = smem_desc_sw128(smem_u32(sW + (size_t)sp * FF_WSLOT_BYTES + 16384));
                const uint32_t d2 = tmem_base + d2_col(tp);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) umma_bf16(d2, ad + 2 * k4, bd + 2 * k4, idesc2, (uint32_t)((jp | k4) != 0));
                umma_commit(&bars->w_empty[sp]);
                umma_commit(&bars->a2_empty[s2]);
                if (jp == NJ - 1) umma_commit(&bars->d2_full[PRO ? 0 : tp]);
            };
            int g = 0, slot = 0, i = 0;
            uint32_t wph = 0;
            int pg = -1, pi = 0, pj = 0, ps = 0;             // the chunk whose second GEMM is still to be issued
            // PRO: output projection of tile ii (ctx tile in sA1[ii & 1], Wo in the next ring slot) -> D0[ii & 1]
            auto issue_g0 = [&](int ii) {
                const int t0 = ii & 1;
                const uint32_t ph0 = ((uint32_t)ii >> 1) & 1u;
                mbar_wait_lean(&bars->a1_full[t0], ph0);
                mbar_wait_lean(&bars->w_full[slot], wph);
                mbar_wait_lean(&bars->d0_empty[t0], ph0 ^ 1u);
                tc_fence_after();
                const uint32_t d0 = tmem_base + d0_col(t0);
                for (int c = 0; c < HC; ++c) {
                    const uint64_t ad = smem_desc_sw128(smem_u32(sA1 + (size_t)t0 * a1_bytes + (size_t)c * 16384));
                    const uint64_t bd = smem_desc_sw128(smem_u32(sW + (size_t)slot * FF_WSLOT_BYTES + (size_t)c * H * 128));
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) umma_bf16(d0, ad + 2 * k4, bd + 2 * k4, idesc2, (uint32_t)((c | k4) != 0));
                }
                umma_commit(&bars->w_empty[slot]);
                umma_commit(&bars->d0_full[t0]);
                if (++slot == FF_WSLOTS) { slot = 0; wph ^= 1u; }
            };
            if (PRO && (int)blockIdx.x < m_tiles) issue_g0(0);
            for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++i) {
                const int t = i & 1;
                if (PRO) mbar_wait_lean(&bars->y_full[t], ((uint32_t)i >> 1) & 1u);      // y = LayerNorm(x2) written over the ctx tile
                else mbar_wait_lean(&bars->a1_full[t], ((uint32_t)i >> 1) & 1u);
                tc_fence_after();
                for (int j = 0; j < NJ; ++j, ++g) {
                    const int s1 = g & 1;
                    mbar_wait_lean(&bars->w_full[slot], wph);
                    mbar_wait_lean(&bars->d1_empty[s1], (((uint32_t)g >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d1 = tmem_base + (uint32_t)(s1 * FF_CH);
                    for (int c = 0; c < HC; ++c) {
                        const uint64_t ad = smem_desc_sw128(smem_u32(sA1 + (size_t)t * a1_bytes + (size_t)c * 16384));
                        const uint64_t bd = smem_desc_sw128(smem_u32(sW + (size_t)slot * FF_WSLOT_BYTES + (size_t)c * 8192));
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) umma_bf16(d1, ad + 2 * k4, bd + 2 * k4, idesc1, (uint32_t)((c | k4) != 0));
                    }
                    umma_commit(&bars->d1_full[s1]);
                    if (j == NJ - 1) umma_commit(&bars->a1_empty[t]);        // every G1 of the tile has read the y tile by then
                    if (pg >= 0) issue_g2(pg, pi, pj, ps);
                    pg = g; pi = i; pj = j; ps = slot;
                    if (++slot == FF_WSLOTS) { slot = 0; wph ^= 1u; }
                    if (PRO && j == jmid && mt + (int)gridDim.x < m_tiles) issue_g0(i + 1);
                }
            }
            if (pg >= 0) issue_g2(pg, pi, pj, ps);
        }
    } else if (warp == 3) {
        // ===================================== TMA producer of the residual blocks =====================================
        if (lane == 0) {
            int rs = 0;
            uint32_t rph = 0;
            for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
                for (int c = 0; c < HQ; ++c) {
                    mbar_wait_lean(&bars->r_empty[rs], rph ^ 1u);
                    mbar_arrive_expect_tx(&bars->r_full[rs], FF_RSLOT_BYTES);
                    tma_load_2d(sRes + (size_t)rs * FF_RSLOT_BYTES, &tmR, &bars->r_full[rs], c * 16, mt * FF_BM);
                    if (++rs == FF_RSLOTS) { rs = 0; rph ^= 1u; }
                }
            }
        }
    } else if (warp >= 4 && warp < 4 + 4 * GW) {
        // ===================================== GELU epilogue: D1 -> a_j (bf16, K-major operand tile) =====================================
        const int wg = (warp - 4) / 4;
        const int q = warp % 4;
        const int rt = q * 32 + lane;                      // row of the tile == TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        constexpr int GC = FF_CH / GW;                     // columns of a chunk per warpgroup: 32 or 16
        int g = 0;
        for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
            for (int j = 0; j < NJ; ++j, ++g) {
                const int s = g & 1;
                const uint32_t ph = ((uint32_t)g >> 1) & 1u;
                const float* bp = a.b1 + j * FF_CH + wg * GC;
                float4 bq[GC / 4];                           // the chunk's bias, requested before the accumulator is waited for
#pragma unroll
                for (int c = 0; c < GC / 4; ++c) bq[c] = __ldg(reinterpret_cast<const float4*>(bp + 4 * c));
                mbar_wait_lean(&bars->d1_full[s], ph);
                tc_fence_after();
                float v[GC];
                if constexpr (GC == 32) tmem_ld32(lane_addr + (uint32_t)(s * FF_CH + wg * GC), v);
                else tmem_ld16(lane_addr + (uint32_t)(s * FF_CH + wg * GC), v);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive_warp(&bars->d1_empty[s]);       // the accumulator stage may be overwritten
#pragma unroll
                for (int c = 0; c < GC / 4; ++c) {
                    v[4 * c] += bq[c].x; v[4 * c + 1] += bq[c].y; v[4 * c + 2] += bq[c].z; v[4 * c + 3] += bq[c].w;
                }
#pragma unroll
                for (int c = 0; c < GC; ++c) v[c] = gelu_erf_fast(v[c]);
                uint4 w[GC / 8];
#pragma unroll
                for (int u = 0; u < GC / 8; ++u) {
                    w[u].x = pack_bf16(v[8 * u], v[8 * u + 1]); w[u].y = pack_bf16(v[8 * u + 2], v[8 * u + 3]);
                    w[u].z = pack_bf16(v[8 * u + 4], v[8 * u + 5]); w[u].w = pack_bf16(v[8 * u + 6], v[8 * u + 7]);
                }
                mbar_wait_lean(&bars->a2_empty[s], ph ^ 1u);            // G2 of chunk g-2 has read this operand buffer
                uint8_t* dst = sA2 + (size_t)s * FF_A2_BYTES;
#pragma unroll
                for (int u = 0; u < GC / 8; ++u) *reinterpret_cast<uint4*>(dst + sw128_offset(rt, wg * (GC / 8) + u)) = w[u];
                fence_proxy_async_smem();                   // generic-proxy writes -> visible to the tensor core
                mbar_arrive_warp(&bars->a2_full[s]);
            }
        }
    } else if (warp >= 4 + 4 * GW) {
        // ===================================== output warpgroup =====================================
        const int q = warp % 4;
        const int rt = q * 32 + lane;                        // row of the tile == TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const float inv_h = 1.0f / (float)H;
        const uint32_t r_off = (uint32_t)rt * 64u;           // this row inside a residual slot; unit u sits at u ^ ((rt >> 1) & 3)
        const int r_swz = (rt >> 1) & 3;
        int rs = 0;
        uint32_t rph = 0;
        // one 16-column block of the fp32 residual rows from the TMA ring (the slot goes back once the values have arrived)
        auto take_residual = [&](float4 (&rq)[4]) {
            mbar_wait_lean(&bars->r_full[rs], rph);
            const uint8_t* slot = sRes + (size_t)rs * FF_RSLOT_BYTES + r_off;
#pragma unroll
            for (int u = 0; u < 4; ++u) rq[u] = *reinterpret_cast<const float4*>(slot + ((u ^ r_swz) << 4));
            // The slot goes back to the producer only once the shared-memory loads have DELIVERED their values.  An arrive that merely
            // follows the load instructions is not enough -- SYNCS.ARRIVE overtook LDS.128 data that was still on its way and the TMA
            // write of the next block won the race (measured: ~1000 wrong 16-byte units per 5 M, all "three blocks ahead") -- and a
            // register dependency does not survive ptxas scheduling.  membar.cta waits for the loads to be performed.
            __threadfence_block();
            mbar_arrive_warp(&bars->r_empty[rs]);
            if (++rs == FF_RSLOTS) { rs = 0; rph ^= 1u; }
        };
        // PRO: x2 = (ctx Wo^T + bo) + x of tile ii -> parked in its D0 columns; y = LayerNorm(x2) -> bf16 A operand over the ctx tile
        auto prologue = [&](int ii) {
            const int t0 = ii & 1;
            const uint32_t d0 = lane_addr + d0_col(t0);
            mbar_wait_lean(&bars->d0_full[t0], ((uint32_t)ii >> 1) & 1u);
            tc_fence_after();
            float shift = 0.f, s1 = 0.f, s2 = 0.f;
            for (int c = 0; c < HQ; ++c) {
                float v[16];
                tmem_ld16(d0 + (uint32_t)(c * 16), v);
                float4 bq[4], rq[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) bq[u] = __ldg(reinterpret_cast<const float4*>(a.bo + c * 16 + 4 * u));
                take_residual(rq);
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 4; ++u) {        // the unfused epilogue's order: + bias, then + residual
                    v[4 * u] = (v[4 * u] + bq[u].x) + rq[u].x; v[4 * u + 1] = (v[4 * u + 1] + bq[u].y) + rq[u].y;
                    v[4 * u + 2] = (v[4 * u + 2] + bq[u].z) + rq[u].z; v[4 * u + 3] = (v[4 * u + 3] + bq[u].w) + rq[u].w;
                }
                if (c == 0) shift = v[0];
                uint32_t w[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float dlt = v[e] - shift;
                    s1 += dlt;
                    s2 = fmaf(dlt, dlt, s2);
                    w[e] = __float_as_uint(v[e]);
                }
                tmem_st16(d0 + (uint32_t)(c * 16), w);
            }
            tmem_st_wait();
            const float dm = s1 * inv_h;
            const float mean = shift + dm;
            const float rstd = rsqrtf(fmaxf(s2 * inv_h - dm * dm, 0.f) + 1e-5f);
            uint8_t* a1 = sA1 + (size_t)t0 * a1_bytes;
            for (int nn = 0; nn < H; nn += 32) {
                float v[32];
                tmem_ld32(d0 + (uint32_t)nn, v);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    const float4 gm = __ldg(reinterpret_cast<const float4*>(a.pro_gamma + nn + c));
                    const float4 bt = __ldg(reinterpret_cast<const float4*>(a.pro_beta + nn + c));
                    v[c] = fmaf((v[c] - mean) * rstd, gm.x, bt.x); v[c + 1] = fmaf((v[c + 1] - mean) * rstd, gm.y, bt.y);
                    v[c + 2] = fmaf((v[c + 2] - mean) * rstd, gm.z, bt.z); v[c + 3] = fmaf((v[c + 3] - mean) * rstd, gm.w, bt.w);
                }
                uint4 w[4];
                pack32_bf16(v, w);
                uint8_t* chunk = a1 + (size_t)(nn / 64) * 16384;       // K-major operand: 64 columns per 128-byte row
#pragma unroll
                for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(chunk + sw128_offset(rt, (nn % 64) / 8 + u)) = w[u];
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive_warp(&bars->y_full[t0]);
        };
        int i = 0;
        if (PRO && (int)blockIdx.x < m_tiles) prologue(0);
        for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++i) {
            const int t = i & 1;
            if (PRO && mt + (int)gridDim.x < m_tiles) prologue(i + 1);      // its G0 was issued in the middle of this tile
            const int row0 = mt * FF_BM + q * 32;            // the warp's first row
            const int row = row0 + lane;
            const WarpTile wt{stage_base + (size_t)(warp - 4 - 4 * GW) * 2048, lane, a.M - row0};
            const uint32_t d2 = lane_addr + d2_col(t);
            const uint32_t d0 = lane_addr + d0_col(t);
            if (PRO) mbar_wait_lean(&bars->d2_full[0], (uint32_t)i & 1u);
            else mbar_wait_lean(&bars->d2_full[t], ((uint32_t)i >> 1) & 1u);
            tc_fence_after();
            float shift = 0.f, s1 = 0.f, s2 = 0.f;
            for (int c = 0; c < HQ; ++c) {                   // 16 columns: accumulator + b2 + residual (the unfused epilogue's order)
                float v[16];
                tmem_ld16(d2 + (uint32_t)(c * 16), v);
                float4 bq[4], rq[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) bq[u] = __ldg(reinterpret_cast<const float4*>(a.b2 + c * 16 + 4 * u));
                if (PRO) {                                   // the residual x2 waits in tensor memory
                    float x2[16];
                    tmem_ld16(d0 + (uint32_t)(c * 16), x2);
                    tmem_ld_wait();
#pragma unroll
                    for (int u = 0; u < 4; ++u) rq[u] = make_float4(x2[4 * u], x2[4 * u + 1], x2[4 * u + 2], x2[4 * u + 3]);
                } else {
                    take_residual(rq);
                    tmem_ld_wait();
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    v[4 * u] = (v[4 * u] + bq[u].x) + rq[u].x; v[4 * u + 1] = (v[4 * u + 1] + bq[u].y) + rq[u].y;
                    v[4 * u + 2] = (v[4 * u + 2] + bq[u].z) + rq[u].z; v[4 * u + 3] = (v[4 * u + 3] + bq[u].w) + rq[u].w;
                }
                if (a.out_f32 && row < a.M) {                // 64 contiguous bytes per row: stores need no staging (nothing waits for them)
                    float4* dst = reinterpret_cast<float4*>(a.out_f32 + (size_t)row * H + c * 16);
#pragma unroll
                    for (int u = 0; u < 4; ++u) dst[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                }
                if (a.ln_out) {          // row statistics about a sample of the row (no cancellation in s2/H - (s1/H)^2);
                    if (c == 0) shift = v[0];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float dlt = v[e] - shift;
                        s1 += dlt;
                        s2 = fmaf(dlt, dlt, s2);
                    }
                    uint32_t w[16];      // the finished row goes back into its accumulator columns until the statistics are complete
#pragma unroll
                    for (int e = 0; e < 16; ++e) w[e] = __float_as_uint(v[e]);
                    tmem_st16(d2 + (uint32_t)(c * 16), w);
                }
            }
            if (PRO) {                   // x2 of this tile is consumed: G0 of tile i + 2 may overwrite the stage
                tc_fence_before();
                mbar_arrive_warp(&bars->d0_empty[t]);
            }
            if (a.ln_out) {
                tmem_st_wait();
                const float dm = s1 * inv_h;
                const float mean = shift + dm;
                const float rstd = rsqrtf(fmaxf(s2 * inv_h - dm * dm, 0.f) + 1e-5f);
                for (int nn = 0; nn < H; nn += 32) {
                    float v[32];
                    tmem_ld32(d2 + (uint32_t)nn, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        const float4 gm = __ldg(reinterpret_cast<const float4*>(a.ln_gamma + nn + c));
                        const float4 bt = __ldg(reinterpret_cast<const float4*>(a.ln_beta + nn + c));
                        v[c] = fmaf((v[c] - mean) * rstd, gm.x, bt.x); v[c + 1] = fmaf((v[c + 1] - mean) * rstd, gm.y, bt.y);
                        v[c + 2] = fmaf((v[c + 2] - mean) * rstd, gm.z, bt.z); v[c + 3] = fmaf((v[c + 3] - mean) * rstd, gm.w, bt.w);
                    }
                    uint4 w[4];
                    pack32_bf16(v, w);
                    tile_store64(wt, w, reinterpret_cast<uint8_t*>(a.ln_out + (size_t)row0 * H + nn), (size_t)H * 2);
                }
            }
            tc_fence_before();
            mbar_arrive_warp(&bars->d2_empty[PRO ? 0 : t]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// GELU warpgroups of the fused feed-forward kernel.  0 (default): 4 without the fused LayerNorm, 2 with it -- measured at
// 204 800 x 128 x 512: fp32 rows only 0.134 (2) / 0.123 ms (4); with the LayerNorm 0.141 (2) / 0.158 ms (4): the output warpgroup's
// second pass then competes with sixteen GELU warps for issue slots.
static int g_ffn_gw = 0;
extern "C" int asme_b200_tc_ffn_tune(int knob, int value) {
    ASME_REQUIRE(knob == 0 && (value == 0 || value == 2 || value == 4), "tc_ffn_tune: knob 0 (GELU warpgroups) takes 0 (automatic), 2 or 4");
    g_ffn_gw = value;
    return ASME_OK;
}

static int ffn_fused_impl(const void* Y, const void* Wo, const float* bo, const float* pro_gamma, const float* pro_beta, const void* W1,
                          const float* b1, const void* W2, const float* b2, const float* residual, int M, int H, int FF, float* out_f32,
                          const float* ln_gamma, const float* ln_beta, void* ln_out, asme_stream_t stream) {
    const bool pro = Wo != nullptr;
    ASME_REQUIRE(Y && W1 && b1 && W2 && b2 && residual, "tc_ffn_fused: null argument");
    ASME_REQUIRE(!pro || (bo && pro_gamma && pro_beta), "tc_block_tail_fused: the output projection needs its bias and the LayerNorm parameters");
    ASME_REQUIRE(out_f32 || ln_out, "tc_ffn_fused: no output requested");
    ASME_REQUIRE(!ln_out || (ln_gamma && ln_beta), "tc_ffn_fused: ln_out needs gamma and beta");
    ASME_REQUIRE(H == 64 || H == 128, "tc_ffn_fused: hidden size %d unsupported (64, 128)", H);
    ASME_REQUIRE(FF >= 64 && FF % 64 == 0, "tc_ffn_fused: intermediate size %d unsupported (multiple of 64)", FF);
    ASME_REQUIRE(M >= 0, "tc_ffn_fused: M=%d", M);
    if (M == 0) return ASME_OK;
    CUtensorMap tmY, tmW1, tmW2, tmR, tmWo;
    int rc = asme_tc_make_tmap_bf16(&tmY, Y, M, H, H, FF_BM);
    if (rc) return rc;
    rc = asme_tc_make_tmap_bf16(&tmW1, W1, FF, H, H, FF_CH);
    if (rc) return rc;
    rc = asme_tc_make_tmap_bf16(&tmW2, W2, H, FF, FF, H);
    if (rc) return rc;
    rc = asme_tc_make_tmap_f32_16(&tmR, residual, M, H, H, FF_BM);
    if (rc) return rc;
    rc = asme_tc_make_tmap_bf16(&tmWo, pro ? Wo : W2, H, pro ? H : FF, pro ? H : FF, H);
    if (rc) return rc;
    FfnFusedArgs a{};
    a.M = M; a.H = H; a.FF = FF; a.b1 = b1; a.b2 = b2; a.residual = residual; a.out_f32 = out_f32;
    a.ln_gamma = ln_gamma; a.ln_beta = ln_beta; a.ln_out = (__nv_bfloat16*)ln_out;
    a.bo = bo; a.pro_gamma = pro_gamma; a.pro_beta = pro_beta;
    const size_t smem = 1024 + 2 * (size_t)(H / 64) * 16384 + (size_t)FF_WSLOTS * FF_WSLOT_BYTES + 2 * FF_A2_BYTES +
                        (size_t)FF_RSLOTS * FF_RSLOT_BYTES + 4 * 2048 + sizeof(FfnBars);
    const int m_tiles = ceil_div(M, FF_BM);
    const int grid = m_tiles < ASME_NUM_SMS ? m_tiles : ASME_NUM_SMS;
    const int gw = g_ffn_gw ? g_ffn_gw : ((ln_out || pro) ? 2 : 4);
    cudaStream_t st = (cudaStream_t)stream;
#define FFN_LAUNCH(GWV, PROV)                                                                                               \
    {                                                                                                                       \
        { const int _rc = asme_ensure_max_smem((const void*)ffn_fused_kernel<GWV, PROV>); if (_rc) return _rc; }            \
        ffn_fused_kernel<GWV, PROV><<<grid, FF_THREADS(GWV), smem, st>>>(tmY, tmW1, tmW2, tmR, tmWo, a);                    \
    }
    if (gw == 4) { if (pro) FFN_LAUNCH(4, true) else FFN_LAUNCH(4, false) }
    else { if (pro) FFN_LAUNCH(2, true) else FFN_LAUNCH(2, false) }
#undef FFN_LAUNCH
    ASME_LAUNCH_OK();
    return ASME_OK;
}

extern "C" int asme_b200_tc_ffn_fused(const void* Y, const void* W1, const float* b1, const void* W2, const float* b2,
                                      const float* residual, int M, int H, int FF, float* out_f32, const float* ln_gamma,
                                      const float* ln_beta, void* ln_out, asme_stream_t stream) {
    return ffn_fused_impl(Y, nullptr, nullptr, nullptr, nullptr, W1, b1, W2, b2, residual, M, H, FF, out_f32, ln_gamma, ln_beta, ln_out, stream);
}
// The whole tail of an encoder block in one kernel (inference): x2 = residual + ctx Wo^T + bo; y = LayerNorm(x2; pro_gamma, pro_beta);
// out = x2 + W2 gelu(W1 y + b1) + b2 [; ln_out = LayerNorm(out; ln_gamma, ln_beta)].  x2, y and the (M, FF) intermediate stay on chip.
extern "C" int asme_b200_tc_block_tail_fused(const void* ctx, const void* Wo, const float* bo, const float* residual,
                                             const float* pro_gamma, const float* pro_beta, const void* W1, const float* b1,
                                             const void* W2, const float* b2, int M, int H, int FF, float* out_f32,
                                             const float* ln_gamma, const float* ln_beta, void* ln_out, asme_stream_t stream) {
    ASME_REQUIRE(Wo, "tc_block_tail_fused: null output-projection weight");
    return ffn_fused_impl(ctx, Wo, bo, pro_gamma, pro_beta, W1, b1, W2, b2, residual, M, H, FF, out_f32, ln_gamma, ln_beta, ln_out, stream);
}
