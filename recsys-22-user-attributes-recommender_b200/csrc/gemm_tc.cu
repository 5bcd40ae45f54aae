// Dense layers of the encoder on the tensor cores (tcgen05 + TMEM, operands staged by TMA): bf16 operands, fp32 accumulation.
//   replaces: nn.Linear q/k/v/o (models/common/layers/transformer_layers.py:190-199), FFN (:220), FFN modifier
//             (models/common/components/representation_modifier/ffn_modifier.py:19) and their autograd.
//
// Two kernel shapes, both fed by 128-byte-swizzled TMA boxes of 64 bf16 columns:
//   "tall"   C[M,N] = A[M,K] x op(B)   M = tokens (large), N, K <= 256.      forward (B = W, K-major) and dX (B = W, MN-major).
//            One CTA per 128 rows: whole A tile + whole weight in shared memory, one accumulator (N TMEM columns), fused
//            epilogue (bias, GELU, gelu' multiply, counter-based dropout, residual) writing fp32 and / or bf16 rows.
//            Several CTAs are resident per SM, so loads, MMAs and epilogues of different tiles overlap.
//   "wgrad"  dW[N,K] = dY[M,N]^T x X[M,K]   contraction over the tokens.  Both operands are MN-major views of row-major
//            activations; split over token ranges, 4-stage TMA ring, fp32 partials reduced deterministically.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_epilogue.cuh"

using namespace tc;

#define G_BM 128
#define G_THREADS 256          // wgrad kernel: warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4-7 epilogue
#define G_TALL_THREADS 384     // tall kernel: two epilogue warpgroups (warps 4-7 and 8-11) split the tile's columns
#define G_NT 128               // output columns per CTA: 128 TMEM columns and ~33 KB of shared memory -> 4 CTAs per SM, whose loads,
                               // MMAs and (latency-bound) epilogues overlap

// instruction descriptor with selectable operand majors (bit 15: A is MN-major, bit 16: B is MN-major)
__host__ __device__ constexpr uint32_t idesc_major(int M, int N, int a_mn, int b_mn) {
    return idesc_bf16_f32(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}
// MN-major operand, SWIZZLE_128B: rows of the shared-memory chunk are K indices, the 64 contiguous elements of a row are
// MN indices.  LBO = distance between 64-element MN blocks, SBO = distance between 8-row K groups (1024 B, dense).
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct TcGemmArgs {
    int M, N, K;
    const float* bias;                 // (N) or NULL
    int act;                           // ASME_ACT_*
    const __nv_bfloat16* gelu_grad_of; // (M,N) or NULL: multiply by gelu'(z)
    float p_drop;
    unsigned long long seed;
    unsigned int site;
    unsigned int post_site;            // second dropout applied AFTER the residual add (block-end dropout); 0 = none
    const float* residual;             // (M,N) fp32 or NULL
    float* out_f32;                    // (M,N) or NULL
    __nv_bfloat16* out_bf16;           // (M,N) or NULL
    __nv_bfloat16* pre_act_bf16;       // (M,N) or NULL: value before the activation
    int ld_bf16;                       // row stride of out_bf16 (>= N), lets QKV land in a wider buffer
    // fused LayerNorm of the fp32 output rows (persistent kernel, N <= 128 = one column tile): ln_out = LN(out_f32) as bf16,
    // ln_stats (2, M) = row mean / rstd for the backward pass.  The rows are the NEXT sublayer's LayerNorm input, so the
    // stand-alone LayerNorm launch (one more read and write of the activations) disappears.
    const float* ln_gamma;             // (N) or NULL
    const float* ln_beta;
    __nv_bfloat16* ln_out;             // (M,N)
    float* ln_stats;                   // (2,M) or NULL
};

#define G_STAGES 2               // K ring depth: the MMAs of a 64-wide chunk take ~100 cycles, two slots keep TMA ahead; a small
                                 // footprint (<= 65 KB) keeps 3-4 CTAs resident per SM, which is what hides the epilogue latency
struct __align__(8) GemmBars {
    uint64_t full[G_STAGES];
    uint64_t empty[G_STAGES];
    uint64_t done;
    uint32_t tmem_base;
};

// four dropout scales for elements idx4*4 .. idx4*4+3 of a site (same stream as dropout_scale in common.cuh)
__device__ __forceinline__ void dropout_scale4(uint64_t seed, uint32_t site, uint64_t idx4, float p, float inv_keep, float (&s)[4]) {
    dropout_scales4(seed, site, idx4, p, inv_keep, s);
}

// epilogue of one 32-column chunk of the 32 output rows of a warp (shared by both tall kernels; ALL lanes call it): bias,
// pre-activation copy, GELU, gelu' multiply, dropout, residual, block-end dropout, fp32 / bf16 stores.
// `row` = this lane's row (row0 + lane, may be >= M), nc = first column of the chunk.
// FAST: the cheap GELU forms (the caller passes it when no fp32 copy of the result leaves the kernel)
template <bool LN = false, bool FAST = false>
__device__ __forceinline__ void tall_epilogue_chunk(const TcGemmArgs& a, float (&v)[32], const int row, const int nc, const float inv_keep,
                                                    float& ln_s, float& ln_s2, const WarpTile& t) {
    const size_t o = (size_t)row * a.N + nc;                       // element index of v[0] (dropout stream index)
    const size_t o0 = (size_t)(row - t.lane) * a.N + nc;           // the same for the warp's first row
    if (a.bias) {
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + nc + c));
            v[c] += b.x; v[c + 1] += b.y; v[c + 2] += b.z; v[c + 3] += b.w;
        }
    }
    if (a.pre_act_bf16) {
        uint4 w[4];
        pack32_bf16(v, w);
        tile_store64(t, w, reinterpret_cast<uint8_t*>(a.pre_act_bf16 + o0), (size_t)a.N * 2);
    }
    if (a.act == ASME_ACT_GELU) {
        if (FAST) {                      // bf16 only: the 7-instruction form (2.6e-5 absolute, far inside the bf16 rounding)
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = gelu_erf_fast(v[c]);
        } else {                         // an fp32 copy leaves the kernel: erf to 3e-7
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = gelu_erf_as(v[c]);
        }
    }
    if (a.gelu_grad_of) {
        uint4 w[4];
        tile_load64(t, w, reinterpret_cast<const uint8_t*>(a.gelu_grad_of + o0), (size_t)a.N * 2);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t ww[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 z2 = *reinterpret_cast<const __nv_bfloat162*>(&ww[e]);
                const float z0 = __low2float(z2), z1 = __high2float(z2);
                v[8 * u + 2 * e] *= FAST ? gelu_erf_grad_fast(z0) : gelu_erf_grad_as(z0);
                v[8 * u + 2 * e + 1] *= FAST ? gelu_erf_grad_fast(z1) : gelu_erf_grad_as(z1);
            }
        }
    }
    if (a.p_drop > 0.f && a.site) {
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
            float s[4];
            dropout_scale4(asme_seed(a.seed), a.site, (uint64_t)(o + c) >> 2, a.p_drop, inv_keep, s);
            v[c] *= s[0]; v[c + 1] *= s[1]; v[c + 2] *= s[2]; v[c + 3] *= s[3];
        }
    }
    if (a.residual) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {                              // two 16-column halves of 64 bytes per row
            uint4 w[4];
            tile_load64(t, w, reinterpret_cast<const uint8_t*>(a.residual + o0 + 16 * h), (size_t)a.N * 4);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[16 * h + 4 * u] += __uint_as_float(w[u].x); v[16 * h + 4 * u + 1] += __uint_as_float(w[u].y);
                v[16 * h + 4 * u + 2] += __uint_as_float(w[u].z); v[16 * h + 4 * u + 3] += __uint_as_float(w[u].w);
            }
        }
    }
    if (a.p_drop > 0.f && a.post_site) {
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
            float s[4];
            dropout_scale4(asme_seed(a.seed), a.post_site, (uint64_t)(o + c) >> 2, a.p_drop, inv_keep, s);
            v[c] *= s[0]; v[c + 1] *= s[1]; v[c + 2] *= s[2]; v[c + 3] *= s[3];
        }
    }
    if (a.out_f32) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint4 w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                w[u] = make_uint4(__float_as_uint(v[16 * h + 4 * u]), __float_as_uint(v[16 * h + 4 * u + 1]),
                                  __float_as_uint(v[16 * h + 4 * u + 2]), __float_as_uint(v[16 * h + 4 * u + 3]));
            tile_store64(t, w, reinterpret_cast<uint8_t*>(a.out_f32 + o0 + 16 * h), (size_t)a.N * 4);
        }
    }
    if (LN) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            ln_s += v[c];
            ln_s2 = fmaf(v[c], v[c], ln_s2);
        }
    }
    if (a.out_bf16) {
        uint4 w[4];
        pack32_bf16(v, w);
        tile_store64(t, w, reinterpret_cast<uint8_t*>(a.out_bf16 + (size_t)(row - t.lane) * a.ld_bf16 + nc), (size_t)a.ld_bf16 * 2);
    }
}

// CTA (m_tile, n_tile): 128 rows x NT (<= G_NT) columns; K streamed in 64-wide chunks through a ring of `stages` slots
// (slot = A chunk 16 KB + B chunk NT x 128 B [K-major] or ceil(NT/64) x 8 KB [MN-major]).
template <bool B_MN>
__global__ void __launch_bounds__(G_TALL_THREADS) tc_gemm_tall_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmB, const TcGemmArgs a,
                                                                  int stages) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u)   /* pointer arithmetic on the __shared__ array keeps the address space: LDS/STS, not generic LD/ST */;
    const int kch = a.K / CHUNK_K;
    const int n0 = blockIdx.y * G_NT;
    const int NT = min(G_NT, a.N - n0);                   // columns of this CTA (multiple of 32)
    const int nblk = (NT + 63) / 64;
    const size_t a_bytes = (size_t)G_BM * 128;
    const size_t b_bytes = B_MN ? (size_t)nblk * 64 * 128 : (size_t)min(G_NT, a.N) * 128;   // = bytes the TMA boxes deliver (OOB rows are zero-filled)
    const size_t b_slot = (size_t)G_NT * 128;             // slot stride sized for the widest tile: keeps every chunk 1024-aligned
    const size_t slot = a_bytes + b_slot;
    GemmBars* bars = reinterpret_cast<GemmBars*>(smem + (size_t)stages * slot);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int m0 = blockIdx.x * G_BM;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < NT) tmem_cols <<= 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        mbar_init(&bars->done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&bars->tmem_base, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            for (int c = 0; c < kch; ++c) {
                const int s = c % stages;
                const uint32_t ph = (uint32_t)(c / stages) & 1u;
                mbar_wait_lean(&bars->empty[s], ph ^ 1u);
                uint8_t* sA = smem + (size_t)s * slot;
                uint8_t* sB = sA + a_bytes;
                mbar_arrive_expect_tx(&bars->full[s], (uint32_t)(a_bytes + b_bytes));
                tma_load_2d(sA, &tmA, &bars->full[s], c * CHUNK_K, m0);
                if (B_MN) {
                    for (int j = 0; j < nblk; ++j) tma_load_2d(sB + (size_t)j * 64 * 128, &tmB, &bars->full[s], n0 + j * 64, c * CHUNK_K);
                } else {
                    tma_load_2d(sB, &tmB, &bars->full[s], c * CHUNK_K, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = idesc_major(G_BM, NT, 0, B_MN ? 1 : 0);
            for (int c = 0; c < kch; ++c) {
                const int s = c % stages;
                const uint32_t ph = (uint32_t)(c / stages) & 1u;
                mbar_wait_lean(&bars->full[s], ph);
                tc_fence_after();
                uint8_t* sA = smem + (size_t)s * slot;
                uint8_t* sB = sA + a_bytes;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const uint64_t ad = smem_desc_advance(smem_desc_sw128(smem_u32(sA)), k4 * 32);
                    uint64_t bd;
                    if (B_MN) bd = smem_desc_mn_sw128(smem_u32(sB + (size_t)k4 * 2048), 64 * 128);   // 16 K rows = 2 KB per step
                    else bd = smem_desc_advance(smem_desc_sw128(smem_u32(sB)), k4 * 32);
                    umma_bf16(tmem_base, ad, bd, idesc, (uint32_t)((c | k4) != 0));
                }
                umma_commit(&bars->empty[s]);
            }
            umma_commit(&bars->done);
        }
    } else if (warp >= 4) {
        const int wg = (warp - 4) / 4;                       // both warpgroups own all 128 rows, each half of the 32-column chunks
        const int q = warp % 4;
        const int row = m0 + q * 32 + lane;
        const bool row_ok = row < a.M;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const float inv_keep = a.p_drop > 0.f ? 1.0f / (1.0f - a.p_drop) : 1.0f;
        const int n_chunks = NT / 32;
        const int c_lo = wg == 0 ? 0 : (n_chunks + 1) / 2, c_hi = wg == 0 ? (n_chunks + 1) / 2 : n_chunks;
        mbar_wait_lean(&bars->done, 0);
        tc_fence_after();
        for (int nn = c_lo * 32; nn < c_hi * 32; nn += 32) {
            float v[32];
            tmem_ld32(lane_addr + (uint32_t)nn, v);
            tmem_ld_wait();
            float unused_s = 0.f, unused_s2 = 0.f;
            const WarpTile wt{nullptr, lane, a.M - (row - lane)};
            if (a.out_f32) tall_epilogue_chunk<false, false>(a, v, row, n0 + nn, inv_keep, unused_s, unused_s2, wt);
            else tall_epilogue_chunk<false, true>(a, v, row, n0 + nn, inv_keep, unused_s, unused_s2, wt);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Persistent tall kernel (default).  The kernel above pays one TMA round trip, one TMEM allocation and one pipeline fill per
// 128-row tile and is latency-bound (15 % issue utilisation, 1.35 waves of short-lived CTAs).  Here a CTA owns one column tile
// for the whole launch: the weight tile is loaded ONCE and stays in shared memory, activation tiles stream through a ring of
// K-chunk slots filled ahead across tile boundaries, and two accumulator stages in tensor memory let the MMAs of tile i+1
// overlap the epilogue of tile i.
// ------------------------------------------------------------------------------------------------------------
#define GP_MAX_STAGES 6
struct __align__(8) GemmPBars {
    uint64_t b_full;
    uint64_t full[GP_MAX_STAGES];
    uint64_t empty[GP_MAX_STAGES];
    uint64_t tfull[2];
    uint64_t tempty[2];
    uint32_t tmem_base;
    float ln_part[2][2][G_BM];         // fused LayerNorm: [warpgroup][sum | sum of squares][row of the tile]
};

template <bool B_MN, bool LN>
__global__ void __launch_bounds__(G_TALL_THREADS, 2) tc_gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                     const __grid_constant__ CUtensorMap tmB, const TcGemmArgs a,
                                                                     int stages) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int kch = a.K / CHUNK_K;
    const int n0 = blockIdx.y * G_NT;
    const int NT = min(G_NT, a.N - n0);
    const int nblk = (NT + 63) / 64;
    const size_t a_bytes = (size_t)G_BM * 128;
    const size_t b_bytes = B_MN ? (size_t)nblk * 64 * 128 : (size_t)min(G_NT, a.N) * 128;   // bytes the TMA boxes of one K chunk deliver
    const size_t b_slot = (size_t)G_NT * 128;
    uint8_t* sB = smem;                                   // [kch][b_slot]: the weight tile, resident for the whole launch
    uint8_t* sA = sB + (size_t)kch * b_slot;              // [stages][a_bytes]
    uint8_t* stage_base = sA + (size_t)stages * a_bytes;       // [8 epilogue warps][2 KB] staging tiles of the coalesced epilogue
    GemmPBars* bars = reinterpret_cast<GemmPBars*>(stage_base + 8 * 2048);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int m_tiles = (a.M + G_BM - 1) / G_BM;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        mbar_init(&bars->b_full, 1);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->tfull[s], 1);
            mbar_init(&bars->tempty[s], 8);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&bars->tmem_base, 2 * G_NT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars->b_full, (uint32_t)(kch * b_bytes));
            for (int c = 0; c < kch; ++c) {
                uint8_t* dst = sB + (size_t)c * b_slot;
                if (B_MN) {
                    for (int j = 0; j < nblk; ++j) tma_load_2d(dst + (size_t)j * 64 * 128, &tmB, &bars->b_full, n0 + j * 64, c * CHUNK_K);
                } else {
                    tma_load_2d(dst, &tmB, &bars->b_full, c * CHUNK_K, n0);
                }
            }
            int s = 0;
            uint32_t ph = 0;
            for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
                for (int c = 0; c < kch; ++c) {
                    mbar_wait_lean(&bars->empty[s], ph ^ 1u);
                    mbar_arrive_expect_tx(&bars->full[s], (uint32_t)a_bytes);
                    tma_load_2d(sA + (size_t)s * a_bytes, &tmA, &bars->full[s], c * CHUNK_K, mt * G_BM);
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = idesc_major(G_BM, NT, 0, B_MN ? 1 : 0);
            mbar_wait_lean(&bars->b_full, 0);
            tc_fence_after();
            int s = 0, i = 0;
            uint32_t ph = 0;
            for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++i) {
                const int as = i & 1;
                mbar_wait_lean(&bars->tempty[as], ((uint32_t)(i >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * G_NT);
                for (int c = 0; c < kch; ++c) {
                    mbar_wait_lean(&bars->full[s], ph);
                    tc_fence_after();
                    const uint8_t* pa = sA + (size_t)s * a_bytes;
                    const uint8_t* pb = sB + (size_t)c * b_slot;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const uint64_t ad = smem_desc_advance(smem_desc_sw128(smem_u32(pa)), k4 * 32);
                        uint64_t bd;
                        if (B_MN) bd = smem_desc_mn_sw128(smem_u32(pb + (size_t)k4 * 2048), 64 * 128);
                        else bd = smem_desc_advance(smem_desc_sw128(smem_u32(pb)), k4 * 32);
                        umma_bf16(d_tmem, ad, bd, idesc, (uint32_t)((c | k4) != 0));
                    }
                    umma_commit(&bars->empty[s]);
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
                umma_commit(&bars->tfull[as]);
            }
        }
    } else if (warp >= 4) {
        const int wg = (warp - 4) / 4;                       // both warpgroups own all 128 rows, each half of the 32-column chunks
        const int q = warp % 4;
        const float inv_keep = a.p_drop > 0.f ? 1.0f / (1.0f - a.p_drop) : 1.0f;
        const bool fast_act = a.out_f32 == nullptr;          // bf16-only results: cheap GELU forms (two specialised epilogue bodies)
        const int n_chunks = NT / 32;
        const int c_lo = wg == 0 ? 0 : (n_chunks + 1) / 2, c_hi = wg == 0 ? (n_chunks + 1) / 2 : n_chunks;
        int i = 0;
        for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++i) {
            const int as = i & 1;
            const int row = mt * G_BM + q * 32 + lane;
            const bool row_ok = row < a.M;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * G_NT);
            mbar_wait_lean(&bars->tfull[as], (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            if (c_lo == c_hi) {                              // a warpgroup without columns (NT = 32) still hands the stage back
                tc_fence_before();
                if (lane == 0) mbar_arrive(&bars->tempty[as]);
                if (LN) {                                    // ... and takes part in the statistics exchange with zeros
                    bars->ln_part[wg][0][q * 32 + lane] = 0.f;
                    bars->ln_part[wg][1][q * 32 + lane] = 0.f;
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    asm volatile("bar.sync 2, 256;" ::: "memory");
                }
                continue;
            }
            float ln_s = 0.f, ln_s2 = 0.f;
            const WarpTile wt{stage_base + (size_t)(warp - 4) * 2048, lane, a.M - (row - lane)};
            for (int nn = c_lo * 32; nn < c_hi * 32; nn += 32) {
                float v[32];
                tmem_ld32(lane_addr + (uint32_t)nn, v);
                tmem_ld_wait();
                if (nn + 32 >= c_hi * 32) {                  // last TMEM read of this stage by this warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->tempty[as]);
                }
                if (fast_act) tall_epilogue_chunk<LN, true>(a, v, row, n0 + nn, inv_keep, ln_s, ln_s2, wt);
                else tall_epilogue_chunk<LN, false>(a, v, row, n0 + nn, inv_keep, ln_s, ln_s2, wt);
            }
            if (LN) {
                // row statistics: the two warpgroups hold disjoint column ranges of the same rows -> exchange the partial sums,
                // then every thread normalises the columns it wrote (re-read from L1 / L2: its own stores of a moment ago)
                const int rt = q * 32 + lane;
                bars->ln_part[wg][0][rt] = ln_s;
                bars->ln_part[wg][1][rt] = ln_s2;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float inv_n = 1.0f / (float)a.N;
                const float mean = (ln_s + bars->ln_part[wg ^ 1][0][rt]) * inv_n;
                const float var = fmaxf((ln_s2 + bars->ln_part[wg ^ 1][1][rt]) * inv_n - mean * mean, 0.f);
                const float rstd = rsqrtf(var + 1e-5f);
                asm volatile("bar.sync 2, 256;" ::: "memory");       // ln_part may be overwritten by the next tile
                if (row_ok) {
                    if (wg == 0 && a.ln_stats) { a.ln_stats[row] = mean; a.ln_stats[(size_t)a.M + row] = rstd; }
                    for (int nn = c_lo * 32; nn < c_hi * 32; nn += 32) {
                        const float* xr = a.out_f32 + (size_t)row * a.N + nn;
                        __nv_bfloat16* yr = a.ln_out + (size_t)row * a.N + nn;
#pragma unroll
                        for (int c = 0; c < 32; c += 8) {
                            const float4 x0 = *reinterpret_cast<const float4*>(xr + c), x1 = *reinterpret_cast<const float4*>(xr + c + 4);
                            const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.ln_gamma + nn + c)), g1 = __ldg(reinterpret_cast<const float4*>(a.ln_gamma + nn + c + 4));
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.ln_beta + nn + c)), b1 = __ldg(reinterpret_cast<const float4*>(a.ln_beta + nn + c + 4));
                            uint4 w;
                            w.x = pack_bf16(fmaf((x0.x - mean) * rstd, g0.x, b0.x), fmaf((x0.y - mean) * rstd, g0.y, b0.y));
                            w.y = pack_bf16(fmaf((x0.z - mean) * rstd, g0.z, b0.z), fmaf((x0.w - mean) * rstd, g0.w, b0.w));
                            w.z = pack_bf16(fmaf((x1.x - mean) * rstd, g1.x, b1.x), fmaf((x1.y - mean) * rstd, g1.y, b1.y));
                            w.w = pack_bf16(fmaf((x1.z - mean) * rstd, g1.z, b1.z), fmaf((x1.w - mean) * rstd, g1.w, b1.w));
                            *reinterpret_cast<uint4*>(yr + c) = w;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * G_NT);
    }
}

static int g_gemm_variant = 1;      // 1: persistent kernel (default), 0: one CTA per tile (kept for A/B measurements)
extern "C" int asme_b200_tc_gemm_tune(int knob, int value) {
    ASME_REQUIRE(knob == 0 && (value == 0 || value == 1), "tc_gemm_tune: knob 0 (tall-kernel variant) takes 0 or 1");
    g_gemm_variant = value;
    return ASME_OK;
}

static int tc_gemm_impl(const void* A, const void* B, int M, int N, int K, int b_is_kn, const float* bias, int act,
                        const void* gelu_grad_of, float p_drop, unsigned long long seed, unsigned int site,
                        unsigned int post_site, const float* residual, float* out_f32, void* out_bf16, int ld_bf16,
                        void* pre_act_bf16, const float* ln_gamma, const float* ln_beta, void* ln_out, float* ln_stats,
                        asme_stream_t stream) {
    ASME_REQUIRE(A && B && (out_f32 || out_bf16), "tc_gemm: null argument");
    ASME_REQUIRE(M >= 0 && N >= 32 && N % 32 == 0, "tc_gemm: N=%d unsupported (multiple of 32)", N);
    ASME_REQUIRE(K >= 64 && K % 64 == 0, "tc_gemm: K=%d unsupported (multiple of 64)", K);
    ASME_REQUIRE(!out_bf16 || (ld_bf16 >= N && ld_bf16 % 8 == 0), "tc_gemm: ld_bf16=%d", ld_bf16);
    const bool ln = ln_gamma != nullptr;
    ASME_REQUIRE(!ln || (ln_beta && ln_out && out_f32 && N <= G_NT), "tc_gemm_ln: needs beta, ln_out, out_f32 and N <= %d (one column tile)", G_NT);
    if (M == 0) return ASME_OK;
    CUtensorMap tmA, tmB;
    int rc = asme_tc_make_tmap_bf16(&tmA, A, M, K, K, G_BM);
    if (rc) return rc;
    // b_is_kn = 0: B is (N,K) row-major (nn.Linear weight, C = A B^T): box 64 x min(N,G_NT) rows.
    // b_is_kn = 1: B is (K,N) row-major (C = A B): box 64 columns x 64 K rows.
    rc = b_is_kn ? asme_tc_make_tmap_bf16(&tmB, B, K, N, N, 64) : asme_tc_make_tmap_bf16(&tmB, B, N, K, K, N < G_NT ? N : G_NT);
    if (rc) return rc;
    TcGemmArgs a{};
    a.M = M; a.N = N; a.K = K; a.bias = bias; a.act = act; a.gelu_grad_of = (const __nv_bfloat16*)gelu_grad_of;
    a.p_drop = p_drop; a.seed = seed; a.site = site; a.post_site = post_site; a.residual = residual; a.out_f32 = out_f32;
    a.out_bf16 = (__nv_bfloat16*)out_bf16; a.pre_act_bf16 = (__nv_bfloat16*)pre_act_bf16; a.ld_bf16 = ld_bf16;
    a.ln_gamma = ln_gamma; a.ln_beta = ln_beta; a.ln_out = (__nv_bfloat16*)ln_out; a.ln_stats = ln_stats;
    const int kch = K / 64;
    cudaStream_t st = (cudaStream_t)stream;
    // weight tile resident (kch x 16 KB) + ring of activation K-chunks; two CTAs per SM whenever they fit (<= ~110 KB each)
    const size_t fixed = 1024 + (size_t)kch * G_NT * 128 + 8 * 2048 + sizeof(GemmPBars);
    int pstages = fixed + 2 * (size_t)G_BM * 128 <= 110 * 1024 ? (int)((110 * 1024 - fixed) / ((size_t)G_BM * 128)) : 0;
    int per_sm = 2;
    if (pstages < 2) {
        pstages = fixed + 2 * (size_t)G_BM * 128 <= 220 * 1024 ? (int)((220 * 1024 - fixed) / ((size_t)G_BM * 128)) : 0;
        per_sm = 1;
    }
    if (pstages > GP_MAX_STAGES) pstages = GP_MAX_STAGES;
    ASME_REQUIRE(!ln || pstages >= 2, "tc_gemm_ln: K=%d does not fit the persistent kernel", K);
    if ((g_gemm_variant == 1 || ln) && pstages >= 2) {      // (a weight tile too large for shared memory falls back to the per-tile kernel)
        const int stages = pstages;
        const size_t smem_p = fixed + (size_t)stages * G_BM * 128;
        const int n_tiles = ceil_div(N, G_NT), m_tiles = ceil_div(M, G_BM);
        int gx = (ASME_NUM_SMS * per_sm) / n_tiles;
        if (gx < 1) gx = 1;
        if (gx > m_tiles) gx = m_tiles;
        const dim3 pgrid(gx, n_tiles);
#define LAUNCH_PERSIST(BMN, LNF)                                                                                              \
        {                                                                                                                     \
            { const int _rc = asme_ensure_max_smem((const void*)tc_gemm_persist_kernel<BMN, LNF>); if (_rc) return _rc; }     \
            tc_gemm_persist_kernel<BMN, LNF><<<pgrid, G_TALL_THREADS, smem_p, st>>>(tmA, tmB, a, stages);                     \
        }
        if (b_is_kn) { if (ln) LAUNCH_PERSIST(true, true) else LAUNCH_PERSIST(true, false) }
        else { if (ln) LAUNCH_PERSIST(false, true) else LAUNCH_PERSIST(false, false) }
#undef LAUNCH_PERSIST
        ASME_LAUNCH_OK();
        return ASME_OK;
    }
    const int stages = kch < G_STAGES ? kch : G_STAGES;
    const size_t smem = 1024 + (size_t)stages * ((size_t)G_BM * 128 + G_NT * 128) + sizeof(GemmBars);
    const dim3 grid(ceil_div(M, G_BM), ceil_div(N, G_NT));
    if (b_is_kn) {
        { const int _rc = asme_ensure_max_smem((const void*)tc_gemm_tall_kernel<true>); if (_rc) return _rc; }
        tc_gemm_tall_kernel<true><<<grid, G_TALL_THREADS, smem, st>>>(tmA, tmB, a, stages);
    } else {
        { const int _rc = asme_ensure_max_smem((const void*)tc_gemm_tall_kernel<false>); if (_rc) return _rc; }
        tc_gemm_tall_kernel<false><<<grid, G_TALL_THREADS, smem, st>>>(tmA, tmB, a, stages);
    }
    ASME_LAUNCH_OK();
    return ASME_OK;
}

extern "C" int asme_b200_tc_gemm(const void* A, const void* B, int M, int N, int K, int b_is_kn, const float* bias, int act,
                                 const void* gelu_grad_of, float p_drop, unsigned long long seed, unsigned int site,
                                 unsigned int post_site, const float* residual, float* out_f32, void* out_bf16, int ld_bf16,
                                 void* pre_act_bf16, asme_stream_t stream) {
    return tc_gemm_impl(A, B, M, N, K, b_is_kn, bias, act, gelu_grad_of, p_drop, seed, site, post_site, residual, out_f32, out_bf16,
                        ld_bf16, pre_act_bf16, nullptr, nullptr, nullptr, nullptr, stream);
}

// the same with a fused LayerNorm of the fp32 output rows: ln_out (M,N) bf16 = LN(out_f32; gamma, beta), ln_stats (2,M) or NULL
extern "C" int asme_b200_tc_gemm_ln(const void* A, const void* B, int M, int N, int K, int b_is_kn, const float* bias, int act,
                                    const void* gelu_grad_of, float p_drop, unsigned long long seed, unsigned int site,
                                    unsigned int post_site, const float* residual, float* out_f32, void* out_bf16, int ld_bf16,
                                    void* pre_act_bf16, const float* ln_gamma, const float* ln_beta, void* ln_out, float* ln_stats,
                                    asme_stream_t stream) {
    ASME_REQUIRE(ln_gamma && ln_beta && ln_out, "tc_gemm_ln: null LayerNorm argument");
    return tc_gemm_impl(A, B, M, N, K, b_is_kn, bias, act, gelu_grad_of, p_drop, seed, site, post_site, residual, out_f32, out_bf16,
                        ld_bf16, pre_act_bf16, ln_gamma, ln_beta, ln_out, ln_stats, stream);
}

// ------------------------------------------------------------------------------------------------------------
// weight gradient: dW[N,K] (+)= dY[M,N]^T X[M,K], contraction over the M tokens, split over token ranges
// ------------------------------------------------------------------------------------------------------------
#define W_SLAB 64              // tokens per pipeline stage
#define W_STAGES 4
struct TcWgradArgs {
    int M, N, K;               // M tokens; output is (N, K)
    int slabs_per_split, n_slabs;
    float* partial;            // [splits][N][K]
    float* partial_bias;       // [splits][N] or NULL: column sums of dY, computed by one extra N=16 MMA against a ones column
};
struct __align__(8) WgradBars {
    uint64_t full[W_STAGES];
    uint64_t empty[W_STAGES];
    uint64_t done;
    uint32_t tmem_base;
};

// CTA (split, n_tile, k_tile): output rows n_tile*128..+127 (dY columns), output columns k_tile*256..+255 (X columns),
// tokens of `split`.
__global__ void __launch_bounds__(G_THREADS, 1) tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmY,
                                                                const __grid_constant__ CUtensorMap tmX, const TcWgradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u)   /* pointer arithmetic on the __shared__ array keeps the address space: LDS/STS, not generic LD/ST */;
    const int nrow0 = blockIdx.y * 128;                  // first output row  (dY column) of this CTA
    const int kcol0 = blockIdx.z * 256;                  // first output column (X column) of this CTA
    const int KT = min(256, a.K - kcol0);                // output columns of this CTA (multiple of 64)
    const int nb_x = KT / 64;
    const size_t blk = (size_t)W_SLAB * 128;             // one 64-column block of one slab: 64 tokens x 128 B
    const size_t stage_bytes = (size_t)(2 + 4) * blk;    // 2 dY blocks + up to 4 X blocks (fixed slot size)
    uint8_t* sOnes = smem + W_STAGES * stage_bytes;      // [64 tokens][128 B]: bf16 1.0 in column 0 (MN-major B of the bias MMA)
    WgradBars* bars = reinterpret_cast<WgradBars*>(sOnes + blk);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int split = blockIdx.x;
    const int s0 = split * a.slabs_per_split;
    const int s1 = min(a.n_slabs, s0 + a.slabs_per_split);
    const bool do_bias = a.partial_bias != nullptr && blockIdx.z == 0;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < KT + (do_bias ? 16 : 0)) tmem_cols <<= 1;
    if (do_bias) {
        for (int i = threadIdx.x; i < (int)(blk / 16); i += blockDim.x) reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        if (threadIdx.x < W_SLAB) *reinterpret_cast<uint16_t*>(sOnes + sw128_offset(threadIdx.x, 0)) = 0x3F80;   // bf16(1.0)
        fence_proxy_async_smem();
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmY);
        tma_prefetch_desc(&tmX);
        for (int s = 0; s < W_STAGES; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        mbar_init(&bars->done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&bars->tmem_base, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            int i = 0;
            for (int sl = s0; sl < s1; ++sl, ++i) {
                const int s = i % W_STAGES;
                const uint32_t ph = (uint32_t)(i / W_STAGES) & 1u;
                mbar_wait_lean(&bars->empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&bars->full[s], (uint32_t)((2 + nb_x) * blk));
                uint8_t* base = smem + (size_t)s * stage_bytes;
                // dY columns beyond N are out of bounds of the tensor map: zero-filled, the extra output rows are not stored
                for (int j = 0; j < 2; ++j) tma_load_2d(base + (size_t)j * blk, &tmY, &bars->full[s], nrow0 + j * 64, sl * W_SLAB);
                for (int j = 0; j < nb_x; ++j) tma_load_2d(base + (size_t)(2 + j) * blk, &tmX, &bars->full[s], kcol0 + j * 64, sl * W_SLAB);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // A = dY^T (M-dim = dY columns, MN-major), B = X^T viewed as [N = X columns][K = tokens] (MN-major)
            const uint32_t idesc = idesc_major(128, KT, 1, 1);
            const uint32_t idesc_b = idesc_major(128, 16, 1, 1);
            int i = 0;
            for (int sl = s0; sl < s1; ++sl, ++i) {
                const int s = i % W_STAGES;
                const uint32_t ph = (uint32_t)(i / W_STAGES) & 1u;
                mbar_wait_lean(&bars->full[s], ph);
                tc_fence_after();
                uint8_t* base = smem + (size_t)s * stage_bytes;
#pragma unroll
                for (int ks = 0; ks < W_SLAB / UMMA_K; ++ks) {
                    const uint64_t ad = smem_desc_mn_sw128(smem_u32(base + (size_t)ks * 2048), (uint32_t)blk);
                    const uint64_t bd = smem_desc_mn_sw128(smem_u32(base + 2 * blk + (size_t)ks * 2048), (uint32_t)blk);
                    umma_bf16(tmem_base, ad, bd, idesc, (uint32_t)((i | ks) != 0));
                    if (do_bias)
                        umma_bf16(tmem_base + (uint32_t)KT, ad, smem_desc_mn_sw128(smem_u32(sOnes + (size_t)ks * 2048), 0), idesc_b,
                                  (uint32_t)((i | ks) != 0));
                }
                umma_commit(&bars->empty[s]);
            }
            umma_commit(&bars->done);
        }
    } else if (warp >= 4) {
        const int q = warp % 4;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        mbar_wait_lean(&bars->done, 0);
        tc_fence_after();
        float* out = a.partial + (size_t)split * a.N * a.K;
        const int n = nrow0 + q * 32 + lane;               // output row (= dY column)
        for (int k0 = 0; k0 < KT; k0 += 32) {
            float v[32];
            tmem_ld32(lane_addr + (uint32_t)k0, v);
            tmem_ld_wait();
            if (n < a.N) {
#pragma unroll
                for (int c = 0; c < 32; c += 4)
                    *reinterpret_cast<float4*>(out + (size_t)n * a.K + kcol0 + k0 + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
            }
        }
        if (do_bias) {
            float v[16];
            tmem_ld16(lane_addr + (uint32_t)KT, v);
            tmem_ld_wait();
            if (n < a.N) a.partial_bias[(size_t)split * a.N + n] = v[0];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// out[i] (+)= sum over splits of partial[split][i]: 32 float4 columns x 8 split groups per block, fixed summation order.
// Two reductions in one launch: blocks [0, blocks0) reduce (partial, n, out), the remaining ones (partial2, n2, out2) -- the
// weight gradient and its bias gradient.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int splits, long long n,
                                                           float* __restrict__ out, int blocks0, const float* __restrict__ partial2,
                                                           long long n2, float* __restrict__ out2, int accumulate) {
    __shared__ float4 red[8][32];
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
    int blk = blockIdx.x;
    if (blk >= blocks0) {          // block-uniform
        blk -= blocks0;
        partial = partial2; n = n2; out = out2;
    }
    const long long i = ((long long)blk * 32 + tx) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) {
        float4 s1 = s;
        int p = ty;
        for (; p + 8 < splits; p += 16) {
            add4(s, ldg4(partial + (size_t)p * n + i));
            add4(s1, ldg4(partial + (size_t)(p + 8) * n + i));
        }
        if (p < splits) add4(s, ldg4(partial + (size_t)p * n + i));
        add4(s, s1);
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && i < n) {
#pragma unroll
        for (int g = 1; g < 8; ++g) add4(s, red[g][tx]);
        float4* o = reinterpret_cast<float4*>(out + i);
        if (accumulate) {
            const float4 c = *o;
            s.x += c.x; s.y += c.y; s.z += c.z; s.w += c.w;
        }
        *o = s;
    }
}

static int wgrad_splits(int M) {
    const int n_slabs = ceil_div(M, W_SLAB);
    const int s = n_slabs < ASME_NUM_SMS ? n_slabs : ASME_NUM_SMS;
    return ceil_div(n_slabs, ceil_div(n_slabs, s));      // no empty splits
}
extern "C" size_t asme_b200_tc_wgrad_workspace_bytes(int M, int N, int K) {
    return ((size_t)wgrad_splits(M < 1 ? 1 : M) * N * K + (size_t)wgrad_splits(M < 1 ? 1 : M) * N) * sizeof(float);
}

extern "C" int asme_b200_tc_wgrad(const void* dY, const void* X, int M, int N, int K, float* dW, float* dbias, int accumulate,
                                  void* ws, size_t ws_bytes, asme_stream_t stream) {
    ASME_REQUIRE(dY && X && dW, "tc_wgrad: null argument");
    ASME_REQUIRE(N >= 64 && N % 64 == 0, "tc_wgrad: N=%d unsupported (multiple of 64)", N);
    ASME_REQUIRE(K >= 64 && K % 64 == 0, "tc_wgrad: K=%d unsupported (multiple of 64)", K);
    if (M == 0) return ASME_OK;
    ASME_REQUIRE(ws_bytes >= asme_b200_tc_wgrad_workspace_bytes(M, N, K), "tc_wgrad: workspace too small");
    CUtensorMap tmY, tmX;
    int rc = asme_tc_make_tmap_bf16(&tmY, dY, M, N, N, W_SLAB);
    if (rc) return rc;
    rc = asme_tc_make_tmap_bf16(&tmX, X, M, K, K, W_SLAB);
    if (rc) return rc;
    TcWgradArgs a{};
    a.M = M; a.N = N; a.K = K;
    a.n_slabs = ceil_div(M, W_SLAB);
    const int splits = wgrad_splits(M);
    a.slabs_per_split = ceil_div(a.n_slabs, splits);
    a.partial = (float*)ws;
    a.partial_bias = dbias ? a.partial + (size_t)splits * N * K : nullptr;
    const size_t stage_bytes = (size_t)(2 + 4) * W_SLAB * 128;
    const size_t smem = 1024 + W_STAGES * stage_bytes + (size_t)W_SLAB * 128 + sizeof(WgradBars);
    cudaStream_t st = (cudaStream_t)stream;
    { const int _rc = asme_ensure_max_smem((const void*)tc_wgrad_kernel); if (_rc) return _rc; }
    tc_wgrad_kernel<<<dim3(splits, ceil_div(N, 128), ceil_div(K, 256)), G_THREADS, smem, st>>>(tmY, tmX, a);
    ASME_LAUNCH_OK();
    const long long n = (long long)N * K;
    const int blocks0 = (int)ceil_div(n / 4, 32), blocks1 = dbias ? ceil_div(N / 4, 32) : 0;
    wgrad_reduce_kernel<<<blocks0 + blocks1, 256, 0, st>>>(a.partial, splits, n, dW, blocks0, a.partial_bias, (long long)N, dbias, accumulate);
    ASME_LAUNCH_OK();
    return ASME_OK;
}
