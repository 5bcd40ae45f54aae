// Masked self-attention for short sequences (S <= 256, head dim 16 / 32 / 64) on the tensor cores:
// tcgen05.mma with fp32 accumulators in tensor memory, Q / K / V tiles staged by TMA straight out of the (T, 3H) bf16
// output of the QKV projection, probabilities handed to the second MMA through 128-byte-swizzled shared memory.
//
// replaces: mask materialisation (models/transformer/sequence_representation.py:33-48) and Attention.forward
//           (models/common/layers/transformer_layers.py:145-155): QK^T/sqrt(d) -> masked_fill(mask==0,-1e9) -> softmax
//           -> dropout -> .V.  The (B,1,S,S) mask never exists: validity comes from the (B,S) key-padding bits and the
//           causal flag; a fully masked row attends uniformly to all S keys (quirk Q4, -1e9 not -inf).
//
// One CTA = one sequence b x one 64-column slice of the hidden size (64/d heads).  Shared memory holds that slice of Q, K
// and V for the whole sequence (3 x 32 KB) plus one probability tile (128 queries x 256 keys bf16, 64 KB).
//   warp 0   TMA producer (three boxes)           warp 1   MMA issuer           warp 2   TMEM allocator
//   warps 4-7  softmax / epilogue: thread = one query row (TMEM lane)
// Per (head, 128-query tile):  S = Q K^T (K-major operands, the head's 32-byte K-steps inside the 128-byte rows)
//   -> softmax in registers (two passes over TMEM: max, then exp / sum / dropout), P -> shared memory (bf16, K-major)
//   -> O = P V (V consumed as an MN-major operand: no transposed copy) -> O / rowsum -> ctx (bf16).
// Row max / row sum-exp use the conventions of the fp32 SIMT kernels (attention_simt.cu),
// so forward and backward kernels of either flavour can be mixed and compared.
#include "common.cuh"
#include "tc_common.cuh"

using namespace tc;

#define AT_THREADS 256
#define AT_MAXS 256
#define MASK_FILL (-1e9f)
#define LOG2E 1.4426950408889634f

__host__ __device__ constexpr uint32_t at_idesc(int M, int N, int a_mn, int b_mn) {
    return idesc_bf16_f32(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}
__device__ __forceinline__ uint64_t at_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t at_pack(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

// dropout threshold of the tensor-core attention forward as the bit-sliced comparison wants it (see drop_keep_bits32)
struct DropPlan {
    uint32_t t[16];              // t[i] = bit i of the 16-bit threshold, replicated to 32 bits
    int low;                     // lowest set bit of the threshold (planes below it cannot change the outcome)
};
static DropPlan make_drop_plan(float p_drop) {
    DropPlan d{};
    uint32_t thr16 = (uint32_t)(p_drop * 65536.0f + 0.5f);
    if (thr16 > 65535u) thr16 = 65535u;
    d.low = 16;
    for (int i = 15; i >= 0; --i) {
        d.t[i] = ((thr16 >> i) & 1u) ? 0xffffffffu : 0u;
        if (d.t[i]) d.low = i;
    }
    return d;
}
struct AttnTcArgs {
    const uint8_t* key_valid;    // (B,S) or NULL
    int B, S, heads, d, H, causal;
    float scale, p_drop, inv_keep;
    DropPlan drop;
    unsigned long long seed;
    unsigned int site;
    __nv_bfloat16* ctx;          // (B*S, H)
    float* stats;                // (2, B*heads*S) or NULL
    uint32_t* keep_bits;         // (B*heads*S, 8) or NULL: bit j of word w of row (b,h,q) = key 32w+j survived the dropout
    const int64_t* only_row;     // (B) or NULL: evaluation of selected positions -- only the 128-query tile that holds flat row
                                 // only_row[b] (= b*S + position) is computed, the other rows of ctx are left untouched
};

struct __align__(8) AttnBars {
    uint64_t loaded;
    uint64_t s_full;
    uint64_t p_full;
    uint64_t o_full;
    uint32_t tmem_base;
    uint32_t kvb[8];             // key-validity bits of the sequence (bit j of word w: key 32w+j may be attended)
};

// ---- dropout stream of the tensor-core attention ----------------------------------------------------------------
// The backward pass never regenerates the stream -- the forward stores the keep bits (1 bit per (query, key)) -- so the only
// contract is "Bernoulli(1 - thr16 / 65536), independent per element" (thr16 = p * 65536 rounded).  The bits of a 32-key chunk are
// drawn BIT-SLICED: plane i is one 32-bit hash word whose bit c is bit i of key c's 16-bit uniform number u_c, and keep_c =
// (u_c >= thr16) falls out of one logic operation per plane on all 32 keys at once -- the keep word needs no per-key extraction,
// compare and re-assembly (those were 8 of the 13.5 instructions per element of the dropout softmax, nearly all on the half-rate
// integer pipe; now 3.5 -- the kernel time did not move, the forward is not issue-bound: DESIGN.md section 6).  Planes below the lowest
// set bit of thr16 cannot change the outcome and are not drawn (p = 0.5: one plane).
// Plane words: (row key + counter * golden ratio) through multiply - xorshift - multiply - xorshift (the row key is a full lowbias32
// mix of (seed, site, row); one round less leaves consecutive planes correlated); tests/test_gpu_tc_attention.py checks keep rates
// and adjacent-key / adjacent-row / adjacent-head independence.
#define DROP_K1 0x21f0aaadu
#define DROP_K2 0x735a2d97u
__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0x735a2d97u; x ^= x >> 15;
    return x;
}
__device__ __forceinline__ uint32_t drop_rowkey(uint64_t seed, uint32_t site, uint64_t row) {
    return (lowbias32((uint32_t)row ^ (uint32_t)seed) ^ lowbias32((uint32_t)(row >> 32) ^ (uint32_t)(seed >> 32) ^ (site * 0x9E3779B9u))) * DROP_K1;
}
// keep bits of the 32 keys of chunk `ch` of one row (bit c = key ch*32+c survives).  Plane word n of chunk ch mixes x = rowkey +
// (16 ch + n) * golden: multiply, xorshift, multiply, xorshift (the first multiply is distributed: rowkey * K1 once per row).
// u < thr is the borrow of u - thr rippling up from the lowest set threshold bit: borrow' = t ? (borrow | ~u) : (borrow & ~u), ONE
// three-input logic operation per plane with the threshold mask t (all ones / all zeros) read from the kernel parameters.
#define DROP_PLANE(i)                                                                   \
    {                                                                                   \
        uint32_t r = base + (uint32_t)(i) * (0x9E3779B9u * DROP_K1);                    \
        r ^= r >> 15; r *= DROP_K2; r ^= r >> 16;                                       \
        borrow = (borrow & ~r) | (dp.t[i] & (borrow | ~r));                             \
    }
__device__ __forceinline__ uint32_t drop_keep_bits32(uint32_t rowkey_k1, int ch, const DropPlan& dp) {
    uint32_t borrow = 0u;
    const uint32_t base = rowkey_k1 + (uint32_t)(ch * 16) * (0x9E3779B9u * DROP_K1);
    switch (dp.low) {            // uniform; falls through the planes from the lowest set threshold bit upwards
        case 0: DROP_PLANE(0)
        case 1: DROP_PLANE(1)
        case 2: DROP_PLANE(2)
        case 3: DROP_PLANE(3)
        case 4: DROP_PLANE(4)
        case 5: DROP_PLANE(5)
        case 6: DROP_PLANE(6)
        case 7: DROP_PLANE(7)
        case 8: DROP_PLANE(8)
        case 9: DROP_PLANE(9)
        case 10: DROP_PLANE(10)
        case 11: DROP_PLANE(11)
        case 12: DROP_PLANE(12)
        case 13: DROP_PLANE(13)
        case 14: DROP_PLANE(14)
        case 15: DROP_PLANE(15)
        default: break;
    }
    return ~borrow;              // keep = not (u < thr)
}
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

#define ATF_THREADS 384        // warp 0 TMA, 1 MMA, 2 TMEM alloc, warps 4-7 and 8-11: two softmax warpgroups
struct __align__(8) AttnFwdShared {
    uint64_t loaded, s_full, p_full, o_full;
    uint32_t tmem_base;
    uint32_t kvb[8];             // key-validity bits of the sequence (bit j of word w: key 32w+j may be attended)
    float xm[2][128];            // per-warpgroup partial row max   (units of log2: score * scale * log2e)
    float xl[2][128];            // per-warpgroup partial row sum
};

// Both warpgroups own all 128 query rows of the tile (TMEM lane = row) and split the key columns in halves; partial row
// maxima / sums are exchanged through shared memory.  Two warps per SM sub-partition hide each other's TMEM-load, MUFU and
// shared-memory latencies (the one-warpgroup version issued on 25 % of the cycles).
__global__ void __launch_bounds__(ATF_THREADS, 1) attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u)   /* pointer arithmetic on the __shared__ array keeps the address space: LDS/STS, not generic LD/ST */;
    uint8_t* sQ = smem;                       // [256 tokens][128 B]   (64 hidden columns of this slice)
    uint8_t* sK = sQ + 32768;
    uint8_t* sV = sK + 32768;
    uint8_t* sP = sV + 32768;                 // [4 key chunks][128 queries][128 B]
    AttnFwdShared* sh = reinterpret_cast<AttnFwdShared*>(sP + 65536);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int b = blockIdx.x, slice = blockIdx.y;
    const int S = a.S, d = a.d;
    const int hps = 64 / d;                   // heads per 64-column slice
    const int n_qt_all = (S + 127) / 128;
    const int qt_only = a.only_row ? (int)((a.only_row[b] - (long long)b * S) / 128) : -1;     // CTA-uniform
    const int n_qt = qt_only >= 0 ? 1 : n_qt_all;   // query tiles this CTA visits; unit u -> (head u / n_qt, tile qt_of(u))
    const int NS = (S + 15) / 16 * 16;        // score columns computed by the first MMA
    const int units = hps * n_qt;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(&sh->loaded, 1);
        mbar_init(&sh->s_full, 1);
        mbar_init(&sh->p_full, 8);                     // one arrival per epilogue warp
        mbar_init(&sh->o_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&sh->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sh->tmem_base;
    const uint32_t tmem_o = tmem_base + 256;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&sh->loaded, 3 * 32768);
            tma_load_2d(sQ, &tmQKV, &sh->loaded, slice * 64, b * S);
            tma_load_2d(sK, &tmQKV, &sh->loaded, a.H + slice * 64, b * S);
            tma_load_2d(sV, &tmQKV, &sh->loaded, 2 * a.H + slice * 64, b * S);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = at_idesc(128, NS, 0, 0);
            const uint32_t idesc_o = at_idesc(128, d, 0, 1);
            const int ks_qk = d / 16;                        // K-steps of Q K^T
            const int ks_pv = (S + 15) / 16;                 // K-steps of P V
            mbar_wait_lean(&sh->loaded, 0);
            tc_fence_after();
            auto issue_s = [&](int u) {
                const int hh = u / n_qt, qt = qt_only >= 0 ? qt_only : u % n_qt;
                const uint64_t qd = smem_desc_sw128(smem_u32(sQ + (size_t)qt * 128 * 128));
                const uint64_t kd = smem_desc_sw128(smem_u32(sK));
                for (int ks = 0; ks < ks_qk; ++ks)
                    umma_bf16(tmem_base, smem_desc_advance(qd, (hh * d + ks * 16) * 2), smem_desc_advance(kd, (hh * d + ks * 16) * 2),
                              idesc_s, (uint32_t)(ks != 0));
                umma_commit(&sh->s_full);
            };
            issue_s(0);
            for (int u = 0; u < units; ++u) {
                const int hh = u / n_qt;
                mbar_wait_lean(&sh->p_full, (uint32_t)u & 1u);
                tc_fence_after();
                for (int ks = 0; ks < ks_pv; ++ks) {
                    const uint64_t pd = smem_desc_advance(smem_desc_sw128(smem_u32(sP + (size_t)(ks / 4) * 16384)), (ks % 4) * 32);
                    const uint64_t vd = at_desc_mn(smem_u32(sV + (size_t)ks * 2048 + (size_t)hh * d * 2), 0);
                    umma_bf16(tmem_o, pd, vd, idesc_o, (uint32_t)(ks != 0));
                }
                umma_commit(&sh->o_full);
                if (u + 1 < units) issue_s(u + 1);           // S of the next unit overlaps this unit's epilogue
            }
        }
    } else if (warp >= 4) {
        const int wg = (warp - 4) / 4;                       // which half of the key columns
        const int q4 = warp % 4;
        const int r = q4 * 32 + lane;                        // row of the query tile == TMEM lane
        const uint32_t lane_addr = ((uint32_t)(q4 * 32) << 16);
        if (warp == 4) {
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const int key = w * 32 + lane;
                bool ok = key < S;
                if (ok && a.key_valid) ok = a.key_valid[(size_t)b * S + key] != 0;
                const uint32_t bits = __ballot_sync(0xffffffffu, ok);
                if (lane == 0) sh->kvb[w] = bits;
            }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");       // the eight softmax warps only
        const int n_chunks = (NS + 31) / 32;
        const int ch_lo = wg == 0 ? 0 : (n_chunks + 1) / 2;
        const int ch_hi = wg == 0 ? (n_chunks + 1) / 2 : n_chunks;
        const float cs = a.scale * LOG2E;                    // exp(x) = exp2(x * log2e); the row max itself stays in natural units so
                                                             // that a fully masked row's max is EXACTLY -1e9 (no cancellation error)
        for (int u = 0; u < units; ++u) {
            const int hh = u / n_qt, qt = qt_only >= 0 ? qt_only : u % n_qt;
            const int head = slice * hps + hh;
            const int qi = qt * 128 + r;                     // query position in the sequence
            const bool q_ok = qi < S;
            const long long bh = (long long)b * a.heads + head;
            const int q_warp_min = qt * 128 + q4 * 32;       // smallest query index of this warp (causal fast-path test)
            mbar_wait_lean(&sh->s_full, (uint32_t)u & 1u);
            tc_fence_after();
            // ---- pass 1: partial row maximum over this warpgroup's key columns
            float m = -INFINITY;
#pragma unroll 1
            for (int ch = ch_lo; ch < ch_hi; ++ch) {
                const uint32_t bits = sh->kvb[ch];
                if (bits == 0u) {            // a chunk of padding keys only (warp-uniform): every score is the mask fill value
                    if (ch * 32 < S) m = fmaxf(m, MASK_FILL);
                    continue;
                }
                float v[32];
                tmem_ld32(tmem_base + lane_addr + (uint32_t)(ch * 32), v);
                tmem_ld_wait();
                const bool plain = bits == 0xffffffffu && (!a.causal || ch * 32 + 31 <= q_warp_min);   // warp-uniform
                if (plain) {
                    float mm = v[0];
#pragma unroll
                    for (int c = 1; c < 32; ++c) mm = fmaxf(mm, v[c]);
                    m = fmaxf(m, mm * a.scale);
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int key = ch * 32 + c;
                        const bool ok = ((bits >> c) & 1u) && (!a.causal || key <= qi);
                        float t = ok ? v[c] * a.scale : MASK_FILL;
                        t = key < S ? t : -INFINITY;
                        m = fmaxf(m, t);
                    }
                }
            }
            sh->xm[wg][r] = m;
            asm volatile("bar.sync 2, 256;" ::: "memory");
            m = fmaxf(m, sh->xm[wg ^ 1][r]);
            const float m2 = m * LOG2E;
            const float e_masked = fast_exp2((MASK_FILL - m) * LOG2E);     // 1 in a fully masked row (uniform attention), else 0
            // ---- pass 2: exp, partial row sum, dropout, P -> shared memory (bf16, 128B-swizzled K-major tile)
            float l = 0.f;
            uint32_t rowkey = 0;
            if (a.p_drop > 0.f) rowkey = drop_rowkey(asme_seed(a.seed), a.site, (uint64_t)bh * S + (uint64_t)(q_ok ? qi : 0));
#pragma unroll 1
            for (int ch = ch_lo; ch < ch_hi; ++ch) {
                const uint32_t bits = sh->kvb[ch];
                uint8_t* chunk_p = sP + (size_t)(ch / 2) * 16384;
                // a chunk of padding keys only, and no row of this warp is fully masked (e_masked == 0 everywhere): the
                // probabilities are exactly zero -- no exponentials, no dropout stream (the keep bits stay zero: they only ever
                // multiply these zeros), just the zero tile the second MMA reads
                if (bits == 0u && __all_sync(0xffffffffu, e_masked == 0.f)) {
#pragma unroll
                    for (int u16 = 0; u16 < 4; ++u16)
                        *reinterpret_cast<uint4*>(chunk_p + sw128_offset(r, (ch & 1) * 4 + u16)) = make_uint4(0u, 0u, 0u, 0u);
                    continue;
                }
                float v[32];
                tmem_ld32(tmem_base + lane_addr + (uint32_t)(ch * 32), v);
                tmem_ld_wait();
                const bool plain = bits == 0xffffffffu && (!a.causal || ch * 32 + 31 <= q_warp_min);
                if (plain) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        v[c] = fast_exp2(fmaf(v[c], cs, -m2));
                        l += v[c];
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int key = ch * 32 + c;
                        const bool ok = ((bits >> c) & 1u) && (!a.causal || key <= qi);
                        float e = ok ? fast_exp2(fmaf(v[c], cs, -m2)) : e_masked;
                        e = key < S ? e : 0.f;
                        l += e;
                        v[c] = e;
                    }
                }
                if (a.p_drop > 0.f) {
                    const uint32_t kb = drop_keep_bits32(rowkey, ch, a.drop);
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] = ((kb >> c) & 1u) ? v[c] * a.inv_keep : 0.f;
                    if (a.keep_bits && q_ok) a.keep_bits[((size_t)bh * S + qi) * 8 + ch] = kb;
                }
                uint8_t* chunk = sP + (size_t)(ch / 2) * 16384;     // 64 keys per chunk, this 32-key half = units (ch&1)*4 .. +3
#pragma unroll
                for (int u16 = 0; u16 < 4; ++u16) {
                    uint4 w;
                    w.x = at_pack(v[u16 * 8 + 0], v[u16 * 8 + 1]); w.y = at_pack(v[u16 * 8 + 2], v[u16 * 8 + 3]);
                    w.z = at_pack(v[u16 * 8 + 4], v[u16 * 8 + 5]); w.w = at_pack(v[u16 * 8 + 6], v[u16 * 8 + 7]);
                    *reinterpret_cast<uint4*>(chunk + sw128_offset(r, (ch & 1) * 4 + u16)) = w;
                }
            }
            sh->xl[wg][r] = l;
            fence_proxy_async_smem();                         // make the generic-proxy P writes visible to the tensor core
            tc_fence_before();
            mbar_arrive_warp(&sh->p_full);
            // ---- O = P V done: normalise and store (each warpgroup half of the head's columns)
            mbar_wait_lean(&sh->o_full, (uint32_t)u & 1u);
            tc_fence_after();
            asm volatile("bar.sync 2, 256;" ::: "memory");   // both partial sums are in shared memory
            l += sh->xl[wg ^ 1][r];
            const float inv_l = 1.0f / l;
            if (wg == 0 && a.stats && q_ok) {
                a.stats[bh * S + qi] = m;                                         // natural units, as the SIMT kernels store it
                a.stats[(long long)a.B * a.heads * S + bh * S + qi] = l;
            }
            const int dcols = d >= 32 ? d / 2 : d;            // columns per warpgroup (d = 16: warpgroup 0 stores everything)
            if (d >= 32 || wg == 0) {
                for (int c0 = wg * (d >= 32 ? dcols : 0); c0 < wg * (d >= 32 ? dcols : 0) + dcols; c0 += 16) {
                    float o[16];
                    tmem_ld16(tmem_o + lane_addr + (uint32_t)c0, o);
                    tmem_ld_wait();
                    if (q_ok) {
                        __nv_bfloat16* dst = a.ctx + ((size_t)b * S + qi) * a.H + slice * 64 + hh * d + c0;
                        uint4 w0, w1;
                        w0.x = at_pack(o[0] * inv_l, o[1] * inv_l); w0.y = at_pack(o[2] * inv_l, o[3] * inv_l);
                        w0.z = at_pack(o[4] * inv_l, o[5] * inv_l); w0.w = at_pack(o[6] * inv_l, o[7] * inv_l);
                        w1.x = at_pack(o[8] * inv_l, o[9] * inv_l); w1.y = at_pack(o[10] * inv_l, o[11] * inv_l);
                        w1.z = at_pack(o[12] * inv_l, o[13] * inv_l); w1.w = at_pack(o[14] * inv_l, o[15] * inv_l);
                        reinterpret_cast<uint4*>(dst)[0] = w0;
                        reinterpret_cast<uint4*>(dst)[1] = w1;
                    }
                }
            }
            tc_fence_before();      // O (and S) TMEM reads are ordered before the MMAs of the next unit via p_full / o_full
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Forward, second version (default): the probabilities never touch shared memory.  The softmax threads write P as packed bf16 pairs
// straight back into the tensor-memory columns the scores came from (tcgen05.st; key pair (2j, 2j+1) -> column j) and the second
// MMA takes its A operand from tensor memory (tcgen05.mma with [a_tmem]).  Without the 64 KB probability tile a CTA needs 96 KB of
// shared memory and 256 tensor-memory columns -- scores / P in [0, NS), O in [128, 128 + d) over score columns that are consumed by
// then -- so TWO CTAs share an SM and each other's TMA round trip, MMA and softmax phases overlap: the first version ran one CTA per
// SM through load -> MMA -> softmax -> MMA -> store strictly in turn (ncu: 36 % issue utilisation, 4 % tensor pipe, 1.73 waves at
// B = 256).  One softmax warpgroup per CTA: thread = query row over ALL key columns, so row maxima / sums need no exchange and a
// thread only ever overwrites score columns it has already read (P chunk c lands in columns [16c, 16c+16), inside score chunk c/2).
// ------------------------------------------------------------------------------------------------------------
#define ATF2_THREADS 256       // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, warps 4-7 softmax / epilogue
#define ATF2_OCOL 128          // first tensor-memory column of O
struct __align__(8) AttnFwd2Shared {
    uint64_t loaded, s_full, p_full, o_full, o_done;
    uint32_t tmem_base;
    uint32_t kvb[8];
};

__global__ void __launch_bounds__(ATF2_THREADS, 2) attn_tc_fwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem;                       // [256 tokens][128 B]   (64 hidden columns of this slice)
    uint8_t* sK = sQ + 32768;
    uint8_t* sV = sK + 32768;
    AttnFwd2Shared* sh = reinterpret_cast<AttnFwd2Shared*>(sV + 32768);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int b = blockIdx.x, slice = blockIdx.y;
    const int S = a.S, d = a.d;
    const int hps = 64 / d;                   // heads per 64-column slice
    const int n_qt_all = (S + 127) / 128;
    const int qt_only = a.only_row ? (int)((a.only_row[b] - (long long)b * S) / 128) : -1;     // CTA-uniform
    const int n_qt = qt_only >= 0 ? 1 : n_qt_all;
    const int NS = (S + 15) / 16 * 16;        // score columns computed by the first MMA
    const int units = hps * n_qt;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(&sh->loaded, 1);
        mbar_init(&sh->s_full, 1);
        mbar_init(&sh->p_full, 4);                     // one arrival per softmax warp
        mbar_init(&sh->o_full, 1);
        mbar_init(&sh->o_done, 4);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&sh->tmem_base, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sh->tmem_base;
    const uint32_t tmem_o = tmem_base + ATF2_OCOL;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&sh->loaded, 3 * 32768);
            tma_load_2d(sQ, &tmQKV, &sh->loaded, slice * 64, b * S);
            tma_load_2d(sK, &tmQKV, &sh->loaded, a.H + slice * 64, b * S);
            tma_load_2d(sV, &tmQKV, &sh->loaded, 2 * a.H + slice * 64, b * S);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = at_idesc(128, NS, 0, 0);
            const uint32_t idesc_o = at_idesc(128, d, 0, 1);
            const int ks_qk = d / 16;                        // K-steps of Q K^T
            const int ks_pv = (S + 15) / 16;                 // K-steps of P V
            mbar_wait_lean(&sh->loaded, 0);
            tc_fence_after();
            for (int u = 0; u < units; ++u) {
                const int hh = u / n_qt, qt = qt_only >= 0 ? qt_only : u % n_qt;
                if (u > 0) {                                 // the score columns double as P and O of the previous unit
                    mbar_wait_lean(&sh->o_done, (uint32_t)(u - 1) & 1u);
                    tc_fence_after();
                }
                const uint64_t qd = smem_desc_sw128(smem_u32(sQ + (size_t)qt * 128 * 128));
                const uint64_t kd = smem_desc_sw128(smem_u32(sK));
                for (int ks = 0; ks < ks_qk; ++ks)
                    umma_bf16(tmem_base, smem_desc_advance(qd, (hh * d + ks * 16) * 2), smem_desc_advance(kd, (hh * d + ks * 16) * 2),
                              idesc_s, (uint32_t)(ks != 0));
                umma_commit(&sh->s_full);
                mbar_wait_lean(&sh->p_full, (uint32_t)u & 1u);
                tc_fence_after();
                for (int ks = 0; ks < ks_pv; ++ks) {         // A = P from tensor memory: 16 keys = 8 packed columns per K-step
                    const uint64_t vd = at_desc_mn(smem_u32(sV + (size_t)ks * 2048 + (size_t)hh * d * 2), 0);
                    umma_bf16_ts(tmem_o, tmem_base + (uint32_t)(ks * 8), vd, idesc_o, (uint32_t)(ks != 0));
                }
                umma_commit(&sh->o_full);
            }
        }
    } else if (warp >= 4) {
        const int q4 = warp % 4;
        const int r = q4 * 32 + lane;                        // row of the query tile == TMEM lane
        const uint32_t lane_addr = ((uint32_t)(q4 * 32) << 16);
        if (warp == 4) {
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const int key = w * 32 + lane;
                bool ok = key < S;
                if (ok && a.key_valid) ok = a.key_valid[(size_t)b * S + key] != 0;
                const uint32_t bits = __ballot_sync(0xffffffffu, ok);
                if (lane == 0) sh->kvb[w] = bits;
            }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");       // the four softmax warps only
        const int n_chunks = (NS + 31) / 32;
        const float cs = a.scale * LOG2E;
        for (int u = 0; u < units; ++u) {
            const int hh = u / n_qt, qt = qt_only >= 0 ? qt_only : u % n_qt;
            const int head = slice * hps + hh;
            const int qi = qt * 128 + r;                     // query position in the sequence
            const bool q_ok = qi < S;
            const long long bh = (long long)b * a.heads + head;
            const int q_warp_min = qt * 128 + q4 * 32;       // smallest query index of this warp (causal fast-path test)
            mbar_wait_lean(&sh->s_full, (uint32_t)u & 1u);
            tc_fence_after();
            // ---- pass 1: row maximum
            float m = -INFINITY;
#pragma unroll 1
            for (int ch = 0; ch < n_chunks; ++ch) {
                const uint32_t bits = sh->kvb[ch];
                if (bits == 0u) {            // a chunk of padding keys only (warp-uniform): every score is the mask fill value
                    if (ch * 32 < S) m = fmaxf(m, MASK_FILL);
                    continue;
                }
                float v[32];
                if (ch * 32 + 32 <= NS) tmem_ld32(tmem_base + lane_addr + (uint32_t)(ch * 32), v);
                else {                       // NS = 16 (mod 32): the last chunk holds 16 score columns (the rest belongs to nobody)
                    float v16[16];
                    tmem_ld16(tmem_base + lane_addr + (uint32_t)(ch * 32), v16);
#pragma unroll
                    for (int c = 0; c < 16; ++c) { v[c] = v16[c]; v[16 + c] = 0.f; }
                }
                tmem_ld_wait();
                const bool plain = bits == 0xffffffffu && (!a.causal || ch * 32 + 31 <= q_warp_min);   // warp-uniform
                if (plain) {
                    float mm = v[0];
#pragma unroll
                    for (int c = 1; c < 32; ++c) mm = fmaxf(mm, v[c]);
                    m = fmaxf(m, mm * a.scale);
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int key = ch * 32 + c;
                        const bool ok = ((bits >> c) & 1u) && (!a.causal || key <= qi);
                        float t = ok ? v[c] * a.scale : MASK_FILL;
                        t = key < S ? t : -INFINITY;
                        m = fmaxf(m, t);
                    }
                }
            }
            const float m2 = m * LOG2E;
            const float e_masked = fast_exp2((MASK_FILL - m) * LOG2E);     // 1 in a fully masked row (uniform attention), else 0
            // ---- pass 2: exp, row sum, dropout, P -> tensor memory (packed bf16 pairs over the score columns already read)
            float l = 0.f;
            uint32_t rowkey = 0;
            if (a.p_drop > 0.f) rowkey = drop_rowkey(asme_seed(a.seed), a.site, (uint64_t)bh * S + (uint64_t)(q_ok ? qi : 0));
#pragma unroll 1
            for (int ch = 0; ch < n_chunks; ++ch) {
                const uint32_t bits = sh->kvb[ch];
                const uint32_t p_addr = tmem_base + lane_addr + (uint32_t)(ch * 16);
                uint32_t w[16];
                if (bits == 0u && __all_sync(0xffffffffu, e_masked == 0.f)) {      // exact zeros: no exponentials, no dropout stream
#pragma unroll
                    for (int c = 0; c < 16; ++c) w[c] = 0u;
                    tmem_st16(p_addr, w);
                    continue;
                }
                float v[32];
                if (ch * 32 + 32 <= NS) tmem_ld32(tmem_base + lane_addr + (uint32_t)(ch * 32), v);
                else {
                    float v16[16];
                    tmem_ld16(tmem_base + lane_addr + (uint32_t)(ch * 32), v16);
#pragma unroll
                    for (int c = 0; c < 16; ++c) { v[c] = v16[c]; v[16 + c] = 0.f; }
                }
                tmem_ld_wait();
                const bool plain = bits == 0xffffffffu && (!a.causal || ch * 32 + 31 <= q_warp_min);
                if (plain) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        v[c] = fast_exp2(fmaf(v[c], cs, -m2));
                        l += v[c];
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int key = ch * 32 + c;
                        const bool ok = ((bits >> c) & 1u) && (!a.causal || key <= qi);
                        float e = ok ? fast_exp2(fmaf(v[c], cs, -m2)) : e_masked;
                        e = key < S ? e : 0.f;
                        l += e;
                        v[c] = e;
                    }
                }
                if (a.p_drop > 0.f) {
                    const uint32_t kb = drop_keep_bits32(rowkey, ch, a.drop);
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] = ((kb >> c) & 1u) ? v[c] * a.inv_keep : 0.f;
                    if (a.keep_bits && q_ok) a.keep_bits[((size_t)bh * S + qi) * 8 + ch] = kb;
                }
#pragma unroll
                for (int c = 0; c < 16; ++c) w[c] = at_pack(v[2 * c], v[2 * c + 1]);
                if (ch * 32 + 32 <= NS) tmem_st16(p_addr, w);
                else {                       // 16 keys -> 8 packed columns
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(p_addr), "r"(w[0]),
                                 "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                                 : "memory");
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive_warp(&sh->p_full);
            // ---- O = P V done: normalise and store
            mbar_wait_lean(&sh->o_full, (uint32_t)u & 1u);
            tc_fence_after();
            const float inv_l = 1.0f / l;
            if (a.stats && q_ok) {
                a.stats[bh * S + qi] = m;                                         // natural units, as the SIMT kernels store it
                a.stats[(long long)a.B * a.heads * S + bh * S + qi] = l;
            }
            for (int c0 = 0; c0 < d; c0 += 16) {
                float o[16];
                tmem_ld16(tmem_o + lane_addr + (uint32_t)c0, o);
                tmem_ld_wait();
                if (q_ok) {
                    __nv_bfloat16* dst = a.ctx + ((size_t)b * S + qi) * a.H + slice * 64 + hh * d + c0;
                    uint4 w0, w1;
                    w0.x = at_pack(o[0] * inv_l, o[1] * inv_l); w0.y = at_pack(o[2] * inv_l, o[3] * inv_l);
                    w0.z = at_pack(o[4] * inv_l, o[5] * inv_l); w0.w = at_pack(o[6] * inv_l, o[7] * inv_l);
                    w1.x = at_pack(o[8] * inv_l, o[9] * inv_l); w1.y = at_pack(o[10] * inv_l, o[11] * inv_l);
                    w1.z = at_pack(o[12] * inv_l, o[13] * inv_l); w1.w = at_pack(o[14] * inv_l, o[15] * inv_l);
                    reinterpret_cast<uint4*>(dst)[0] = w0;
                    reinterpret_cast<uint4*>(dst)[1] = w1;
                }
            }
            tc_fence_before();
            mbar_arrive_warp(&sh->o_done);       // scores of the next unit may overwrite P and O
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

static int g_attn_fwd_variant = 2;      // 2: probabilities in tensor memory, two CTAs per SM (default); 1: the first kernel (A/B)
static int tc_attn_fwd_impl(const void* qkv, const uint8_t* key_valid, int B, int S, int heads, int d, int causal,
                            float p_drop, unsigned long long seed, unsigned int site, void* ctx, float* stats,
                            uint32_t* keep_bits, const int64_t* only_row, asme_stream_t stream);
extern "C" int asme_b200_tc_attn_fwd(const void* qkv, const uint8_t* key_valid, int B, int S, int heads, int d, int causal,
                                     float p_drop, unsigned long long seed, unsigned int site, void* ctx, float* stats,
                                     uint32_t* keep_bits, asme_stream_t stream) {
    return tc_attn_fwd_impl(qkv, key_valid, B, S, heads, d, causal, p_drop, seed, site, ctx, stats, keep_bits, nullptr, stream);
}
// evaluation of selected positions: only the query tile that holds flat row only_row[b] of every sequence b is computed
extern "C" int asme_b200_tc_attn_fwd_rows(const void* qkv, const uint8_t* key_valid, int B, int S, int heads, int d, int causal,
                                          const int64_t* only_row, void* ctx, asme_stream_t stream) {
    ASME_REQUIRE(only_row, "tc_attn_fwd_rows: null row list");
    return tc_attn_fwd_impl(qkv, key_valid, B, S, heads, d, causal, 0.f, 0ull, 0u, ctx, nullptr, nullptr, only_row, stream);
}
static int tc_attn_fwd_impl(const void* qkv, const uint8_t* key_valid, int B, int S, int heads, int d, int causal,
                            float p_drop, unsigned long long seed, unsigned int site, void* ctx, float* stats,
                            uint32_t* keep_bits, const int64_t* only_row, asme_stream_t stream) {
    ASME_REQUIRE(qkv && ctx, "tc_attn_fwd: null argument");
    ASME_REQUIRE(S >= 1 && S <= AT_MAXS, "tc_attn_fwd: S=%d unsupported (1..%d)", S, AT_MAXS);
    ASME_REQUIRE(d == 16 || d == 32 || d == 64, "tc_attn_fwd: head dim %d unsupported (16, 32, 64)", d);
    const int H = heads * d;
    ASME_REQUIRE(H % 64 == 0, "tc_attn_fwd: hidden size %d must be a multiple of 64", H);
    ASME_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "tc_attn_fwd: p_drop=%f", p_drop);
    if (B == 0) return ASME_OK;
    CUtensorMap tm;
    int rc = asme_tc_make_tmap_bf16(&tm, qkv, (long long)B * S, 3 * H, 3 * H, 256);
    if (rc) return rc;
    AttnTcArgs a{};
    a.key_valid = key_valid; a.B = B; a.S = S; a.heads = heads; a.d = d; a.H = H; a.causal = causal;
    a.scale = 1.0f / sqrtf((float)d); a.p_drop = p_drop; a.inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    a.seed = seed; a.site = site; a.ctx = (__nv_bfloat16*)ctx; a.stats = stats; a.keep_bits = keep_bits; a.only_row = only_row;
    a.drop = make_drop_plan(p_drop);
    // second kernel whenever the grid fills the machine: with fewer CTAs than SMs nothing shares an SM and the first kernel's two
    // softmax warpgroups per CTA are the shorter critical path (64 x 256 x 64: 0.077 vs 0.091 ms; 1024 x 200 x 128: 0.165 vs 0.112 ms)
    if (g_attn_fwd_variant == 3 || (g_attn_fwd_variant == 2 && (long long)B * (H / 64) >= ASME_NUM_SMS)) {
        const size_t smem2 = 1024 + 3 * 32768 + sizeof(AttnFwd2Shared);
        { const int _rc = asme_ensure_max_smem((const void*)attn_tc_fwd2_kernel); if (_rc) return _rc; }
        attn_tc_fwd2_kernel<<<dim3(B, H / 64), ATF2_THREADS, smem2, (cudaStream_t)stream>>>(tm, a);
        ASME_LAUNCH_OK();
        return ASME_OK;
    }
    const size_t smem = 1024 + 3 * 32768 + 65536 + sizeof(AttnFwdShared);
    { const int _rc = asme_ensure_max_smem((const void*)attn_tc_fwd_kernel); if (_rc) return _rc; }
    attn_tc_fwd_kernel<<<dim3(B, H / 64), ATF_THREADS, smem, (cudaStream_t)stream>>>(tm, a);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------
// One CTA = one sequence b x one 64-column slice (64/d heads); shared memory holds that slice of Q, K, V and dO for the
// whole sequence (4 x 32 KB), two 128 x 128 bf16 operand tiles (2 x 32 KB), the per-query constants and the dropout bits.
// Per head, two sweeps of 128 x 128 sub-tiles (MMAs: T1 = scores, T2 = dP, both recomputed in tensor memory):
//   sweep A, rows = queries:  dS -> tile X;                 dQ[rows]  += X  K[cols]    (K consumed MN-major)
//   sweep B, rows = keys   :  P^T.drop -> Y1, dS^T -> Y2;   dV[rows]  += Y1 dO[cols],  dK[rows] += Y2 Q[cols]
// The transposed quantities come from transposed MMAs (K Q^T, V dO^T), never from a data transpose.  Gradients through
// masked_fill positions are zero (the reference replaces the value), also in fully masked rows.
struct AttnBwdArgs {
    const uint8_t* key_valid;
    int B, S, heads, d, H, causal;
    float scale, p_drop, inv_keep;
    const __nv_bfloat16* ctx;      // (B*S, H) forward output
    const __nv_bfloat16* d_ctx;    // (B*S, H) (also read through its tensor map)
    const float* stats;            // (2, B*heads*S)
    const uint32_t* keep_bits;     // (B*heads*S, 8) or NULL when p_drop == 0
    __nv_bfloat16* d_qkv;          // (B*S, 3H)
};
struct __align__(16) AttnBwdShared {
    float4 qc[2 * AT_MAXS];        // per query of the current head (single-sweep kernel: of the current PAIR of heads, head-major):
                                   // {rowmax * log2e, 1 / rowsum, D = sum_d dO.O, exp(-1e9 - rowmax)}
    uint32_t keep[2 * AT_MAXS * 8];// dropout keep bits, same indexing
    uint64_t loaded, t_full, x_full, acc_done;
    uint32_t tmem_base;
    uint32_t kvb[8];
};

__global__ void __launch_bounds__(ATF_THREADS, 1) attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                                    const __grid_constant__ CUtensorMap tmDO, const AttnBwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u)   /* pointer arithmetic on the __shared__ array keeps the address space: LDS/STS, not generic LD/ST */;
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + 32768;
    uint8_t* sV = sK + 32768;
    uint8_t* sDO = sV + 32768;
    uint8_t* sX = sDO + 32768;                // [2 chunks of 64 columns][128 rows][128 B]
    uint8_t* sY = sX + 32768;
    AttnBwdShared* sh = reinterpret_cast<AttnBwdShared*>(sY + 32768);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int b = blockIdx.x, slice = blockIdx.y;
    const int S = a.S, d = a.d;
    const int hps = 64 / d;
    const int n_t = (S + 127) / 128;          // 128-wide tiles of the sequence (rows and columns alike)
    const int subs_per_head = 2 * n_t * n_t;  // sweep A then sweep B

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmDO);
        mbar_init(&sh->loaded, 1);
        mbar_init(&sh->t_full, 1);
        mbar_init(&sh->x_full, 256);
        mbar_init(&sh->acc_done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&sh->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sh->tmem_base;
    const uint32_t tm_t1 = tmem_base, tm_t2 = tmem_base + 128, tm_acc0 = tmem_base + 256, tm_acc1 = tmem_base + 320;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&sh->loaded, 4 * 32768);
            tma_load_2d(sQ, &tmQKV, &sh->loaded, slice * 64, b * S);
            tma_load_2d(sK, &tmQKV, &sh->loaded, a.H + slice * 64, b * S);
            tma_load_2d(sV, &tmQKV, &sh->loaded, 2 * a.H + slice * 64, b * S);
            tma_load_2d(sDO, &tmDO, &sh->loaded, slice * 64, b * S);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait_lean(&sh->loaded, 0);
            tc_fence_after();
            const int ks_d = d / 16;
            // sub-tile i of a head: sweep = i / (n_t*n_t), rt = row tile, ct = column tile
            auto issue_t = [&](int hh, int i) {
                const int sweep = i / (n_t * n_t), rt = (i / n_t) % n_t, ct = i % n_t;
                const int ncols = min(128, S - ct * 128);
                const uint32_t idesc = at_idesc(128, (ncols + 15) / 16 * 16, 0, 0);
                // sweep A: T1 = Q[rt] K[ct]^T, T2 = dO[rt] V[ct]^T      sweep B: T1 = K[rt] Q[ct]^T, T2 = V[rt] dO[ct]^T
                uint8_t* a1 = sweep == 0 ? sQ : sK;
                uint8_t* b1 = sweep == 0 ? sK : sQ;
                uint8_t* a2 = sweep == 0 ? sDO : sV;
                uint8_t* b2 = sweep == 0 ? sV : sDO;
                const uint32_t koff = (uint32_t)(hh * d * 2);
                for (int ks = 0; ks < ks_d; ++ks)
                    umma_bf16(tm_t1, smem_desc_advance(smem_desc_sw128(smem_u32(a1 + (size_t)rt * 16384)), koff + ks * 32),
                              smem_desc_advance(smem_desc_sw128(smem_u32(b1 + (size_t)ct * 16384)), koff + ks * 32), idesc, (uint32_t)(ks != 0));
                for (int ks = 0; ks < ks_d; ++ks)
                    umma_bf16(tm_t2, smem_desc_advance(smem_desc_sw128(smem_u32(a2 + (size_t)rt * 16384)), koff + ks * 32),
                              smem_desc_advance(smem_desc_sw128(smem_u32(b2 + (size_t)ct * 16384)), koff + ks * 32), idesc, (uint32_t)(ks != 0));
                umma_commit(&sh->t_full);
            };
            int g = 0;                                        // global sub-tile counter (barrier phases)
            issue_t(0, 0);
            for (int hh = 0; hh < hps; ++hh) {
                for (int i = 0; i < subs_per_head; ++i, ++g) {
                    const int sweep = i / (n_t * n_t), ct = i % n_t;
                    const int ncols = min(128, S - ct * 128);
                    const int ks_c = (ncols + 15) / 16;
                    const uint32_t idesc_acc = at_idesc(128, d, 0, 1);
                    mbar_wait_lean(&sh->x_full, (uint32_t)g & 1u);
                    tc_fence_after();
                    const uint32_t acc_on = ct != 0;          // first column tile of a row tile overwrites the accumulators
                    const uint32_t coff = (uint32_t)(hh * d * 2);
                    for (int ks = 0; ks < ks_c; ++ks) {
                        const uint64_t xd = smem_desc_advance(smem_desc_sw128(smem_u32(sX + (size_t)(ks / 4) * 16384)), (ks % 4) * 32);
                        uint8_t* bsrc = sweep == 0 ? sK : sQ;  // dQ += dS K        |  dK += dS^T Q
                        const uint64_t bd = at_desc_mn(smem_u32(bsrc + (size_t)(ct * 128 + ks * 16) * 128 + coff), 0);
                        umma_bf16(tm_acc0, xd, bd, idesc_acc, (uint32_t)(acc_on | (ks != 0)));
                    }
                    if (sweep == 1) {
                        for (int ks = 0; ks < ks_c; ++ks) {   // dV += (P^T . drop) dO
                            const uint64_t yd = smem_desc_advance(smem_desc_sw128(smem_u32(sY + (size_t)(ks / 4) * 16384)), (ks % 4) * 32);
                            const uint64_t bd = at_desc_mn(smem_u32(sDO + (size_t)(ct * 128 + ks * 16) * 128 + coff), 0);
                            umma_bf16(tm_acc1, yd, bd, idesc_acc, (uint32_t)(acc_on | (ks != 0)));
                        }
                    }
                    umma_commit(&sh->acc_done);
                    // next sub-tile's T MMAs (its t_full commit also covers the accumulate MMAs above)
                    if (i + 1 < subs_per_head) issue_t(hh, i + 1);
                    else if (hh + 1 < hps) issue_t(hh + 1, 0);
                }
            }
        }
    } else if (warp >= 4) {
        // two warpgroups: both own all 128 rows of a sub-tile (TMEM lane = row) and split its 128 columns in halves
        const int wg = (warp - 4) / 4;
        const int q4 = warp % 4;
        const int r = q4 * 32 + lane;
        const int et = threadIdx.x - 128;                    // 0..255 among the epilogue threads
        const uint32_t lane_addr = ((uint32_t)(q4 * 32) << 16);
        const float cs = a.scale * LOG2E;
        if (warp == 4) {
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const int key = w * 32 + lane;
                bool ok = key < S;
                if (ok && a.key_valid) ok = a.key_valid[(size_t)b * S + key] != 0;
                const uint32_t bits = __ballot_sync(0xffffffffu, ok);
                if (lane == 0) sh->kvb[w] = bits;
            }
        }
        int g = 0;
        for (int hh = 0; hh < hps; ++hh) {
            const int head = slice * hps + hh;
            const long long bh = (long long)b * a.heads + head;
            asm volatile("bar.sync 1, 256;" ::: "memory");   // previous head's constants no longer in use (and kvb written)
            // per-query constants and dropout bits of this head
            for (int q = et; q < S; q += 256) {
                const __nv_bfloat16* o = a.ctx + ((size_t)b * S + q) * a.H + slice * 64 + hh * d;
                const __nv_bfloat16* go = a.d_ctx + ((size_t)b * S + q) * a.H + slice * 64 + hh * d;
                float D = 0.f;
                for (int c = 0; c < d; c += 8) {
                    const uint4 ov = *reinterpret_cast<const uint4*>(o + c);
                    const uint4 gv = *reinterpret_cast<const uint4*>(go + c);
                    const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __nv_bfloat162 o2 = *reinterpret_cast<const __nv_bfloat162*>(&ow[e]);
                        const __nv_bfloat162 g2 = *reinterpret_cast<const __nv_bfloat162*>(&gw[e]);
                        D += __low2float(o2) * __low2float(g2) + __high2float(o2) * __high2float(g2);
                    }
                }
                const float m = a.stats[bh * S + q];
                const float l = a.stats[(long long)a.B * a.heads * S + bh * S + q];
                sh->qc[q] = make_float4(m * LOG2E, 1.0f / l, D, exp2f((MASK_FILL - m) * LOG2E));
                if (a.keep_bits) {
                    const uint4* kb = reinterpret_cast<const uint4*>(a.keep_bits + ((size_t)bh * S + q) * 8);
                    reinterpret_cast<uint4*>(sh->keep + q * 8)[0] = kb[0];
                    reinterpret_cast<uint4*>(sh->keep + q * 8)[1] = kb[1];
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int i = 0; i < subs_per_head; ++i, ++g) {
                const int sweep = i / (n_t * n_t), rt = (i / n_t) % n_t, ct = i % n_t;
                const int row = rt * 128 + r;                  // sweep A: query, sweep B: key
                const bool row_ok = row < S;
                const int ncols = min(128, S - ct * 128);
                const int warp_row0 = rt * 128 + q4 * 32;      // first row of this warp
                mbar_wait_lean(&sh->t_full, (uint32_t)g & 1u);
                tc_fence_after();
                float4 rc = make_float4(0.f, 1.f, 0.f, 0.f);
                if (sweep == 0 && row_ok) rc = sh->qc[row];
                const bool key_row_ok = sweep == 1 && row_ok && ((sh->kvb[row >> 5] >> (row & 31)) & 1u);
                const bool warp_rows_plain = sweep == 0 ? (warp_row0 + 31 < S)
                                                        : (__all_sync(0xffffffffu, key_row_ok) != 0);
                uint32_t keep_w0 = 0xffffffffu, keep_w1 = 0xffffffffu;   // sweep A: keep bits of this row for this warpgroup's 64 keys
                if (sweep == 0 && a.keep_bits && row_ok) {
                    keep_w0 = sh->keep[row * 8 + ct * 4 + wg * 2];
                    keep_w1 = sh->keep[row * 8 + ct * 4 + wg * 2 + 1];
                }
#pragma unroll 1
                for (int c16 = wg * 4; c16 < wg * 4 + 4; ++c16) {          // 16-column chunks of this warpgroup's half
                    const int col0 = ct * 128 + c16 * 16;                  // sweep A: first key, sweep B: first query
                    if (c16 * 16 >= ((ncols + 31) / 32) * 32) break;       // beyond the columns the MMAs touch (warp-uniform)
                    float t1[16], t2[16];
                    tmem_ld16(tm_t1 + lane_addr + (uint32_t)(c16 * 16), t1);
                    tmem_ld16(tm_t2 + lane_addr + (uint32_t)(c16 * 16), t2);
                    tmem_ld_wait();
                    float pd[16], ds[16];
                    if (sweep == 0) {
                        const uint32_t kv16 = (sh->kvb[col0 >> 5] >> (col0 & 31)) & 0xffffu;
                        const uint32_t kp16 = ((((c16 >> 1) & 1) ? keep_w1 : keep_w0) >> ((c16 & 1) * 16)) & 0xffffu;
                        const bool plain = warp_rows_plain && kv16 == 0xffffu && (!a.causal || col0 + 15 <= warp_row0);
                        if (plain) {
#pragma unroll
                            for (int c = 0; c < 16; ++c) {
                                const float p = fast_exp2(fmaf(t1[c], cs, -rc.x)) * rc.y;
                                const float keep = ((kp16 >> c) & 1u) ? a.inv_keep : 0.f;
                                ds[c] = p * (t2[c] * keep - rc.z) * a.scale;
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < 16; ++c) {
                                const int col = col0 + c;
                                const bool ok = ((kv16 >> c) & 1u) && (!a.causal || col <= row);
                                const float e = ok ? fast_exp2(fmaf(t1[c], cs, -rc.x)) : rc.w;
                                const float p = (col < S && row_ok) ? e * rc.y : 0.f;
                                const float keep = ((kp16 >> c) & 1u) ? a.inv_keep : 0.f;
                                ds[c] = ok ? p * (t2[c] * keep - rc.z) * a.scale : 0.f;
                            }
                        }
                    } else {
                        const bool plain = warp_rows_plain && (col0 + 15 < S) && (!a.causal || warp_row0 + 31 <= col0);
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const int col = col0 + c;
                            const bool col_ok = plain || col < S;
                            const float4 cc = col_ok ? sh->qc[col] : make_float4(0.f, 1.f, 0.f, 0.f);
                            const bool ok = plain || (key_row_ok && col_ok && (!a.causal || row <= col));
                            const float e = ok ? fast_exp2(fmaf(t1[c], cs, -cc.x)) : cc.w;
                            const float p = (col_ok && row_ok) ? e * cc.y : 0.f;
                            float keep = a.inv_keep;
                            if (a.keep_bits && col_ok) keep = ((sh->keep[col * 8 + (row >> 5)] >> (row & 31)) & 1u) ? a.inv_keep : 0.f;
                            pd[c] = p * keep;
                            ds[c] = ok ? p * (t2[c] * keep - cc.z) * a.scale : 0.f;
                        }
                    }
                    uint8_t* xchunk = sX + (size_t)(c16 / 4) * 16384;      // 64 columns per chunk; this 16-column piece = 2 units
                    uint8_t* ychunk = sY + (size_t)(c16 / 4) * 16384;
#pragma unroll
                    for (int u16 = 0; u16 < 2; ++u16) {
                        uint4 w;
                        w.x = at_pack(ds[u16 * 8 + 0], ds[u16 * 8 + 1]); w.y = at_pack(ds[u16 * 8 + 2], ds[u16 * 8 + 3]);
                        w.z = at_pack(ds[u16 * 8 + 4], ds[u16 * 8 + 5]); w.w = at_pack(ds[u16 * 8 + 6], ds[u16 * 8 + 7]);
                        *reinterpret_cast<uint4*>(xchunk + sw128_offset(r, (c16 & 3) * 2 + u16)) = w;
                        if (sweep == 1) {
                            w.x = at_pack(pd[u16 * 8 + 0], pd[u16 * 8 + 1]); w.y = at_pack(pd[u16 * 8 + 2], pd[u16 * 8 + 3]);
                            w.z = at_pack(pd[u16 * 8 + 4], pd[u16 * 8 + 5]); w.w = at_pack(pd[u16 * 8 + 6], pd[u16 * 8 + 7]);
                            *reinterpret_cast<uint4*>(ychunk + sw128_offset(r, (c16 & 3) * 2 + u16)) = w;
                        }
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(&sh->x_full);
                if (ct == n_t - 1) {
                    // all column tiles of this row tile accumulated: store dQ (sweep A) or dK, dV (sweep B); each warpgroup
                    // stores half of the head's columns (d = 16: warpgroup 0 stores everything)
                    mbar_wait_lean(&sh->acc_done, (uint32_t)g & 1u);
                    tc_fence_after();
                    const int n_out = sweep == 0 ? 1 : 2;
                    const int dcols = d >= 32 ? d / 2 : d;
                    if (d >= 32 || wg == 0) {
                        const int cbeg = d >= 32 ? wg * dcols : 0;
                        for (int which = 0; which < n_out; ++which) {
                            const uint32_t tm = which == 0 ? tm_acc0 : tm_acc1;
                            const int colbase = (sweep == 0 ? 0 : (which == 0 ? a.H : 2 * a.H)) + slice * 64 + hh * d;
                            for (int c0 = cbeg; c0 < cbeg + dcols; c0 += 16) {
                                float o[16];
                                tmem_ld16(tm + lane_addr + (uint32_t)c0, o);
                                tmem_ld_wait();
                                if (row_ok) {
                                    __nv_bfloat16* dst = a.d_qkv + ((size_t)b * S + row) * 3 * a.H + colbase + c0;
                                    uint4 w0, w1;
                                    w0.x = at_pack(o[0], o[1]); w0.y = at_pack(o[2], o[3]); w0.z = at_pack(o[4], o[5]); w0.w = at_pack(o[6], o[7]);
                                    w1.x = at_pack(o[8], o[9]); w1.y = at_pack(o[10], o[11]); w1.z = at_pack(o[12], o[13]); w1.w = at_pack(o[14], o[15]);
                                    reinterpret_cast<uint4*>(dst)[0] = w0;
                                    reinterpret_cast<uint4*>(dst)[1] = w1;
                                }
                            }
                        }
                    }
                    tc_fence_before();
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}


// ------------------------------------------------------------------------------------------------------------
// backward, single sweep (default).  The two-sweep kernel above recomputes every score tile twice (once with queries as rows
// for dQ, once transposed for dK / dV) and its transposed sweep is the expensive one (per-element constants from shared
// memory).  Here each 128 x 128 sub-tile is visited ONCE with queries as rows: dS -> X and P.drop -> Y are written to shared
// memory as before, and the same two tiles feed three accumulations
//     dQ[rt] += X    K[ct]      (X as a K-major A operand, as before)
//     dK[ct] += X^T  Q[rt]      (X as an MN-major A operand: the same bytes, contraction running down the query rows)
//     dV[ct] += Y^T  dO[rt]
// Loop order: key tile ct outer, query tile rt inner, so dK / dV accumulate over the inner loop and every query tile keeps
// its own dQ accumulator: tensor memory = 2 x 128 (T1, T2) + (2 + n_t) x d <= 512 columns.
// ------------------------------------------------------------------------------------------------------------
// WGS epilogue warpgroups (2 or 4), each owning 128 / WGS columns of a sub-tile: the element-wise stage is latency-bound (23 % issue
// utilisation with two warpgroups), more resident warps hide it.
// Order of work, from a clock64 timeline of one CTA (B = 256, S = 200, d = 32: 66 k cycles per CTA before, 53 k after):
//  * the MMA warp issues the score / dP MMAs of sub-tile i+1 BEFORE the three accumulations of sub-tile i (which take ~2 k cycles of
//    tensor-pipe time: 24 MMAs of 128 x d x 16 are operand-fetch bound), so the element-wise stage of i+1 starts at once and waits for
//    acc_done(i) only before its first write to X / Y;
//  * the finished dQ / dK / dV tiles of sub-tile i are stored during sub-tile i+1 (after that same wait) instead of right after the
//    x_full arrival of i, where every storing warp idled through its own accumulations;
//  * a warp whose 32 queries do not exist (ragged last query tile) writes zeros, the warp holding the last real queries runs the plain
//    arithmetic and zeroes its missing rows afterwards (it was the slowest warp of the tile by 3.6 k cycles on the general path);
//  * the per-query constants of two heads are gathered in one pass (the pass is global-load latency; 2 S (head, query) pairs fill the
//    epilogue threads).
template <int WGS>
__global__ void __launch_bounds__(128 + 128 * WGS, 1) attn_tc_bwd1_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                                     const __grid_constant__ CUtensorMap tmDO, const AttnBwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + 32768;
    uint8_t* sV = sK + 32768;
    uint8_t* sDO = sV + 32768;
    uint8_t* sX = sDO + 32768;                // dS:     [2 chunks of 64 keys][128 queries][128 B]
    uint8_t* sY = sX + 32768;                 // P.drop: same layout
    AttnBwdShared* sh = reinterpret_cast<AttnBwdShared*>(sY + 32768);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int b = blockIdx.x, slice = blockIdx.y;
    const int S = a.S, d = a.d;
    const int hps = 64 / d;
    const int n_t = (S + 127) / 128;
    const int subs_per_head = n_t * n_t;      // sub-tile i: key tile ct = i / n_t, query tile rt = i % n_t

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmDO);
        mbar_init(&sh->loaded, 1);
        mbar_init(&sh->t_full, 1);
        mbar_init(&sh->x_full, 4 * WGS);               // one arrival per epilogue warp
        mbar_init(&sh->acc_done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&sh->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sh->tmem_base;
    const uint32_t tm_t1 = tmem_base, tm_t2 = tmem_base + 128, tm_dk = tmem_base + 256, tm_dv = tmem_base + 256 + d,
                   tm_dq0 = tmem_base + 256 + 2 * d;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&sh->loaded, 4 * 32768);
            tma_load_2d(sQ, &tmQKV, &sh->loaded, slice * 64, b * S);
            tma_load_2d(sK, &tmQKV, &sh->loaded, a.H + slice * 64, b * S);
            tma_load_2d(sV, &tmQKV, &sh->loaded, 2 * a.H + slice * 64, b * S);
            tma_load_2d(sDO, &tmDO, &sh->loaded, slice * 64, b * S);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait_lean(&sh->loaded, 0);
            tc_fence_after();
            const int ks_d = d / 16;
            auto issue_t = [&](int hh, int i) {
                const int ct = i / n_t, rt = i % n_t;
                const int ncols = min(128, S - ct * 128);
                const uint32_t idesc = at_idesc(128, (ncols + 15) / 16 * 16, 0, 0);
                const uint32_t koff = (uint32_t)(hh * d * 2);
                // T1 = Q[rt] K[ct]^T (scores),  T2 = dO[rt] V[ct]^T (dP)
                for (int ks = 0; ks < ks_d; ++ks)
                    umma_bf16(tm_t1, smem_desc_advance(smem_desc_sw128(smem_u32(sQ + (size_t)rt * 16384)), koff + ks * 32),
                              smem_desc_advance(smem_desc_sw128(smem_u32(sK + (size_t)ct * 16384)), koff + ks * 32), idesc, (uint32_t)(ks != 0));
                for (int ks = 0; ks < ks_d; ++ks)
                    umma_bf16(tm_t2, smem_desc_advance(smem_desc_sw128(smem_u32(sDO + (size_t)rt * 16384)), koff + ks * 32),
                              smem_desc_advance(smem_desc_sw128(smem_u32(sV + (size_t)ct * 16384)), koff + ks * 32), idesc, (uint32_t)(ks != 0));
                umma_commit(&sh->t_full);
            };
            const uint32_t idesc_k = at_idesc(128, d, 0, 1);      // A K-major  (X),    B MN-major
            const uint32_t idesc_mn = at_idesc(128, d, 1, 1);     // A MN-major (X^T),  B MN-major
            int g = 0;
            issue_t(0, 0);
            for (int hh = 0; hh < hps; ++hh) {
                const uint32_t coff = (uint32_t)(hh * d * 2);
                for (int i = 0; i < subs_per_head; ++i, ++g) {
                    const int ct = i / n_t, rt = i % n_t;
                    const int ks_c = (min(128, S - ct * 128) + 15) / 16;      // K-steps over the keys of the tile
                    const int ks_r = (min(128, S - rt * 128) + 15) / 16;      // K-steps over the queries of the tile
                    mbar_wait_lean(&sh->x_full, (uint32_t)g & 1u);
                    tc_fence_after();
                    // the score / dP MMAs of the NEXT sub-tile go first (T1 / T2 have been read: x_full): its element-wise stage
                    // starts while the three accumulations below stream X and Y, and waits for acc_done only before its first
                    // write to them
                    if (i + 1 < subs_per_head) issue_t(hh, i + 1);
                    else if (hh + 1 < hps) issue_t(hh + 1, 0);
                    const uint32_t tm_dq = tm_dq0 + (uint32_t)(rt * d);
                    for (int ks = 0; ks < ks_c; ++ks) {       // dQ[rt] += dS K[ct]
                        const uint64_t xd = smem_desc_advance(smem_desc_sw128(smem_u32(sX + (size_t)(ks / 4) * 16384)), (ks % 4) * 32);
                        const uint64_t bd = at_desc_mn(smem_u32(sK + (size_t)(ct * 128 + ks * 16) * 128 + coff), 0);
                        umma_bf16(tm_dq, xd, bd, idesc_k, (uint32_t)((ct != 0) | (ks != 0)));
                    }
                    for (int ks = 0; ks < ks_r; ++ks) {       // dK[ct] += dS^T Q[rt]
                        const uint64_t ad = at_desc_mn(smem_u32(sX + (size_t)ks * 2048), 16384);
                        const uint64_t bd = at_desc_mn(smem_u32(sQ + (size_t)(rt * 128 + ks * 16) * 128 + coff), 0);
                        umma_bf16(tm_dk, ad, bd, idesc_mn, (uint32_t)((rt != 0) | (ks != 0)));
                    }
                    for (int ks = 0; ks < ks_r; ++ks) {       // dV[ct] += (P.drop)^T dO[rt]
                        const uint64_t ad = at_desc_mn(smem_u32(sY + (size_t)ks * 2048), 16384);
                        const uint64_t bd = at_desc_mn(smem_u32(sDO + (size_t)(rt * 128 + ks * 16) * 128 + coff), 0);
                        umma_bf16(tm_dv, ad, bd, idesc_mn, (uint32_t)((rt != 0) | (ks != 0)));
                    }
                    umma_commit(&sh->acc_done);
                }
            }
        }
    } else if (warp >= 4) {
        const int wg = (warp - 4) / 4;
        const int q4 = warp % 4;
        const int r = q4 * 32 + lane;
        const int et = threadIdx.x - 128;                    // 0 .. 128*WGS-1 among the epilogue threads
        constexpr int C16_PER_WG = 8 / WGS;                  // 16-column chunks of a sub-tile per warpgroup
        const uint32_t lane_addr = ((uint32_t)(q4 * 32) << 16);
        const float cs = a.scale * LOG2E;
        if (warp == 4) {
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const int key = w * 32 + lane;
                bool ok = key < S;
                if (ok && a.key_valid) ok = a.key_valid[(size_t)b * S + key] != 0;
                const uint32_t bits = __ballot_sync(0xffffffffu, ok);
                if (lane == 0) sh->kvb[w] = bits;
            }
        }
        // dQ (rows = queries of tile rt), dK, dV (rows = keys of tile ct) of head hh: tensor memory -> bf16 -> global; the head's d
        // output columns go in 16-column pieces dealt to the warpgroups
        auto store_tiles = [&](int hh_s, int ct_s, int rt_s, bool store_q, bool store_kv) {
            for (int which = store_q ? 0 : 1; which < (store_kv ? 3 : 1); ++which) {      // 0 = dQ, 1 = dK, 2 = dV
                const uint32_t tm = which == 0 ? tm_dq0 + (uint32_t)(rt_s * d) : (which == 1 ? tm_dk : tm_dv);
                const int out_row = (which == 0 ? rt_s : ct_s) * 128 + r;
                const int colbase = which * a.H + slice * 64 + hh_s * d;
                for (int c0 = (WGS - 1 - wg) * 16; c0 < d; c0 += WGS * 16) {      // last warpgroups first: a ragged key tile leaves them the fewest chunks
                    float o[16];
                    tmem_ld16(tm + lane_addr + (uint32_t)c0, o);
                    tmem_ld_wait();
                    if (out_row < S) {
                        __nv_bfloat16* dst = a.d_qkv + ((size_t)b * S + out_row) * 3 * a.H + colbase + c0;
                        uint4 w0, w1;
                        w0.x = at_pack(o[0], o[1]); w0.y = at_pack(o[2], o[3]); w0.z = at_pack(o[4], o[5]); w0.w = at_pack(o[6], o[7]);
                        w1.x = at_pack(o[8], o[9]); w1.y = at_pack(o[10], o[11]); w1.z = at_pack(o[12], o[13]); w1.w = at_pack(o[14], o[15]);
                        reinterpret_cast<uint4*>(dst)[0] = w0;
                        reinterpret_cast<uint4*>(dst)[1] = w1;
                    }
                }
            }
        };
        bool pend_q = false, pend_kv = false;
        int pend_hh = 0, pend_ct = 0, pend_rt = 0;
        int g = 0;
        for (int hh = 0; hh < hps; ++hh) {
            if ((hh & 1) == 0) {
                // per-query constants and keep bits of this head and the next in ONE pass (a pass is pure global-load latency: 2 x S
                // (head, query) pairs keep 400 of the 512 threads busy at S = 200 instead of 200 of them twice)
                asm volatile("bar.sync 1, %0;" ::"n"(128 * WGS) : "memory");      // the previous pair's constants are no longer in use
                const int n_pair = min(2, hps - hh);
                for (int idx = et; idx < n_pair * S; idx += 128 * WGS) {
                    const int h2 = idx >= S ? 1 : 0, q = idx - h2 * S;
                    const long long bh = (long long)b * a.heads + slice * hps + hh + h2;
                    const __nv_bfloat16* o = a.ctx + ((size_t)b * S + q) * a.H + slice * 64 + (hh + h2) * d;
                    const __nv_bfloat16* go = a.d_ctx + ((size_t)b * S + q) * a.H + slice * 64 + (hh + h2) * d;
                    const float m_in = a.stats[bh * S + q];
                    const float l_in = a.stats[(long long)a.B * a.heads * S + bh * S + q];
                    uint4 kb0 = make_uint4(0u, 0u, 0u, 0u), kb1 = kb0;
                    if (a.keep_bits) {
                        const uint4* kb = reinterpret_cast<const uint4*>(a.keep_bits + ((size_t)bh * S + q) * 8);
                        kb0 = kb[0]; kb1 = kb[1];
                    }
                    float D = 0.f;
#pragma unroll 4
                    for (int c = 0; c < d; c += 8) {          // (unrolled: the 16-byte loads of a row are in flight together)
                        const uint4 ov = __ldg(reinterpret_cast<const uint4*>(o + c));
                        const uint4 gv = __ldg(reinterpret_cast<const uint4*>(go + c));
                        const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __nv_bfloat162 o2 = *reinterpret_cast<const __nv_bfloat162*>(&ow[e]);
                            const __nv_bfloat162 g2 = *reinterpret_cast<const __nv_bfloat162*>(&gw[e]);
                            D += __low2float(o2) * __low2float(g2) + __high2float(o2) * __high2float(g2);
                        }
                    }
                    sh->qc[h2 * AT_MAXS + q] = make_float4(m_in * LOG2E, 1.0f / l_in, D, exp2f((MASK_FILL - m_in) * LOG2E));
                    if (a.keep_bits) {
                        reinterpret_cast<uint4*>(sh->keep + (h2 * AT_MAXS + q) * 8)[0] = kb0;
                        reinterpret_cast<uint4*>(sh->keep + (h2 * AT_MAXS + q) * 8)[1] = kb1;
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(128 * WGS) : "memory");
            }
            const float4* qc_h = sh->qc + (hh & 1) * AT_MAXS;
            const uint32_t* keep_h = sh->keep + (hh & 1) * AT_MAXS * 8;
            for (int i = 0; i < subs_per_head; ++i, ++g) {
                const int ct = i / n_t, rt = i % n_t;
                const int row = rt * 128 + r;                  // query of this thread
                const bool row_ok = row < S;
                const int ncols = min(128, S - ct * 128);
                const int warp_row0 = rt * 128 + q4 * 32;
                mbar_wait_lean(&sh->t_full, (uint32_t)g & 1u);
                tc_fence_after();
                float4 rc = make_float4(0.f, 1.f, 0.f, 0.f);
                if (row_ok) rc = qc_h[row];
                const bool warp_rows_plain = warp_row0 + 31 < S;
                const bool warp_rows_none = warp_row0 >= S;          // no query of this warp exists: dS = P = 0
                bool xy_free = g == 0;                               // the accumulate MMAs of the previous sub-tile have read X and Y
#pragma unroll 1
                for (int c16 = wg * C16_PER_WG; c16 < (wg + 1) * C16_PER_WG; ++c16) {
                    const int col0 = ct * 128 + c16 * 16;                  // first key of the chunk
                    if (c16 * 16 >= ((ncols + 31) / 32) * 32) break;
                    const uint32_t kv16 = (sh->kvb[col0 >> 5] >> (col0 & 31)) & 0xffffu;
                    // padding keys only and no fully masked row in this warp, or no query of this warp exists: P = dS = 0 exactly
                    // (warp-uniform test)
                    if (warp_rows_none || (kv16 == 0u && __all_sync(0xffffffffu, rc.w == 0.f))) {
                        if (!xy_free) { mbar_wait_lean(&sh->acc_done, (uint32_t)(g - 1) & 1u); xy_free = true; }
#pragma unroll
                        for (int u16 = 0; u16 < 2; ++u16) {
                            const uint32_t off = sw128_offset(r, (c16 & 3) * 2 + u16);
                            *reinterpret_cast<uint4*>(sX + (size_t)(c16 / 4) * 16384 + off) = make_uint4(0u, 0u, 0u, 0u);
                            *reinterpret_cast<uint4*>(sY + (size_t)(c16 / 4) * 16384 + off) = make_uint4(0u, 0u, 0u, 0u);
                        }
                        continue;
                    }
                    float t1[16], t2[16];
                    tmem_ld16(tm_t1 + lane_addr + (uint32_t)(c16 * 16), t1);
                    tmem_ld16(tm_t2 + lane_addr + (uint32_t)(c16 * 16), t2);
                    tmem_ld_wait();
                    float pd[16], ds[16];
                    uint32_t keep_w = 0xffffffffu;                      // keep bits of this row for the 32 keys around the chunk
                    if (a.keep_bits && row_ok) keep_w = keep_h[row * 8 + ct * 4 + (c16 >> 1)];
                    const uint32_t kp16 = (keep_w >> ((c16 & 1) * 16)) & 0xffffu;
                    // every key of the chunk attendable by every query of the warp that exists (missing queries: zeroed after the math)
                    const bool plain = kv16 == 0xffffu && (!a.causal || col0 + 15 <= warp_row0);
                    if (plain) {
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const float p = fast_exp2(fmaf(t1[c], cs, -rc.x)) * rc.y;
                            const float keep = ((kp16 >> c) & 1u) ? a.inv_keep : 0.f;
                            pd[c] = p * keep;
                            ds[c] = p * (t2[c] * keep - rc.z) * a.scale;
                        }
                        if (!warp_rows_plain) {                         // warp-uniform: the last warp with queries of a ragged tile
#pragma unroll
                            for (int c = 0; c < 16; ++c) {
                                pd[c] = row_ok ? pd[c] : 0.f;
                                ds[c] = row_ok ? ds[c] : 0.f;
                            }
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const int col = col0 + c;
                            const bool ok = ((kv16 >> c) & 1u) && (!a.causal || col <= row);
                            const float e = ok ? fast_exp2(fmaf(t1[c], cs, -rc.x)) : rc.w;
                            const float p = (col < S && row_ok) ? e * rc.y : 0.f;
                            const float keep = ((kp16 >> c) & 1u) ? a.inv_keep : 0.f;
                            pd[c] = p * keep;
                            ds[c] = ok ? p * (t2[c] * keep - rc.z) * a.scale : 0.f;
                        }
                    }
                    // X and Y are written only now, and only now must the previous sub-tile's accumulate MMAs have finished reading them
                    if (!xy_free) { mbar_wait_lean(&sh->acc_done, (uint32_t)(g - 1) & 1u); xy_free = true; }
                    uint8_t* xchunk = sX + (size_t)(c16 / 4) * 16384;
                    uint8_t* ychunk = sY + (size_t)(c16 / 4) * 16384;
#pragma unroll
                    for (int u16 = 0; u16 < 2; ++u16) {
                        uint4 w;
                        w.x = at_pack(ds[u16 * 8 + 0], ds[u16 * 8 + 1]); w.y = at_pack(ds[u16 * 8 + 2], ds[u16 * 8 + 3]);
                        w.z = at_pack(ds[u16 * 8 + 4], ds[u16 * 8 + 5]); w.w = at_pack(ds[u16 * 8 + 6], ds[u16 * 8 + 7]);
                        *reinterpret_cast<uint4*>(xchunk + sw128_offset(r, (c16 & 3) * 2 + u16)) = w;
                        w.x = at_pack(pd[u16 * 8 + 0], pd[u16 * 8 + 1]); w.y = at_pack(pd[u16 * 8 + 2], pd[u16 * 8 + 3]);
                        w.z = at_pack(pd[u16 * 8 + 4], pd[u16 * 8 + 5]); w.w = at_pack(pd[u16 * 8 + 6], pd[u16 * 8 + 7]);
                        *reinterpret_cast<uint4*>(ychunk + sw128_offset(r, (c16 & 3) * 2 + u16)) = w;
                    }
                }
                // the gradient tiles the PREVIOUS sub-tile completed leave now: their accumulate MMAs have run next to this
                // sub-tile's element-wise stage (acc_done was awaited before the first write to X / Y above; warps without
                // such a write wait here), so no warp ever idles through the accumulate MMAs of its own sub-tile
                if (pend_q || pend_kv) {
                    if (!xy_free) { mbar_wait_lean(&sh->acc_done, (uint32_t)(g - 1) & 1u); xy_free = true; }
                    tc_fence_after();
                    store_tiles(pend_hh, pend_ct, pend_rt, pend_q, pend_kv);
                    tc_fence_before();
                }
                fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive_warp(&sh->x_full);
                pend_kv = rt == n_t - 1; pend_q = ct == n_t - 1;
                pend_hh = hh; pend_ct = ct; pend_rt = rt;
            }
        }
        if (pend_q || pend_kv) {                              // the last sub-tile's
            mbar_wait_lean(&sh->acc_done, (uint32_t)(g - 1) & 1u);
            tc_fence_after();
            store_tiles(pend_hh, pend_ct, pend_rt, pend_q, pend_kv);
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static int g_attn_bwd_variant = 1;      // 1: single sweep (default), 0: the two-sweep kernel (kept for A/B measurements)
static int g_attn_bwd_wgs = 4;          // epilogue warpgroups of the single-sweep kernel: 2 or 4 (4: 3-6 % faster, bit-identical)
extern "C" int asme_b200_tc_attn_tune(int knob, int value) {
    if (knob == 0) {
        ASME_REQUIRE(value == 0 || value == 1, "tc_attn_tune: knob 0 (backward variant) takes 0 or 1");
        g_attn_bwd_variant = value;
    } else if (knob == 1) {
        ASME_REQUIRE(value == 2 || value == 4, "tc_attn_tune: knob 1 (epilogue warpgroups of the single-sweep backward) takes 2 or 4");
        g_attn_bwd_wgs = value;
    } else if (knob == 2) {
        ASME_REQUIRE(value >= 1 && value <= 3, "tc_attn_tune: knob 2 (forward kernel) takes 1, 2 or 3 (3: the second kernel for every grid size)");
        g_attn_fwd_variant = value;
    } else {
        ASME_REQUIRE(false, "tc_attn_tune: unknown knob %d", knob);
    }
    return ASME_OK;
}

extern "C" int asme_b200_tc_attn_bwd(const void* qkv, const uint8_t* key_valid, int B, int S, int heads, int d, int causal,
                                     float p_drop, const void* ctx, const void* d_ctx, const float* stats,
                                     const uint32_t* keep_bits, void* d_qkv, asme_stream_t stream) {
    ASME_REQUIRE(qkv && ctx && d_ctx && stats && d_qkv, "tc_attn_bwd: null argument");
    ASME_REQUIRE(S >= 1 && S <= AT_MAXS, "tc_attn_bwd: S=%d unsupported (1..%d)", S, AT_MAXS);
    ASME_REQUIRE(d == 16 || d == 32 || d == 64, "tc_attn_bwd: head dim %d unsupported (16, 32, 64)", d);
    const int H = heads * d;
    ASME_REQUIRE(H % 64 == 0, "tc_attn_bwd: hidden size %d must be a multiple of 64", H);
    ASME_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "tc_attn_bwd: p_drop=%f", p_drop);
    ASME_REQUIRE(p_drop == 0.f || keep_bits, "tc_attn_bwd: keep_bits of the forward pass are required when p_drop > 0");
    if (B == 0) return ASME_OK;
    CUtensorMap tmQ, tmD;
    int rc = asme_tc_make_tmap_bf16(&tmQ, qkv, (long long)B * S, 3 * H, 3 * H, 256);
    if (rc) return rc;
    rc = asme_tc_make_tmap_bf16(&tmD, d_ctx, (long long)B * S, H, H, 256);
    if (rc) return rc;
    AttnBwdArgs a{};
    a.key_valid = key_valid; a.B = B; a.S = S; a.heads = heads; a.d = d; a.H = H; a.causal = causal;
    a.scale = 1.0f / sqrtf((float)d); a.p_drop = p_drop; a.inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    a.ctx = (const __nv_bfloat16*)ctx; a.d_ctx = (const __nv_bfloat16*)d_ctx; a.stats = stats;
    a.keep_bits = p_drop > 0.f ? keep_bits : nullptr; a.d_qkv = (__nv_bfloat16*)d_qkv;
    const size_t smem = 1024 + 6 * 32768 + sizeof(AttnBwdShared);
    if (g_attn_bwd_variant == 1) {
        if (g_attn_bwd_wgs == 4) {
            { const int _rc = asme_ensure_max_smem((const void*)attn_tc_bwd1_kernel<4>); if (_rc) return _rc; }
            attn_tc_bwd1_kernel<4><<<dim3(B, H / 64), 128 + 128 * 4, smem, (cudaStream_t)stream>>>(tmQ, tmD, a);
        } else {
            { const int _rc = asme_ensure_max_smem((const void*)attn_tc_bwd1_kernel<2>); if (_rc) return _rc; }
            attn_tc_bwd1_kernel<2><<<dim3(B, H / 64), 128 + 128 * 2, smem, (cudaStream_t)stream>>>(tmQ, tmD, a);
        }
    } else {
        { const int _rc = asme_ensure_max_smem((const void*)attn_tc_bwd_kernel); if (_rc) return _rc; }
        attn_tc_bwd_kernel<<<dim3(B, H / 64), ATF_THREADS, smem, (cudaStream_t)stream>>>(tmQ, tmD, a);
    }
    ASME_LAUNCH_OK();
    return ASME_OK;
}
