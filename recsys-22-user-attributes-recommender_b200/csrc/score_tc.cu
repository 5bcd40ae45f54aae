// Full-catalog scoring on the 5th-generation tensor cores (tcgen05 + TMEM, operands staged by TMA), fused with
//   * top-k (score desc, id asc) + capture of the target's score          (evaluation, K12+K15+K18-K20)
//   * exact rank counts  #{s_j > s_t}, #{j < t : s_j == s_t}              (rank / full-MRR metrics)
//   * online log-softmax partials (max, sum-exp, target logit)            (cross-entropy forward, K12+K16)
// so the (rows x V) logits only ever exist as 128 x 256 fp32 accumulator tiles in tensor memory.
//
// replaces: ItemEmbeddingProjectionLayer / LinearProjectionLayer (models/common/layers/layers.py:105-143) followed by
//           AllItemsSampler + per-metric argsort (metrics/container/metrics_sampler.py:45-71, metrics/common.py:18-27)
//           or nn.CrossEntropyLoss (modules/masked_training_module.py:107-111).
//
// Kernel shape (one CTA per SM, 384 threads, cta_group::1):
//   warp 0      TMA producer: A = 128 hidden rows (loaded once, stationary), B = 256-item table tiles, ring of stages
//   warp 1      MMA issuer  : per tile kch*4 tcgen05.mma (M=128, N=256, K=16), accumulators double-buffered in TMEM
//   warp 2      TMEM allocator (512 columns = 2 accumulator stages)
//   warps 4-11  epilogue: two warpgroups, each owning 128 of the tile's 256 columns; thread = one row (TMEM lane)
// CTA (m_tile, split) sweeps the item tiles [split*tps, (split+1)*tps); per-(split, warpgroup) partial results are
// merged by a second tiny kernel.
#include "common.cuh"
#include "tc_common.cuh"
#include <type_traits>
#include <string.h>

#include <limits.h>

using namespace tc;

// ------------------------------------------------------------------------------------------------------------
// host: tensor-map encoder through the runtime's driver entry point
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static int make_tmap_bf16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows, int box_cols,
                          CUtensorMapSwizzle swz);
int asme_tc_make_tmap_bf16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
    return make_tmap_bf16(map, base, rows, cols, ld, box_rows, CHUNK_K, CU_TENSOR_MAP_SWIZZLE_128B);
}
// fp32 row blocks for epilogues that read activations through TMA: box (16 columns x box_rows), 64-byte swizzle
static int ensure_encoder();
int asme_tc_make_tmap_f32_16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
    { const int rc = ensure_encoder(); if (rc) return rc; }
    ASME_REQUIRE(((uintptr_t)base & 15) == 0 && (ld * 4) % 16 == 0 && cols % 16 == 0, "tensor map (fp32): base / row stride / width must be 16-aligned");
    ASME_REQUIRE(box_rows >= 1 && box_rows <= 256, "tensor map (fp32): box_rows=%d", box_rows);
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {16u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        asme_set_error("cuTensorMapEncodeTiled (fp32) failed with CUresult %d (rows=%lld cols=%lld ld=%lld box_rows=%d)", (int)r, rows, cols, ld, box_rows);
        return ASME_ERR_CUDA;
    }
    return ASME_OK;
}
// 16-column boxes with the 32-byte swizzle: the K tail of operands whose padded width is 64 k + 16 (bias folded into the
// contraction) -- a quarter of the bytes of a 64-column box
static int make_tmap_bf16_tail16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
    return make_tmap_bf16(map, base, rows, cols, ld, box_rows, 16, CU_TENSOR_MAP_SWIZZLE_32B);
}
static int ensure_encoder() {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        ASME_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) {
            asme_set_error("cuTensorMapEncodeTiled is not available from this driver");
            return ASME_ERR_CUDA;
        }
        g_encode = (PFN_encodeTiled)fn;
    }
    return ASME_OK;
}
static int make_tmap_bf16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows, int box_cols,
                          CUtensorMapSwizzle swz) {
    { const int rc = ensure_encoder(); if (rc) return rc; }
    ASME_REQUIRE(((uintptr_t)base & 15) == 0 && (ld * 2) % 16 == 0, "tensor map: base / row stride must be 16-byte aligned");
    ASME_REQUIRE(box_rows >= 1 && box_rows <= 256, "tensor map: box_rows=%d", box_rows);
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        asme_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld box_rows=%d)", (int)r, rows, cols,
                       ld, box_rows);
        return ASME_ERR_CUDA;
    }
    return ASME_OK;
}

// ------------------------------------------------------------------------------------------------------------
// fp32 -> bf16 operand preparation (round to nearest even), zero padded to ld_out columns
// ------------------------------------------------------------------------------------------------------------
__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long rows, int cols, int ld_in,
                                 int ld_out, const int32_t* __restrict__ n_live) {
    const int per_row = ld_out / 4;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * per_row) return;
    const long long r = i / per_row;
    const int c = (int)(i % per_row) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* src = x + r * ld_in + c;
    // rows past the live count are written as zeros (not skipped): the result is a TMA / MMA operand, and stale bits there could
    // be NaNs that a masked-out (x 0) product would still propagate
    if (n_live != nullptr && r >= (long long)__ldg(n_live)) { /* zeros */ }
    else if (c + 3 < cols && (ld_in % 4) == 0) v = ldg4(src);
    else {
        if (c + 0 < cols) v.x = src[0];
        if (c + 1 < cols) v.y = src[1];
        if (c + 2 < cols) v.z = src[2];
        if (c + 3 < cols) v.w = src[3];
    }
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(y + r * ld_out + c) = o;
}

// y[r, :cols] = bf16(x[r, :cols]); y[r, cols], y[r, cols+1] = the two extra columns that fold a bias into the contraction:
//   mode 1 (activations): 1, 1        mode 2 (weights): bf16(bias[r]), bf16(bias[r] - bf16(bias[r]))   (hi + lo: 2^-17 relative)
// remaining columns up to ld_out are zero.  h.w + b  ==  [h, 1, 1].[w, b_hi, b_lo]  -- the scoring epilogue no longer adds a bias.
__global__ void cast_bf16_ext_kernel(const float* __restrict__ x, const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                                     long long rows, int cols, int ld_in, int ld_out, int mode) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ld_out) return;
    const long long r = i / ld_out;
    const int c = (int)(i % ld_out);
    float v = 0.f;
    if (c < cols) v = x[r * ld_in + c];
    else if (c < cols + 2) {
        if (mode == 1) v = 1.f;
        else {
            const float b = bias[r];
            const float hi = __bfloat162float(__float2bfloat16_rn(b));
            v = c == cols ? hi : b - hi;
        }
    }
    y[i] = __float2bfloat16_rn(v);
}
extern "C" int asme_b200_cast_bf16_ext(const float* x, const float* bias, void* y, long long rows, int cols, int ld_in, int ld_out,
                                       int mode, asme_stream_t stream) {
    ASME_REQUIRE(x && y && (mode == 1 || (mode == 2 && bias)), "cast_bf16_ext: bad argument");
    ASME_REQUIRE(cols >= 1 && ld_in >= cols && ld_out >= cols + 2, "cast_bf16_ext: cols=%d ld_in=%d ld_out=%d", cols, ld_in, ld_out);
    if (rows == 0) return ASME_OK;
    const long long n = rows * ld_out;
    cast_bf16_ext_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, bias, (__nv_bfloat16*)y, rows, cols, ld_in, ld_out, mode);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

extern "C" int asme_b200_cast_bf16(const float* x, void* y, long long rows, int cols, int ld_in, int ld_out, const int32_t* n_live,
                                   asme_stream_t stream) {
    ASME_REQUIRE(x && y, "cast_bf16: null argument");
    ASME_REQUIRE(cols >= 1 && ld_in >= cols && ld_out >= cols && ld_out % 4 == 0, "cast_bf16: cols=%d ld_in=%d ld_out=%d", cols, ld_in,
                 ld_out);
    if (rows == 0) return ASME_OK;
    const long long n = rows * (ld_out / 4);
    cast_bf16_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)y, rows, cols, ld_in, ld_out, n_live);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ------------------------------------------------------------------------------------------------------------
// the scoring kernel
// ------------------------------------------------------------------------------------------------------------
#define BM 128
#define BN 256
#define A_CHUNK_BYTES (BM * CHUNK_ROW_BYTES)   // 16 KB
#define B_CHUNK_BYTES (BN * CHUNK_ROW_BYTES)   // 32 KB
#define MAX_EPI_WGS 4           // epilogue warpgroups (template parameter WGS: 2 or 4): each owns 256/WGS columns of every tile
#define TC_THREADS(WGS) (128 + 128 * (WGS))
#define EPI_WARP0 4
#define MAX_STAGES 6
#define L2E 1.4426950408889634f

enum { EPI_TOPK = 0, EPI_CE = 1, EPI_PROBE = 2 };   // EPI_PROBE: diagnostic, epilogue = TMEM handshake only (pipeline ceiling)

struct ScoreTcArgs {
    int R, Vloc, v0, k, kch, stages;
    int last_ksteps;             // K-steps (of 16 columns) in the last 64-wide chunk: 4, or fewer when Kp % 64 != 0 (bias columns)
    int pend_cap;                // FIFO entries per epilogue thread (drained when more than pend_cap - 8 are pending)
    int tail16;                  // 1: the last chunk is ONE K-step staged as a 16-column box with the 32-byte swizzle (tmAt / tmBt)
    int n_tiles, tiles_per_split;
    int tile_lo, tile_hi;        // this launch sweeps tiles [split*tps + tile_lo, min(n_tiles, split*tps + tile_hi)) of every split
    int part0;                   // first partial-result slot written by this launch
    const float* thr_init;       // (R, thr_stride) or NULL: a lower bound of each row's k-th best score (from a sample sweep)
    int thr_stride, thr_col;
    int sample_mode;             // 1: threshold pass -- only the maximum of every 32-column chunk is offered to the lists (the k-th best
                                 //    chunk maximum is a lower bound of the k-th best score: k distinct chunks hold a score >= it)
    float thr_floor;             // diagnostic: -inf normally; +inf makes every list reject everything (cost of the insertion-free sweep)
    const float* bias;
    const float2* bias_bounds;   // (ceil(Vloc/32)) {max, min} of the bias over every 32-column chunk, or NULL.  With it the top-k sweeps
                                 // run the bias-free fast path and add the bias only to chunks whose best raw score + max bias could
                                 // pass the row's threshold (asme_b200_bias_chunk_bounds)
    const int64_t* target;
    const float* target_score;   // count mode: the pivot
    float* pv;                   // top-k partial values  [parts][R][k]      | CE: partial row max    [parts][R]
    int* pi;                     // top-k partial ids     [parts][R][k]
    int* pg;                     // partial #greater      [parts][R]
    int* pt;                     // partial #tie-lower    [parts][R]
    float* ps;                   // CE: partial sum-exp   [parts][R]
    float* captured;             // [R]: score of the target column as computed by this kernel (owner writes)
    const int32_t* n_live;       // device count of live rows (row selections of capacity R) or NULL: row tiles past it exit at once
    int seq_parts;               // 0 / 1: one part per CTA.  P > 1 (round-robin tiles, top-k only): the CTA sweeps the parts split*P .. split*P+P-1
                                 // one after the other -- part g = tiles g, g + tile_step, g + 2 tile_step, ... -- and every epilogue THREAD
                                 // writes its own list at the end of each part (slot (g * WGS + warpgroup): no fold), keeps the list's tail as
                                 // its threshold and starts the next part with an empty list.  Many row tiles (a rank of the vocab-sharded
                                 // evaluation scores ALL users against its slice) then need neither 16 short-lived CTAs per row tile (7 waves,
                                 // each paying the insertion storm of an unseeded list) nor a threshold pass.
    int unfolded;                // 1: every epilogue thread writes its own list (slot (split * WGS + warpgroup)) although the CTA sweeps a single
                                 // part: 4 lists per CTA, so 5..15 natural catalog splits already give the union its 16+ lists and the catalog
                                 // need not be cut into 16 splits (256 CTAs = 1.73 waves at 2048 rows)
    int tile_step;               // 0 / 1: a split sweeps a contiguous range of tiles; S > 1: split s sweeps tiles s, s+S, s+2S, ... (round
                                 // robin: whatever order the catalog is in, every split sees an even share of the best items)
};

struct __align__(8) ScoreTcBarriers {
    uint64_t a_full;
    uint64_t full[MAX_STAGES];
    uint64_t empty[MAX_STAGES];
    uint64_t tfull[2];
    uint64_t tempty[2];
    uint32_t tmem_base;
};

// Sorted insertion into this thread's REGISTER-resident list (descending; equal scores keep stream order, and a thread
// sees its columns in ascending id order, so ties resolve to the lowest item id).  Branch-free and fully unrolled: an
// insertion is ~6 instructions per entry with no memory latency (the first version kept the lists in shared memory and
// was latency-bound: ~800 cycles per insertion).
template <int KC>
__device__ __forceinline__ void reg_insert(float (&lv)[KC], int (&li)[KC], float x, int id) {
#pragma unroll
    for (int p = KC - 1; p > 0; --p) {
        const bool shift = lv[p - 1] < x;            // the entry above moves down into slot p
        const bool here = !shift && (lv[p] < x);     // x lands in slot p
        lv[p] = shift ? lv[p - 1] : (here ? x : lv[p]);
        li[p] = shift ? li[p - 1] : (here ? id : li[p]);
    }
    const bool here0 = lv[0] < x;
    lv[0] = here0 ? x : lv[0];
    li[0] = here0 ? id : li[0];
}

__device__ __forceinline__ bool tc_better(float v, int id, float v2, int id2) { return v > v2 || (v == v2 && id < id2); }

// insertion with the full order (score desc, id asc): parts arrive in arbitrary id order
template <int KC>
__device__ __forceinline__ void tie_insert(float (&lv)[KC], int (&li)[KC], float x, int id) {
#pragma unroll
    for (int p = KC - 1; p > 0; --p) {
        const bool shift = tc_better(x, id, lv[p - 1], li[p - 1]);
        const bool here = !shift && tc_better(x, id, lv[p], li[p]);
        lv[p] = shift ? lv[p - 1] : (here ? x : lv[p]);
        li[p] = shift ? li[p - 1] : (here ? id : li[p]);
    }
    const bool here0 = tc_better(x, id, lv[0], li[0]);
    lv[0] = here0 ? x : lv[0];
    li[0] = here0 ? id : li[0];
}

// Deferred insertion.  A candidate (score above the thread's threshold) is rare once the thresholds have converged, but
// handling it on the spot costs the whole warp ~1000 cycles with one active lane.  Candidates are therefore only PUSHED
// (8 bytes into a per-thread shared-memory FIFO) where they are found and INSERTED later in batches: the warp drains all
// its FIFOs together when any of them may overflow and once at the end, so the serial insertion code runs with many
// lanes active.  FIFO order == stream order == ascending item id, so ties still resolve to the lowest id.
#define PEND_CAP_MAX 16         // FIFO entries per epilogue thread (run-time a.pend_cap <= this); drained when more than pend_cap - 8 are pending
template <int KC>
__device__ __forceinline__ void pend_drain(float (&lv)[KC], int (&li)[KC], float& thr, float thr0, int& cnt, const uint2* pend,
                                           int stride) {
    for (int e = 0; e < cnt; ++e) {
        const uint2 u = pend[(size_t)e * stride];
        const float x = __uint_as_float(u.x);
        if (x > thr) {
            reg_insert<KC>(lv, li, x, (int)u.y);
            thr = fmaxf(thr0, lv[KC - 1]);
        }
    }
    cnt = 0;
}

// v[c] for a run-time c in 0..31 without dynamic register indexing: 5-level select tree (31 SEL)
__device__ __forceinline__ float sel32(const float (&v)[32], int c) {
    float a[16], b[8], d[4], e[2];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (c & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (c & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i] = (c & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
    for (int i = 0; i < 2; ++i) e[i] = (c & 8) ? d[2 * i + 1] : d[2 * i];
    return (c & 16) ? e[1] : e[0];
}

__device__ __forceinline__ float max8(const float* v) {
    return fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
}

// K-major operand chunk of 16 columns staged with the 32-byte swizzle: rows of 32 bytes, 8-row groups 256 bytes apart
__device__ __forceinline__ uint64_t smem_desc_sw32(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;                              // layout type: SWIZZLE_32B
    return d;
}

// KC = capacity of the per-thread top-k list (0: no top-k); the first a.k (<= KC) entries are reported
// PAIR: the CTA and its cluster neighbour (the next row tile) form a cta_group::2 pair -- one M=256 MMA per K-step issued by the
// even CTA, each CTA staging its own 128 rows of A and its own 128 of the tile's 256 table rows, so that shared memory
// moves every table byte once per pair instead of once per CTA (a single CTA is shared-memory-bandwidth bound: per tile it
// writes 64 KB and reads 96 KB of operands in the 1024 cycles the MMAs need).
template <int EPI, int KC, bool COUNT, int WGS, bool PAIR>
__global__ void __launch_bounds__(TC_THREADS(WGS), 1) score_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmB,
                                                                  const __grid_constant__ CUtensorMap tmAt,
                                                                  const __grid_constant__ CUtensorMap tmBt, const ScoreTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u)   /* pointer arithmetic on the __shared__ array keeps the address space: LDS/STS, not generic LD/ST */;
    const int kch = a.kch, stages = a.stages;
    uint8_t* sA = smem;
    uint8_t* sB = sA + (size_t)kch * A_CHUNK_BYTES;
    // `stages` ring slots of ONE 64-wide K chunk each
    constexpr uint32_t B_SLOT_BYTES = PAIR ? B_CHUNK_BYTES / 2 : B_CHUNK_BYTES;   // table rows of one K chunk staged by this CTA
    ScoreTcBarriers* bars = reinterpret_cast<ScoreTcBarriers*>(sB + (size_t)stages * B_SLOT_BYTES);
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    constexpr bool TOPK = KC > 0;
    constexpr int KL = KC > 0 ? KC : 1;
    constexpr int EPI_COLS = 256 / WGS;
    // top-k only: FIFO of pending candidates, entry e of epilogue thread i at pend[e * (WGS * 128) + i]
    uint2* pend_base = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(bars) + ((sizeof(ScoreTcBarriers) + 15) & ~(size_t)15));

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int m0 = blockIdx.x * BM;
    const int split = blockIdx.y;
    const int tstep = a.tile_step > 1 ? a.tile_step : 1;
    const int nseq = a.seq_parts > 1 ? a.seq_parts : 1;          // parts this CTA sweeps one after the other
    const bool unfolded = nseq > 1 || a.unfolded != 0;           // per-thread lists go out as they are (top-k sweeps only)
    // first / end tile of sequential part pp (pp = 0 when the CTA has one part)
    auto part_t0 = [&](int pp) {
        return tstep > 1 ? min(a.n_tiles, split * nseq + pp + a.tile_lo * tstep) : min(a.n_tiles, split * a.tiles_per_split + a.tile_lo);
    };
    auto part_t1 = [&](int pp) {
        return tstep > 1 ? min(a.n_tiles, split * nseq + pp + a.tile_hi * tstep) : min(a.n_tiles, split * a.tiles_per_split + a.tile_hi);
    };
    if (a.n_live != nullptr && m0 >= __ldg(a.n_live)) return;      // whole CTA, before any barrier / TMEM allocation (never with PAIR)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(&bars->a_full, 1);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->tfull[s], 1);
            mbar_init(&bars->tempty[s], (PAIR ? 2 : 1) * 4 * WGS);
        }
        fence_barrier_init();
    }
    __syncwarp();
    if (PAIR) cluster_sync_all();      // the neighbour's barriers exist before anything signals them
    if (warp == 2) {
        if (PAIR) tmem_alloc_pair(&bars->tmem_base, 512);
        else tmem_alloc(&bars->tmem_base, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    // programmatic dependent launch: the next kernel of the stream (the merge) may be scheduled now; it waits for this grid's
    // completion itself before it reads anything
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            // pair: the leader's barriers count the bytes of both CTAs
            const uint32_t a_tail_bytes = a.tail16 ? (uint32_t)BM * 32u : (uint32_t)A_CHUNK_BYTES;
            if (leader) mbar_arrive_expect_tx(&bars->a_full, (PAIR ? 2u : 1u) * ((uint32_t)(kch - 1) * A_CHUNK_BYTES + a_tail_bytes));
            for (int c = 0; c < kch; ++c) {
                if (PAIR) tma_load_2d_pair(sA + (size_t)c * A_CHUNK_BYTES, &tmA, &bars->a_full, c * CHUNK_K, m0);
                else if (a.tail16 && c == kch - 1) tma_load_2d(sA + (size_t)c * A_CHUNK_BYTES, &tmAt, &bars->a_full, c * CHUNK_K, m0);
                else tma_load_2d(sA + (size_t)c * A_CHUNK_BYTES, &tmA, &bars->a_full, c * CHUNK_K, m0);
            }
            int s = 0;                                    // ring slot and its phase, advanced without integer division
            uint32_t ph = 0;
            for (int pp = 0; pp < nseq; ++pp)
            for (int t = part_t0(pp), t1 = part_t1(pp); t < t1; t += tstep) {
                for (int c = 0; c < kch; ++c) {
                    mbar_wait(&bars->empty[s], ph ^ 1u);
                    const bool tail = a.tail16 && c == kch - 1;
                    if (leader) mbar_arrive_expect_tx(&bars->full[s], tail ? (uint32_t)BN * 32u : (uint32_t)B_CHUNK_BYTES);
                    if (PAIR) tma_load_2d_pair(sB + (size_t)s * B_SLOT_BYTES, &tmB, &bars->full[s], c * CHUNK_K, t * BN + (int)rank * (BN / 2));
                    else if (tail) tma_load_2d(sB + (size_t)s * B_SLOT_BYTES, &tmBt, &bars->full[s], c * CHUNK_K, t * BN);
                    else tma_load_2d(sB + (size_t)s * B_SLOT_BYTES, &tmB, &bars->full[s], c * CHUNK_K, t * BN);
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =====================================
        if (lane == 0 && leader) {
            const uint32_t idesc = idesc_bf16_f32(PAIR ? 2 * BM : BM, BN);
            mbar_wait(&bars->a_full, 0);
            tc_fence_after();
            // This single thread paces the tensor core: one 128x256x16 MMA is 128 cycles, so a whole tile (kch*4 MMAs) must be
            // issued in well under kch*512 cycles.  Descriptors are therefore built once and advanced by adds, the ring slot /
            // phase are tracked without integer division, and the four MMAs of a chunk are issued back to back.
            const uint64_t adesc0 = smem_desc_sw128(smem_u32(sA));
            const uint64_t bdesc0 = smem_desc_sw128(smem_u32(sB));
            int i = 0, s = 0;
            uint32_t ph = 0;
            for (int pp = 0; pp < nseq; ++pp)
            for (int t = part_t0(pp), t1 = part_t1(pp); t < t1; t += tstep, ++i) {
                const int as = i & 1;
                const uint32_t aph = (uint32_t)(i >> 1) & 1u;
                mbar_wait(&bars->tempty[as], aph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)as * BN;
                for (int c = 0; c < kch; ++c) {
                    mbar_wait(&bars->full[s], ph);
                    tc_fence_after();
                    const uint64_t ad = adesc0 + (uint64_t)(c * (A_CHUNK_BYTES >> 4));
                    const uint64_t bd = bdesc0 + (uint64_t)(s * (B_SLOT_BYTES >> 4));
                    const int ksteps = (c < kch - 1) ? 4 : a.last_ksteps;
                    if (PAIR) {
                        for (int k4 = 0; k4 < ksteps; ++k4) umma_bf16_pair(d_tmem, ad + 2 * k4, bd + 2 * k4, idesc, (uint32_t)((c | k4) != 0));
                        umma_commit_pair(&bars->empty[s]);
                    } else if (a.tail16 && c == kch - 1) {
                        umma_bf16(d_tmem, smem_desc_sw32(smem_u32(sA + (size_t)c * A_CHUNK_BYTES)),
                                  smem_desc_sw32(smem_u32(sB + (size_t)s * B_SLOT_BYTES)), idesc, (uint32_t)(c != 0));
                        umma_commit(&bars->empty[s]);
                    } else {
                        if (ksteps == 4) {
                            umma_bf16(d_tmem, ad, bd, idesc, (uint32_t)(c != 0));
                            umma_bf16(d_tmem, ad + 2, bd + 2, idesc, 1u);
                            umma_bf16(d_tmem, ad + 4, bd + 4, idesc, 1u);
                            umma_bf16(d_tmem, ad + 6, bd + 6, idesc, 1u);
                        } else {
                            for (int k4 = 0; k4 < ksteps; ++k4) umma_bf16(d_tmem, ad + 2 * k4, bd + 2 * k4, idesc, (uint32_t)((c | k4) != 0));
                        }
                        umma_commit(&bars->empty[s]);     // the slot may be refilled once these MMAs have read it
                    }
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
                if (PAIR) umma_commit_pair(&bars->tfull[as]);
                else umma_commit(&bars->tfull[as]);        // accumulator ready for the epilogue
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===================================== epilogue =====================================
        const int wg = (warp - EPI_WARP0) / 4;       // which EPI_COLS-column slice of the tile
        const int q = warp % 4;                      // TMEM lane quarter this warp may access
        const int tid = q * 32 + lane;               // row within the M tile == TMEM lane
        const int row = m0 + tid;
        const bool row_ok = row < a.R;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);

        long long tgl_ll = -1;
        if (a.target && row_ok) tgl_ll = a.target[row] - (long long)a.v0;
        // local target column; targets owned by a lower / higher shard become -1 / INT_MAX (ties: "id < target")
        const int tgl = tgl_ll < 0 ? -1 : (tgl_ll >= (long long)a.Vloc ? INT_MAX : (int)tgl_ll);

        // --- per-thread state ---
        float lv[KL];
        int li[KL];
#pragma unroll
        for (int p = 0; p < KL; ++p) {
            lv[p] = -INFINITY;
            li[p] = INT_MAX;
        }
        // threshold: nothing at or below it can enter the list.  thr0 comes from a sample sweep (strictly below the
        // sample's k-th best, so ties with it still pass the strict compare); once the own list is full its tail takes over.
        float thr0 = -INFINITY;
        if (TOPK && a.thr_init) asm volatile("griddepcontrol.wait;" ::: "memory");      // thresholds come from the previous launch
        if (TOPK && a.thr_init && row_ok) {
            const float t_row = a.thr_init[(size_t)row * a.thr_stride + a.thr_col];
            thr0 = t_row == -INFINITY ? t_row : nextafterf(t_row, -INFINITY);
        }
        thr0 = fmaxf(thr0, a.thr_floor);
        float thr = thr0;
        uint2* pend = pend_base + ((warp - EPI_WARP0) * 32 + lane);
        int cnt = 0;
        int cg = 0, ct = 0;
        float st = INFINITY;
        if (EPI == EPI_TOPK && COUNT && row_ok) st = a.target_score[row];
        float run_m = -INFINITY, run_s = 0.f;     // CE: running max (in log2 units) and sum of exp2

        int i = 0;
        // bias bounds of this warp's two chunks of the current / next tile (BOUND tiles, four warpgroups)
        float2 bb_cur0 = make_float2(0.f, 0.f), bb_cur1 = bb_cur0;
        const bool use_bb = EPI == EPI_TOPK && TOPK && !COUNT && EPI_COLS == 64 && a.bias_bounds != nullptr;
        const int bb_last = ((a.Vloc + 31) >> 5) - 1;
        auto load_bb = [&](int tt, float2& b0, float2& b1) {
            const int c = (tt * BN + wg * EPI_COLS) >> 5;
            b0 = __ldg(a.bias_bounds + min(c, bb_last));
            b1 = __ldg(a.bias_bounds + min(c + 1, bb_last));
        };
      for (int pp = 0; pp < nseq; ++pp) {
        const int t0 = part_t0(pp), t1 = part_t1(pp);
        if (use_bb && t0 < t1) load_bb(t0, bb_cur0, bb_cur1);
        for (int t = t0; t < t1; t += tstep, ++i) {
            const int as = i & 1;
            const uint32_t aph = (uint32_t)(i >> 1) & 1u;
            float2 bb_nxt0 = bb_cur0, bb_nxt1 = bb_cur1;
            if (use_bb && t + tstep < t1) load_bb(t + tstep, bb_nxt0, bb_nxt1);
            mbar_wait_lean(&bars->tfull[as], aph);
            tc_fence_after();
            if (EPI == EPI_PROBE) {          // a.k == 0: no TMEM reads at all (what TMA + MMA + handshakes sustain);
                if (a.k != 0) {              // a.k != 0: read the whole accumulator and discard it (TMEM read throughput)
                    float v[32];
#pragma unroll 1
                    for (int ch = 0; ch < EPI_COLS / 32; ++ch) {
                        tmem_ld32(lane_addr + (uint32_t)(as * BN + wg * EPI_COLS + ch * 32), v);
                        tmem_ld_wait();
                    }
                    if (v[0] == 12345.678f) a.captured[0] = v[1];      // keep the loads alive (a.captured is NULL in the probe)
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                        if (PAIR) mbar_arrive_leader(&bars->tempty[as]);
                        else mbar_arrive(&bars->tempty[as]);
                    }
                continue;
            }
            // PLAIN tiles (all 256 columns valid, no bias vector, no target column of this warp's rows inside) skip the per-chunk
            // bookkeeping: the epilogue warps are issue-bound, every instruction per chunk counts
            // MODE 0: general tile; 1: PLAIN; 2: BOUND = a full tile without target column whose bias is only bounded per chunk
            auto chunk = [&](auto mode_c, const int ch) {
                constexpr int MODE = decltype(mode_c)::value;
                constexpr bool PLAIN = MODE != 0;
                constexpr bool BOUND = MODE == 2;
                const int col0 = t * BN + wg * EPI_COLS + ch * 32;     // local column of v[0]
                float2 bb = make_float2(0.f, 0.f);
                if (BOUND) {      // warp-uniform; the array (V/4 bytes) does not stay in L1 next to 227 KB of shared memory, so the
                                  // bounds of a tile were requested a whole tile ahead (two chunks per warp with four warpgroups)
                    if (EPI_COLS == 64) bb = ch == 0 ? bb_cur0 : bb_cur1;
                    else bb = __ldg(a.bias_bounds + (col0 >> 5));
                }
                float v[32];
                tmem_ld32(lane_addr + (uint32_t)(as * BN + wg * EPI_COLS + ch * 32), v);
                tmem_ld_wait();
                if (ch == EPI_COLS / 32 - 1) {   // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (PAIR) mbar_arrive_leader(&bars->tempty[as]);
                        else mbar_arrive(&bars->tempty[as]);
                    }
                }
                if (!PLAIN && col0 >= a.Vloc) return;           // warp-uniform
                const bool full_valid = PLAIN || col0 + 32 <= a.Vloc;
                if (!PLAIN && a.bias) {
                    if (full_valid) {
#pragma unroll
                        for (int c = 0; c < 32; c += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + col0 + c));
                            v[c] += b.x; v[c + 1] += b.y; v[c + 2] += b.z; v[c + 3] += b.w;
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            if (col0 + c < a.Vloc) v[c] += __ldg(a.bias + col0 + c);
                    }
                }
                if (!full_valid) {
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (col0 + c >= a.Vloc) v[c] = -INFINITY;
                }
                // score of the target column exactly as this kernel computes it
                if (!PLAIN && a.captured && tgl >= col0 && tgl < col0 + 32) {
                    float x = 0.f;
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (col0 + c == tgl) x = v[c];
                    a.captured[row] = x;
                }
                if (EPI == EPI_CE) {
                    float m8[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) m8[j] = max8(v + 8 * j);
                    const float mc = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])) * L2E;
                    const float m_new = fmaxf(run_m, mc);
                    float s = 0.f;
#pragma unroll
                    for (int c = 0; c < 32; ++c) s += exp2f(fmaf(v[c], L2E, -m_new));
                    run_s = run_s * exp2f(run_m - m_new) + s;
                    run_m = m_new;
                } else {
                    if (COUNT) {
                        bool any_eq = false;
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            cg += v[c] > st;
                            any_eq |= v[c] == st;
                        }
                        if (any_eq) {
#pragma unroll
                            for (int c = 0; c < 32; ++c) ct += (v[c] == st) && (col0 + c < tgl);
                        }
                    }
                    if (TOPK) {
                        float m8[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) m8[j] = max8(v + 8 * j);
                        float mc = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
                        // BOUND: rounding is monotone, so fl(best raw + max bias) >= every biased score of the chunk (main sweep: nothing
                        // that could pass the threshold is skipped) and fl(best raw + min bias) <= the chunk's best biased score (threshold
                        // pass: k distinct chunks hold a score >= the k-th best of these lower bounds)
                        if (BOUND) mc += a.sample_mode ? bb.y : bb.x;
                        if (__any_sync(0xffffffffu, mc > thr)) {      // rare once the thresholds have converged
                            if (BOUND && !a.sample_mode) {            // now the exact biased scores of this chunk
#pragma unroll
                                for (int c = 0; c < 32; c += 4) {
                                    const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + col0 + c));
                                    v[c] += b.x; v[c + 1] += b.y; v[c + 2] += b.z; v[c + 3] += b.w;
                                }
#pragma unroll
                                for (int j = 0; j < 4; ++j) m8[j] = max8(v + 8 * j);
                            }
                            if (a.sample_mode) {      // warp-uniform
                                if (__any_sync(0xffffffffu, cnt > a.pend_cap - 8)) pend_drain<KL>(lv, li, thr, thr0, cnt, pend, WGS * 128);
                                if (mc > thr) {
                                    pend[(size_t)cnt * (WGS * 128)] = make_uint2(__float_as_uint(mc), (uint32_t)(a.v0 + col0));
                                    ++cnt;
                                }
                            } else
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (__any_sync(0xffffffffu, cnt > a.pend_cap - 8)) pend_drain<KL>(lv, li, thr, thr0, cnt, pend, WGS * 128);
                                if (m8[j] > thr) {
#pragma unroll
                                    for (int c = 0; c < 8; ++c) {
                                        const float x = v[8 * j + c];
                                        if (x > thr) {
                                            pend[(size_t)cnt * (WGS * 128)] = make_uint2(__float_as_uint(x), (uint32_t)(a.v0 + col0 + 8 * j + c));
                                            ++cnt;
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
            };
            const bool tile_full = (t + 1) * BN <= a.Vloc;
            const bool target_here = a.captured && tgl >= t * BN && tgl < (t + 1) * BN;
            const bool clean = tile_full && !__any_sync(0xffffffffu, target_here);
            if (clean && !a.bias) {
#pragma unroll 1
                for (int ch = 0; ch < EPI_COLS / 32; ++ch) chunk(std::integral_constant<int, 1>{}, ch);
            } else if (EPI == EPI_TOPK && TOPK && !COUNT && clean && a.bias_bounds) {
#pragma unroll 1
                for (int ch = 0; ch < EPI_COLS / 32; ++ch) chunk(std::integral_constant<int, 2>{}, ch);
            } else {
#pragma unroll 1
                for (int ch = 0; ch < EPI_COLS / 32; ++ch) chunk(std::integral_constant<int, 0>{}, ch);
            }
            bb_cur0 = bb_nxt0; bb_cur1 = bb_nxt1;
        }
        if (TOPK && unfolded) {       // end of a (sequential) part: this thread's own list goes out, its tail seeds the next part
            pend_drain<KL>(lv, li, thr, thr0, cnt, pend, WGS * 128);
            if (row_ok) {
                const size_t o = ((size_t)a.part0 + (size_t)(split * nseq + pp) * WGS + wg) * a.R + row;
#pragma unroll
                for (int p = 0; p < KL; ++p) {
                    if (p < a.k) {
                        a.pv[o * a.k + p] = lv[p];
                        a.pi[o * a.k + p] = li[p];
                    }
                }
            }
            // (strictly below the tail: an equal score with a lower item id still has to pass the strict compare)
            if (li[KL - 1] != INT_MAX) thr0 = fmaxf(thr0, nextafterf(lv[KL - 1], -INFINITY));
            thr = thr0;
#pragma unroll
            for (int p = 0; p < KL; ++p) { lv[p] = -INFINITY; li[p] = INT_MAX; }
        }
      }
        if (TOPK) pend_drain<KL>(lv, li, thr, thr0, cnt, pend, WGS * 128);
        // --- fold the warpgroups' results into warpgroup 0 through the table ring (idle: every MMA of this CTA has completed),
        //     so that the CTA emits ONE partial result per row
        if (EPI != EPI_PROBE && !(TOPK && unfolded)) {
            constexpr int SLOTS = EPI == EPI_CE ? 2 : (2 * KL + 2);       // 32-bit words staged per row and warpgroup
            uint32_t* xs = reinterpret_cast<uint32_t*>(sB);               // [WGS-1][SLOTS][128]
            if (wg > 0) {
                uint32_t* mine = xs + (size_t)(wg - 1) * SLOTS * 128 + tid;
                if (EPI == EPI_CE) {
                    mine[0] = __float_as_uint(run_m);
                    mine[128] = __float_as_uint(run_s);
                } else {
#pragma unroll
                    for (int p = 0; p < KL; ++p) {
                        mine[(size_t)(2 * p) * 128] = __float_as_uint(lv[p]);
                        mine[(size_t)(2 * p + 1) * 128] = (uint32_t)li[p];
                    }
                    mine[(size_t)(2 * KL) * 128] = (uint32_t)cg;
                    mine[(size_t)(2 * KL + 1) * 128] = (uint32_t)ct;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(128 * WGS) : "memory");      // the epilogue warps only
            if (wg == 0) {
#pragma unroll 1
                for (int w = 0; w < WGS - 1; ++w) {
                    const uint32_t* other = xs + (size_t)w * SLOTS * 128 + tid;
                    if (EPI == EPI_CE) {
                        const float m2 = __uint_as_float(other[0]), s2 = __uint_as_float(other[128]);
                        const float m_new = fmaxf(run_m, m2);
                        if (m_new > -INFINITY) run_s = run_s * exp2f(run_m - m_new) + s2 * exp2f(m2 - m_new);
                        run_m = m_new;
                    } else {
                        if (TOPK) {
#pragma unroll 1
                            for (int p = 0; p < KL; ++p) {                 // sorted: stop at the first entry that cannot enter
                                const float x = __uint_as_float(other[(size_t)(2 * p) * 128]);
                                const int id = (int)other[(size_t)(2 * p + 1) * 128];
                                if (id == INT_MAX || !tc_better(x, id, lv[KL - 1], li[KL - 1])) break;
                                tie_insert<KL>(lv, li, x, id);
                            }
                        }
                        if (COUNT) {
                            cg += (int)other[(size_t)(2 * KL) * 128];
                            ct += (int)other[(size_t)(2 * KL + 1) * 128];
                        }
                    }
                }
            }
        }
        // --- write this split's partial results ---
        if (row_ok && wg == 0 && !(TOPK && unfolded)) {
            const size_t part = (size_t)a.part0 + (size_t)split;
            const size_t o = part * a.R + row;
            if (EPI == EPI_CE) {
                a.pv[o] = run_m;      // log2 units
                a.ps[o] = run_s;
            } else {
                if (TOPK) {
#pragma unroll
                    for (int p = 0; p < KL; ++p) {
                        if (p < a.k) {
                            a.pv[o * a.k + p] = lv[p];
                            a.pi[o * a.k + p] = li[p];
                        }
                    }
                }
                if (COUNT) {
                    a.pg[o] = cg;
                    a.pt[o] = ct;
                }
            }
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();      // neither CTA frees tensor memory (or exits) while the pair's MMAs / signals may touch it
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------------------
// merge of the partial results: one warp per row
// ------------------------------------------------------------------------------------------------------------
// One warp per row.  Phase 1: lane l folds parts l, l+32, ... into a register list (all loads of a part are independent and
// the lanes work on different parts, so the kernel pays a handful of memory latencies instead of one per part).  Phase 2:
// k rounds of warp arg-max over the list heads; the winning lane pops its head.
template <int KC>
__global__ void __launch_bounds__(128) tc_topk_merge_kernel(const float* __restrict__ pv, const int* __restrict__ pi,
                                                            const int* __restrict__ pg, const int* __restrict__ pt, int parts, int R,
                                                            int k, float* __restrict__ out_v, int32_t* __restrict__ out_i,
                                                            int32_t* __restrict__ out_g, int32_t* __restrict__ out_t) {
    const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    asm volatile("griddepcontrol.wait;" ::: "memory");                 // the partial lists come from the previous launch
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (row >= R) return;
    if (KC > 0 && pv) {
        constexpr int KL = KC > 0 ? KC : 1;
        float lv[KL];
        int li[KL];
#pragma unroll
        for (int p = 0; p < KL; ++p) { lv[p] = -INFINITY; li[p] = INT_MAX; }
        for (int p = lane; p < parts; p += 32) {
            const size_t o = ((size_t)p * R + row) * k;
            float cv[KL];
            int ci[KL];
#pragma unroll
            for (int j = 0; j < KL; ++j) {
                cv[j] = j < k ? __ldg(pv + o + j) : -INFINITY;
                ci[j] = j < k ? __ldg(pi + o + j) : INT_MAX;
            }
#pragma unroll
            for (int j = 0; j < KL; ++j) {
                // a part is sorted: once an entry is empty or cannot enter, neither can the rest
                if (ci[j] == INT_MAX || !tc_better(cv[j], ci[j], lv[KL - 1], li[KL - 1])) break;
                tie_insert<KL>(lv, li, cv[j], ci[j]);
            }
        }
        float ov = -INFINITY;
        int oi = INT_MAX;
        for (int r = 0; r < k; ++r) {
            float bv = lv[0];
            int bi = li[0];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const float v2 = __shfl_xor_sync(0xffffffffu, bv, s);
                const int i2 = __shfl_xor_sync(0xffffffffu, bi, s);
                if (tc_better(v2, i2, bv, bi)) { bv = v2; bi = i2; }
            }
            if (bi != INT_MAX && li[0] == bi) {      // item ids are unique across parts: exactly one lane pops
#pragma unroll
                for (int p = 0; p < KL - 1; ++p) { lv[p] = lv[p + 1]; li[p] = li[p + 1]; }
                lv[KL - 1] = -INFINITY;
                li[KL - 1] = INT_MAX;
            }
            if (lane == r) { ov = bv; oi = bi; }
        }
        if (lane < k) {
            out_v[(size_t)row * k + lane] = ov;
            out_i[(size_t)row * k + lane] = oi == INT_MAX ? -1 : oi;
        }
    }
    if (pg) {
        int g = 0, t = 0;
        for (int p = lane; p < parts; p += 32) {
            g += pg[(size_t)p * R + row];
            t += pt[(size_t)p * R + row];
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            g += __shfl_xor_sync(0xffffffffu, g, s);
            t += __shfl_xor_sync(0xffffffffu, t, s);
        }
        if (lane == 0) { out_g[row] = g; out_t[row] = t; }
    }
}

extern int g_pdl_merge;
template <int KC>
static cudaError_t launch_merge_one(const float* pv, const int* pi, const int* pg, const int* pt, int parts, int R, int k, float* out_v,
                                    int32_t* out_i, int32_t* out_g, int32_t* out_t, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ceil_div(R, 4));
    cfg.blockDim = dim3(128);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl_merge ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, tc_topk_merge_kernel<KC>, pv, pi, pg, pt, parts, R, k, out_v, out_i, out_g, out_t);
}

static int launch_merge(const float* pv, const int* pi, const int* pg, const int* pt, int parts, int R, int k, float* out_v,
                        int32_t* out_i, int32_t* out_g, int32_t* out_t, cudaStream_t st) {
#define MERGE(KC) ASME_CUDA_OK(launch_merge_one<KC>(pv, pi, pg, pt, parts, R, k, out_v, out_i, out_g, out_t, st))
    if (!pv) MERGE(0);
    else if (k == 1) MERGE(1);
    else if (k <= 5) MERGE(5);
    else if (k <= 10) MERGE(10);
    else if (k <= 20) MERGE(20);
    else MERGE(32);
#undef MERGE
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// Candidate merge (exact top-k, csrc/rescore.cu): the union of the per-part lists is merged into the k_out best by bf16 score, and
// ``bound[row]`` = an upper bound of the bf16 score of EVERY item that is in none of the row's part lists.  Such an item was dropped
// by a sorted list that was full (its score is <= that list's last entry; the last entry of a list only grows), or it never passed
// the sample threshold: bound = max(last entry of every full part list, last entry of a full lane list that folded several parts,
// threshold).  Entries cut by the k_out limit are bounded by the k_out-th output itself (the caller adds that).
template <int KC>
__global__ void __launch_bounds__(128) tc_cand_merge_kernel(const float* __restrict__ pv, const int* __restrict__ pi, int parts, int R, int k,
                                                            int k_out, const float* __restrict__ thr_init, int thr_stride, int thr_col,
                                                            float* __restrict__ out_v, int32_t* __restrict__ out_i,
                                                            float* __restrict__ bound) {
    const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= R) return;
    constexpr int KL = KC;
    float lv[KL];
    int li[KL];
#pragma unroll
    for (int p = 0; p < KL; ++p) { lv[p] = -INFINITY; li[p] = INT_MAX; }
    float tau = -INFINITY;
    int folded = 0;
    for (int p = lane; p < parts; p += 32, ++folded) {
        const size_t o = ((size_t)p * R + row) * k;
        float cv[KL];
        int ci[KL];
#pragma unroll
        for (int j = 0; j < KL; ++j) {
            cv[j] = j < k ? __ldg(pv + o + j) : -INFINITY;
            ci[j] = j < k ? __ldg(pi + o + j) : INT_MAX;
        }
#pragma unroll
        for (int j = 0; j < KL; ++j)
            if (j == k - 1 && ci[j] != INT_MAX) tau = fmaxf(tau, cv[j]);          // a full part list: whatever it dropped is <= its tail
#pragma unroll
        for (int j = 0; j < KL; ++j) {
            if (ci[j] == INT_MAX || !tc_better(cv[j], ci[j], lv[KL - 1], li[KL - 1])) break;
            tie_insert<KL>(lv, li, cv[j], ci[j]);
        }
    }
    if (folded > 1 && li[KL - 1] != INT_MAX) tau = fmaxf(tau, lv[KL - 1]);        // entries this lane's fold pushed out
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) tau = fmaxf(tau, __shfl_xor_sync(0xffffffffu, tau, s));
    if (thr_init != nullptr) tau = fmaxf(tau, thr_init[(size_t)row * thr_stride + thr_col]);
    for (int r = 0; r < k_out; ++r) {
        float bv = lv[0];
        int bi = li[0];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const float v2 = __shfl_xor_sync(0xffffffffu, bv, s);
            const int i2 = __shfl_xor_sync(0xffffffffu, bi, s);
            if (tc_better(v2, i2, bv, bi)) { bv = v2; bi = i2; }
        }
        if (bi != INT_MAX && li[0] == bi) {
#pragma unroll
            for (int p = 0; p < KL - 1; ++p) { lv[p] = lv[p + 1]; li[p] = li[p + 1]; }
            lv[KL - 1] = -INFINITY;
            li[KL - 1] = INT_MAX;
        }
        if (lane == 0) {
            out_v[(size_t)row * k_out + r] = bv;
            out_i[(size_t)row * k_out + r] = bi == INT_MAX ? -1 : bi;
        }
    }
    if (lane == 0) bound[row] = tau;
}
// classic candidate path: (R, kc) lists -> (R, k_out) rows with empty tail slots; bound = -inf (the list is the true bf16 top kc)
__global__ void spread_candidates_kernel(const float* __restrict__ src_v, const int32_t* __restrict__ src_i, int R, int kc, int k_out,
                                         float* __restrict__ dst_v, int32_t* __restrict__ dst_i, float* __restrict__ bound) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)R * k_out) return;
    const int r = (int)(i / k_out), c = (int)(i % k_out);
    if (src_v != dst_v) {
        dst_v[i] = c < kc ? src_v[(size_t)r * kc + c] : -INFINITY;
        dst_i[i] = c < kc ? src_i[(size_t)r * kc + c] : -1;
    }
    // a full kc-entry list cut the catalog: everything outside it scores at most its last entry; a shorter list holds everything
    if (c == 0) bound[r] = (kc < k_out && src_i[(size_t)r * kc + kc - 1] >= 0) ? src_v[(size_t)r * kc + kc - 1] : -INFINITY;
}

// CE partials: (max in log2 units, sum of exp2) per part -> natural-log row max and sum-exp
__global__ void tc_ce_merge_kernel(const float* __restrict__ pm, const float* __restrict__ ps, int parts, int R,
                                   float* __restrict__ row_max, float* __restrict__ row_sumexp, const int32_t* __restrict__ n_live) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= asme_live_rows(R, n_live)) return;
    float m = -INFINITY;
    for (int p = 0; p < parts; ++p) m = fmaxf(m, pm[(size_t)p * R + r]);
    float s = 0.f;
    for (int p = 0; p < parts; ++p) {
        const float mp = pm[(size_t)p * R + r];
        if (mp > -INFINITY) s += ps[(size_t)p * R + r] * exp2f(mp - m);
    }
    row_max[r] = m * 0.6931471805599453f;
    row_sumexp[r] = s;
}

// {max, min} of the bias over every 32-column chunk [32 c, 32 c + 32) of the (local) catalog slice: one warp per chunk
__global__ void __launch_bounds__(256) bias_chunk_bounds_kernel(const float* __restrict__ bias, int V, float2* __restrict__ bounds) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
    if (c * 32 >= V) return;
    const int j = c * 32 + lane;
    const float b = j < V ? __ldg(bias + j) : __ldg(bias + c * 32);
    float hi = b, lo = b;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    }
    if (lane == 0) bounds[c] = make_float2(hi, lo);
}
extern "C" int asme_b200_bias_chunk_bounds(const float* bias, int V, float* bounds, asme_stream_t stream) {
    ASME_REQUIRE(bias && bounds && V >= 1, "bias_chunk_bounds: bad argument");
    ASME_REQUIRE(((uintptr_t)bounds & 7) == 0, "bias_chunk_bounds: bounds must be 8-byte aligned");
    const int chunks = ceil_div(V, 32);
    bias_chunk_bounds_kernel<<<ceil_div(chunks, 8), 256, 0, (cudaStream_t)stream>>>(bias, V, reinterpret_cast<float2*>(bounds));
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ------------------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------------------
struct ScorePlan {
    int m_tiles, n_tiles, splits, tiles_per_split, parts, kch, stages, last_ksteps, wgs, pair, grid_x, tail16;
    size_t smem;
    CUtensorMap tmAt, tmBt;      // 16-column / 32-byte-swizzle maps of the K tail (tail16 only)
};

// Tuning knobs (diagnostics; defaults are what the product path uses).  Results do not depend on them.
static int g_epi_wgs = 4;          // epilogue warpgroups of the top-k sweeps: 4 (measured best; k > 10 needs the registers of 2)
static int g_epi_wgs_other = 4;    // ... of the CE / count-only sweeps (their epilogues are ALU-bound: 4 is faster)
static int g_sample_div = 16;      // the sample sweep scores 1/g_sample_div of every split's tiles (0: no sample sweep)
static float g_thr_floor = -INFINITY;
static int g_pair = 0;             // 1: CTA pairs (cta_group::2) whenever there are at least two row tiles; 0: single CTAs.  Measured:
                                   // no gain (the single-CTA pipeline already runs at the power-capped tensor peak, it is not shared-
                                   // memory bound) and the pair couples the two epilogues' jitter -> off by default, kept as a knob
static int g_tail16 = 1;           // stage a 16-column K tail (Kp = 64 k + 16: folded bias) as a 32-byte-swizzled quarter-size box
static int g_pend_cap = 12;        // candidate FIFO depth per epilogue thread: 8 bytes x 512 threads per entry come out of the B ring's
                                   // shared memory -- 12 instead of 16 buys the ring a fourth slot at Kp = 144 (three K chunks per tile)
static int g_cand_seq = 1;         // candidate sweeps over many row tiles: 1 = sequential parts per CTA with unfolded thread lists;
                                   // 0 = always cut the catalog into CAND_MIN_PARTS splits (the first scheme, kept for A/B)
static int g_pdl = 0;              // programmatic dependent launch between the launches of one top-k call
int g_pdl_merge = 0;            // measured: slower (early-scheduled dependents take SM resources from the sweep) -> off, kept as a knob

extern "C" int asme_b200_tc_score_tune(int knob, int value) {
    switch (knob) {
        case 0: ASME_REQUIRE(value == 2 || value == 4, "tc_score_tune: epilogue warpgroups must be 2 or 4"); g_epi_wgs = value; break;
        case 1: ASME_REQUIRE(value >= 0, "tc_score_tune: sample divisor must be >= 0"); g_sample_div = value; break;
        case 2: g_thr_floor = value ? INFINITY : -INFINITY; break;
        case 3: ASME_REQUIRE(value == 2 || value == 4, "tc_score_tune: epilogue warpgroups must be 2 or 4"); g_epi_wgs_other = value; break;
        case 4: g_pair = value ? 1 : 0; break;
        case 5: g_pdl = g_pdl_merge = value ? 1 : 0; break;
        case 6: g_tail16 = value ? 1 : 0; break;
        case 7: ASME_REQUIRE(value >= 8 && value <= PEND_CAP_MAX, "tc_score_tune: FIFO depth must be 8..16"); g_pend_cap = value; break;
        case 8: g_cand_seq = value ? 1 : 0; break;
        default: ASME_REQUIRE(false, "tc_score_tune: unknown knob %d", knob);
    }
    return ASME_OK;
}

// ``plan_rows`` (0 = R): how many of the R rows are expected to be live (row selections pass their capacity as R and the count on
// the device); only the split of the catalog over CTAs is chosen from it -- a wrong guess costs balance, never correctness.
static int make_plan(int R, int Kp, int Vloc, ScorePlan* p, bool topk = false, int plan_rows = 0) {
    ASME_REQUIRE(R >= 1 && Vloc >= 1, "tc score: bad shape R=%d Vloc=%d", R, Vloc);
    ASME_REQUIRE(Kp >= 16 && Kp <= 272 && Kp % 16 == 0, "tc score: padded hidden size %d unsupported (multiple of 16, <= 272)", Kp);
    p->kch = ceil_div(Kp, CHUNK_K);
    p->last_ksteps = (Kp % CHUNK_K) ? (Kp % CHUNK_K) / UMMA_K : CHUNK_K / UMMA_K;
    p->tail16 = 0;
    p->m_tiles = ceil_div(R, BM);
    p->n_tiles = ceil_div(Vloc, BN);
    p->pair = (g_pair && p->m_tiles >= 2) ? 1 : 0;
    p->grid_x = p->pair ? (p->m_tiles + 1) / 2 * 2 : p->m_tiles;      // an odd last row tile is paired with an empty one
    p->tail16 = (g_tail16 && !p->pair && Kp % CHUNK_K == 16 && Kp > 16) ? 1 : 0;
    const int fill_x = (plan_rows > 0 && plan_rows < R) ? ceil_div(plan_rows, BM) : p->grid_x;
    int splits = ASME_NUM_SMS / fill_x;
    if (splits < 1) splits = 1;
    if (splits > p->n_tiles) splits = p->n_tiles;
    p->tiles_per_split = ceil_div(p->n_tiles, splits);
    p->splits = ceil_div(p->n_tiles, p->tiles_per_split);
    p->wgs = topk ? g_epi_wgs : g_epi_wgs_other;
    p->parts = p->splits;      // the warpgroups of a CTA fold their results before writing
    const size_t fixed = 1024 + (size_t)p->kch * A_CHUNK_BYTES + ((sizeof(ScoreTcBarriers) + 15) & ~(size_t)15) +
                         (topk ? (size_t)p->wgs * 128 * g_pend_cap * sizeof(uint2) : 0);
    const size_t stage = (size_t)B_CHUNK_BYTES / (p->pair ? 2 : 1);
    const size_t budget = 227 * 1024;
    ASME_REQUIRE(fixed + 2 * stage <= budget, "tc score: shared memory budget exceeded (Kp=%d)", Kp);
    int stages = (int)((budget - fixed) / stage);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    p->stages = stages;
    p->smem = fixed + stages * stage;
    return ASME_OK;
}

// Threshold pass: with per-thread lists every row pays ~k*ln(n/k) insertions in EACH of its lists while the local thresholds
// converge.  For long sweeps a first launch scores 1/16 of every split's tiles and keeps only chunk maxima (one candidate per
// 32 columns: cheap); the merged k-th best chunk maximum becomes every list's initial threshold in the second launch, which
// sweeps everything (the 1/16 is scored twice) with rare insertions.  Results are identical either way; only the work differs.
static int sample_tiles(const ScorePlan& p) { return (g_sample_div > 0 && p.tiles_per_split >= 2 * g_sample_div) ? p.tiles_per_split / g_sample_div : 0; }

extern "C" size_t asme_b200_tc_score_topk_workspace_bytes(int R, int Kp, int Vloc, int k) {
    ScorePlan p;
    if (make_plan(R < 1 ? 1 : R, Kp, Vloc, &p)) return 0;
    return (size_t)2 * (p.splits + 1) * MAX_EPI_WGS * (R < 1 ? 1 : R) * ((size_t)k * 8 + 8);    // + 1: either pairing mode fits
}

template <typename K>
static int tc_set_smem(K kernel, size_t bytes) {
    { const int _rc = asme_ensure_max_smem((const void*)kernel); if (_rc) return _rc; }
    return ASME_OK;
}

static int make_tail_maps(ScorePlan* p, const void* Hb, int R, const void* Wb, int Vloc, int Kp) {
    if (!p->tail16) {
        memset(&p->tmAt, 0, sizeof(CUtensorMap));
        memset(&p->tmBt, 0, sizeof(CUtensorMap));
        return ASME_OK;
    }
    int rc = make_tmap_bf16_tail16(&p->tmAt, Hb, R, Kp, Kp, BM);
    if (rc) return rc;
    return make_tmap_bf16_tail16(&p->tmBt, Wb, Vloc, Kp, Kp, BN);
}

template <int EPI, int KC, bool COUNT, int WGS, bool PAIR>
static int launch_score_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const ScoreTcArgs& a, const ScorePlan& p, cudaStream_t st,
                            bool pdl) {
    auto kernel = score_tc_kernel<EPI, KC, COUNT, WGS, PAIR>;
    int rc = tc_set_smem(kernel, p.smem);
    if (rc) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.grid_x, p.splits);
    cfg.blockDim = dim3(TC_THREADS(WGS));
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 2 : 1;
    ASME_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, p.tmAt, p.tmBt, a));
    asme_count_launch();
    return ASME_OK;
}
template <int EPI, int KC, bool COUNT>
static int launch_score(const CUtensorMap& tmA, const CUtensorMap& tmB, const ScoreTcArgs& a, const ScorePlan& p, cudaStream_t st,
                        bool pdl = false) {
    if (p.wgs == 4)
        return p.pair ? launch_score_one<EPI, KC, COUNT, 4, true>(tmA, tmB, a, p, st, pdl)
                      : launch_score_one<EPI, KC, COUNT, 4, false>(tmA, tmB, a, p, st, pdl);
    return p.pair ? launch_score_one<EPI, KC, COUNT, 2, true>(tmA, tmB, a, p, st, pdl)
                  : launch_score_one<EPI, KC, COUNT, 2, false>(tmA, tmB, a, p, st, pdl);
}

static int launch_topk(const CUtensorMap& tmA, const CUtensorMap& tmB, const ScoreTcArgs& a, const ScorePlan& p, bool topk,
                       bool count, cudaStream_t st, bool pdl = false) {
#define LAUNCH_KC(KC) return count ? launch_score<EPI_TOPK, KC, true>(tmA, tmB, a, p, st, pdl) : launch_score<EPI_TOPK, KC, false>(tmA, tmB, a, p, st, pdl)
    if (!topk) return launch_score<EPI_TOPK, 0, true>(tmA, tmB, a, p, st, pdl);
    if (a.k == 1) LAUNCH_KC(1);
    if (a.k <= 5) LAUNCH_KC(5);
    if (a.k <= 10) LAUNCH_KC(10);
    if (a.k <= 20) LAUNCH_KC(20);
    LAUNCH_KC(32);
#undef LAUNCH_KC
}

extern "C" int asme_b200_tc_score_topk(const void* Hb, int R, int Kp, const void* Wb, const float* bias, int v0, int Vloc,
                                       const int64_t* target, const float* target_score_in, int k, float* topk_val,
                                       int32_t* topk_idx, float* target_score_out, int32_t* n_greater, int32_t* n_tie_lower,
                                       void* ws, size_t ws_bytes, asme_stream_t stream) {
    ASME_REQUIRE(Hb && Wb, "tc_score_topk: null operand");
    ASME_REQUIRE(k >= 0 && k <= 32, "tc_score_topk: k=%d unsupported (0..32)", k);
    const bool topk = k > 0, count = target_score_in != nullptr;
    ASME_REQUIRE(topk || count, "tc_score_topk: nothing to do (k = 0 and no target_score_in)");
    ASME_REQUIRE(!topk || (topk_val && topk_idx), "tc_score_topk: top-k outputs missing");
    ASME_REQUIRE(!count || (target && n_greater && n_tie_lower), "tc_score_topk: count mode needs target, n_greater, n_tie_lower");
    ASME_REQUIRE(!target_score_out || target, "tc_score_topk: target_score_out needs target");
    ASME_REQUIRE(!bias || ((uintptr_t)bias & 15) == 0, "tc_score_topk: bias must be 16-byte aligned");
    if (R == 0) return ASME_OK;
    ScorePlan p;
    int rc = make_plan(R, Kp, Vloc, &p, topk);
    if (rc) return rc;
    if (k > 10 && p.wgs != 2) {      // the 20- and 32-entry register lists do not fit 96 registers per thread
        const int keep = g_epi_wgs;
        g_epi_wgs = 2;
        rc = make_plan(R, Kp, Vloc, &p, topk);
        g_epi_wgs = keep;
        if (rc) return rc;
    }
    const size_t need = (size_t)2 * p.splits * MAX_EPI_WGS * R * ((size_t)k * 8 + 8);
    if (ws_bytes < need) {
        asme_set_error("tc_score_topk: workspace too small (%zu < %zu)", ws_bytes, need);
        return ASME_ERR_WORKSPACE;
    }
    CUtensorMap tmA, tmB;
    rc = asme_tc_make_tmap_bf16(&tmA, Hb, R, Kp, Kp, BM);
    if (rc) return rc;
    rc = asme_tc_make_tmap_bf16(&tmB, Wb, Vloc, Kp, Kp, p.pair ? BN / 2 : BN);
    if (rc) return rc;
    rc = make_tail_maps(&p, Hb, R, Wb, Vloc, Kp);
    if (rc) return rc;
    const int n_sample = topk ? sample_tiles(p) : 0;
    const int total_parts = n_sample > 0 ? 2 * p.parts : p.parts;
    ScoreTcArgs a{};
    a.R = R; a.Vloc = Vloc; a.v0 = v0; a.k = k; a.kch = p.kch; a.stages = p.stages; a.last_ksteps = p.last_ksteps; a.tail16 = p.tail16; a.pend_cap = g_pend_cap;
    a.n_tiles = p.n_tiles; a.tiles_per_split = p.tiles_per_split;
    a.bias = bias; a.target = target; a.target_score = target_score_in; a.thr_floor = g_thr_floor;
    a.pv = (float*)ws;
    a.pi = (int*)(a.pv + (size_t)total_parts * R * k);
    a.pg = a.pi + (size_t)total_parts * R * k;
    a.pt = a.pg + (size_t)total_parts * R;
    a.ps = nullptr;
    a.captured = target_score_out;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_sample > 0) {
        // launch 1 (threshold pass): the first tiles of every split, chunk maxima only -> partial slots [0, parts); the k-th best
        // of their merge seeds every list's threshold in launch 2 (a valid lower bound of the true k-th best score)
        a.tile_lo = 0; a.tile_hi = n_sample; a.part0 = 0; a.thr_init = nullptr; a.sample_mode = 1;
        a.captured = nullptr;
        rc = launch_topk(tmA, tmB, a, p, topk, false, st);
        if (rc) return rc;
        rc = launch_merge(a.pv, a.pi, nullptr, nullptr, p.parts, R, k, topk_val, topk_idx, nullptr, nullptr, st);
        if (rc) return rc;
        // launch 2: the whole sweep -> partial slots [parts, 2 parts)
        a.tile_lo = 0; a.tile_hi = p.tiles_per_split; a.part0 = p.parts; a.sample_mode = 0;
        a.captured = target_score_out;
        a.thr_init = topk_val; a.thr_stride = k; a.thr_col = k - 1;
        rc = launch_topk(tmA, tmB, a, p, topk, count, st, g_pdl != 0);
        if (rc) return rc;
        return launch_merge(a.pv + (size_t)p.parts * R * k, a.pi + (size_t)p.parts * R * k, count ? a.pg + (size_t)p.parts * R : nullptr,
                            a.pt + (size_t)p.parts * R, p.parts, R, k, topk_val, topk_idx, n_greater, n_tie_lower, st);
    }
    a.tile_lo = 0; a.tile_hi = p.tiles_per_split; a.part0 = 0; a.thr_init = nullptr;
    rc = launch_topk(tmA, tmB, a, p, topk, count, st);
    if (rc) return rc;
    return launch_merge(topk ? a.pv : nullptr, a.pi, count ? a.pg : nullptr, a.pt, total_parts, R, k, topk_val, topk_idx, n_greater,
                        n_tie_lower, st);
}

// Candidate generation for the exact top-k (csrc/rescore.cu).  With >= 4 catalog splits the sweep keeps its cheap k-entry thread
// lists (the 20- / 32-entry lists need the two-warpgroup epilogue and cost up to 2x) and the k_out candidates come from the UNION of
// the splits' lists: tiles are dealt round robin, so every split sees an even share of the best items however the catalog is ordered,
// and ``bound`` tells the certificate what the lists may have dropped.  With fewer splits (very many rows) the thread lists hold
// k_out entries themselves and the merged list is the true bf16 top-k_out (bound = -inf).
// Union mode needs enough parts that no single part holds k of the row's best ~2k items (the certificate fails for a row where one
// does): with P parts dealt round robin that happens with probability ~ P C(2k, k) P^-k -- 1e-7 per row at P = 16, k = 10.  So the
// catalog is cut into at least CAND_MIN_PARTS splits even when fewer would fill the machine (more CTAs than SMs run in waves; every
// CTA still sweeps many tiles), and catalogs of fewer tiles than that take the classic path with k_out-entry thread lists.
#define CAND_MIN_PARTS 16
static int cand_list_len(int k) { return k <= 2 ? 5 : 10; }      // thread-list entries in union mode: at least 2k spare-ish, 5 or 10
// seq (out): sequential parts per CTA.  When the row tiles alone (nearly) fill the machine the natural split count is below
// CAND_MIN_PARTS; instead of cutting the catalog into 16 short-lived CTAs per row tile, every CTA sweeps ``seq`` parts in turn and
// its four warpgroups write their lists unfolded: splits * seq * WGS lists per row (ScoreTcArgs::seq_parts).
static int make_cand_plan(int R, int Kp, int Vloc, int k, ScorePlan* p, bool* use_union, int* seq, int* unfolded = nullptr) {
    int rc = make_plan(R, Kp, Vloc, p, true);
    if (rc) return rc;
    *seq = 1;
    if (unfolded) *unfolded = 0;
    *use_union = k <= 10 && !p->pair && p->n_tiles >= CAND_MIN_PARTS;
    if (*use_union && p->splits < CAND_MIN_PARTS) {
        // measured on one GPU with the slice shapes of G ranks (1024 G rows x 1 000 003 / G items): G = 8 (2 natural splits)
        // 0.53 vs 0.87 ms; G = 2 (9 natural splits) 0.475 vs 0.415 ms -- the first scheme's threshold pass wins while a row tile
        // still gets several CTAs
        if (g_cand_seq && p->wgs == 4 && p->splits <= 3 && p->n_tiles >= 4 * CAND_MIN_PARTS) {
            int sq = ceil_div(CAND_MIN_PARTS, p->splits * p->wgs);
            if (sq < 2) sq = 2;            // (seq = 1 would be the folded single-part kernel path)
            *seq = sq;
            p->tiles_per_split = ceil_div(p->n_tiles, p->splits * sq);      // tiles per PART
            p->parts = p->splits * sq * p->wgs;                            // lists per row
        } else if (g_cand_seq && unfolded && p->wgs == 4 && p->splits * p->wgs >= CAND_MIN_PARTS) {
            // 5..15 natural splits (a 2-rank slice: 2048 rows): keep them -- one wave of CTAs -- and let the four warpgroups of
            // every CTA write their lists unfolded: splits * 4 >= 16 lists per row, threshold pass as usual
            *unfolded = 1;
            p->parts = p->splits * p->wgs;
        } else {
            p->tiles_per_split = ceil_div(p->n_tiles, CAND_MIN_PARTS);
            p->splits = ceil_div(p->n_tiles, p->tiles_per_split);
            p->parts = p->splits;
        }
    }
    return ASME_OK;
}
extern "C" size_t asme_b200_tc_score_candidates_workspace_bytes(int R, int Kp, int Vloc, int k, int k_out) {
    const int kk = k > k_out ? k : k_out;
    if (R < 1) R = 1;
    ScorePlan p;
    bool use_union = false;
    int seq = 1, unf = 0;
    if (make_cand_plan(R, Kp, Vloc, k, &p, &use_union, &seq, &unf)) return 0;
    const size_t classic = asme_b200_tc_score_topk_workspace_bytes(R, Kp, Vloc, 32) + (size_t)R * 32 * 8;
    const size_t lists = (size_t)(p.parts > p.splits ? p.parts : p.splits) + MAX_EPI_WGS;
    const size_t uni = (size_t)2 * lists * MAX_EPI_WGS * R * ((size_t)kk * 8 + 8) + (size_t)R * kk * 8;
    return classic > uni ? classic : uni;
}
extern "C" int asme_b200_tc_score_candidates(const void* Hb, int R, int Kp, const void* Wb, const float* bias, const float* bias_bounds,
                                             int v0, int Vloc,
                                             const int64_t* target, int k, int k_out, float* cand_val, int32_t* cand_idx, float* bound,
                                             float* target_score_out, void* ws, size_t ws_bytes, asme_stream_t stream) {
    ASME_REQUIRE(!bias_bounds || (bias && ((uintptr_t)bias_bounds & 7) == 0), "tc_score_candidates: bias_bounds needs the bias and 8-byte alignment");
    ASME_REQUIRE(Hb && Wb && cand_val && cand_idx && bound, "tc_score_candidates: null argument");
    ASME_REQUIRE(k >= 1 && k <= 32 && k <= k_out && k_out <= 64, "tc_score_candidates: k=%d k_out=%d unsupported (1 <= k <= 32, k <= k_out <= 64)", k, k_out);
    ASME_REQUIRE(!target_score_out || target, "tc_score_candidates: target_score_out needs target");
    ASME_REQUIRE(!bias || ((uintptr_t)bias & 15) == 0, "tc_score_candidates: bias must be 16-byte aligned");
    if (R == 0) return ASME_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ScorePlan p;
    bool use_union = false;
    int seq = 1, unfolded = 0;
    int rc = make_cand_plan(R, Kp, Vloc, k, &p, &use_union, &seq, &unfolded);
    if (rc) return rc;
    const int kl = cand_list_len(k);          // entries of the sweep's lists (>= k)
    if (!use_union) {          // classic: the true bf16 top 32 (the longest thread lists); unused candidate slots stay empty
        const int kc = k_out < 32 ? k_out : 32;
        float* tmp_v = cand_val;
        int32_t* tmp_i = cand_idx;
        if (kc != k_out) {     // the sweep writes (R, kc) rows: stage them behind the sweep's own scratch and spread them out afterwards
            const size_t off = asme_b200_tc_score_topk_workspace_bytes(R, Kp, Vloc, 32);
            ASME_REQUIRE(ws_bytes >= off + (size_t)R * kc * 8, "tc_score_candidates: workspace too small");
            tmp_v = (float*)((uint8_t*)ws + off);
            tmp_i = (int32_t*)(tmp_v + (size_t)R * kc);
        }
        rc = asme_b200_tc_score_topk(Hb, R, Kp, Wb, bias, v0, Vloc, target, nullptr, kc, tmp_v, tmp_i, target_score_out, nullptr,
                                     nullptr, ws, ws_bytes, stream);
        if (rc) return rc;
        spread_candidates_kernel<<<ceil_div((long long)R * k_out, 256), 256, 0, st>>>(tmp_v, tmp_i, R, kc, k_out, cand_val, cand_idx, bound);
        ASME_LAUNCH_OK();
        return ASME_OK;
    }
    ASME_REQUIRE(ws_bytes >= asme_b200_tc_score_candidates_workspace_bytes(R, Kp, Vloc, k, k_out), "tc_score_candidates: workspace too small");
    CUtensorMap tmA, tmB;
    rc = asme_tc_make_tmap_bf16(&tmA, Hb, R, Kp, Kp, BM);
    if (rc) return rc;
    rc = asme_tc_make_tmap_bf16(&tmB, Wb, Vloc, Kp, Kp, BN);
    if (rc) return rc;
    rc = make_tail_maps(&p, Hb, R, Wb, Vloc, Kp);
    if (rc) return rc;
    const int n_sample = seq > 1 ? 0 : sample_tiles(p);      // sequential parts seed their own thresholds
    const int total_parts = n_sample > 0 ? 2 * p.parts : p.parts;
    ScoreTcArgs a{};
    a.R = R; a.Vloc = Vloc; a.v0 = v0; a.k = kl; a.kch = p.kch; a.stages = p.stages; a.last_ksteps = p.last_ksteps; a.tail16 = p.tail16; a.pend_cap = g_pend_cap;
    a.n_tiles = p.n_tiles; a.tiles_per_split = p.tiles_per_split; a.tile_step = p.splits * seq; a.seq_parts = seq; a.unfolded = unfolded;
    a.bias = bias; a.bias_bounds = reinterpret_cast<const float2*>(bias_bounds); a.target = target; a.target_score = nullptr; a.thr_floor = g_thr_floor;
    a.pv = (float*)ws;
    a.pi = (int*)(a.pv + (size_t)total_parts * R * kl);
    a.pg = a.pi + (size_t)total_parts * R * kl;
    a.pt = a.pg + (size_t)total_parts * R;
    float* thr_v = (float*)(a.pt + (size_t)total_parts * R);
    int32_t* thr_i = (int32_t*)(thr_v + (size_t)R * kl);
    a.ps = nullptr;
    const float* thr_init = nullptr;
    const float* fpv = a.pv;
    const int* fpi = a.pi;
    if (n_sample > 0) {
        a.tile_lo = 0; a.tile_hi = n_sample; a.part0 = 0; a.thr_init = nullptr; a.sample_mode = 1; a.captured = nullptr;
        rc = launch_topk(tmA, tmB, a, p, true, false, st);
        if (rc) return rc;
        rc = launch_merge(a.pv, a.pi, nullptr, nullptr, p.parts, R, kl, thr_v, thr_i, nullptr, nullptr, st);
        if (rc) return rc;
        a.tile_lo = 0; a.tile_hi = p.tiles_per_split; a.part0 = p.parts; a.sample_mode = 0;
        a.thr_init = thr_v; a.thr_stride = kl; a.thr_col = kl - 1;
        thr_init = thr_v;
        fpv = a.pv + (size_t)p.parts * R * kl;
        fpi = a.pi + (size_t)p.parts * R * kl;
    } else {
        a.tile_lo = 0; a.tile_hi = p.tiles_per_split; a.part0 = 0; a.thr_init = nullptr;
    }
    a.captured = target_score_out;
    rc = launch_topk(tmA, tmB, a, p, true, false, st);
    if (rc) return rc;
#define CAND_MERGE(KC) tc_cand_merge_kernel<KC><<<ceil_div(R, 4), 128, 0, st>>>(fpv, fpi, p.parts, R, kl, k_out, thr_init, kl, kl - 1, cand_val, cand_idx, bound)
    if (kl == 5) CAND_MERGE(5);
    else CAND_MERGE(10);
#undef CAND_MERGE
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// diagnostic: the scoring sweep with an empty epilogue (upper bound of what the TMA -> MMA -> TMEM pipeline sustains)
extern "C" int asme_b200_tc_score_pipeline_probe(const void* Hb, int R, int Kp, const void* Wb, int Vloc, int read_tmem,
                                                 asme_stream_t stream) {
    ASME_REQUIRE(Hb && Wb, "tc_score_pipeline_probe: null operand");
    ScorePlan p;
    int rc = make_plan(R, Kp, Vloc, &p);
    if (rc) return rc;
    CUtensorMap tmA, tmB;
    rc = asme_tc_make_tmap_bf16(&tmA, Hb, R, Kp, Kp, BM);
    if (rc) return rc;
    rc = asme_tc_make_tmap_bf16(&tmB, Wb, Vloc, Kp, Kp, p.pair ? BN / 2 : BN);
    if (rc) return rc;
    rc = make_tail_maps(&p, Hb, R, Wb, Vloc, Kp);
    if (rc) return rc;
    ScoreTcArgs a{};
    a.R = R; a.Vloc = Vloc; a.kch = p.kch; a.stages = p.stages; a.last_ksteps = p.last_ksteps; a.tail16 = p.tail16; a.pend_cap = g_pend_cap; a.k = read_tmem ? 1 : 0;
    a.n_tiles = p.n_tiles; a.tiles_per_split = p.tiles_per_split; a.tile_lo = 0; a.tile_hi = p.tiles_per_split;
    return launch_score<EPI_PROBE, 0, false>(tmA, tmB, a, p, (cudaStream_t)stream);
}

extern "C" size_t asme_b200_tc_score_ce_workspace_bytes(int R, int Kp, int Vloc, int plan_rows) {
    ScorePlan p;
    if (make_plan(R < 1 ? 1 : R, Kp, Vloc, &p, false, plan_rows)) return 0;
    return (size_t)(p.splits + 1) * MAX_EPI_WGS * (R < 1 ? 1 : R) * 8;
}

extern "C" int asme_b200_tc_score_ce_partial(const void* Hb, int R, int Kp, const void* Wb, const float* bias, int v0, int Vloc,
                                             const int64_t* target, float* row_max, float* row_sumexp, float* target_logit,
                                             void* ws, size_t ws_bytes, const int32_t* n_live, int plan_rows, asme_stream_t stream) {
    ASME_REQUIRE(Hb && Wb && target && row_max && row_sumexp && target_logit, "tc_score_ce_partial: null argument");
    ASME_REQUIRE(!bias || ((uintptr_t)bias & 15) == 0, "tc_score_ce_partial: bias must be 16-byte aligned");
    if (R == 0) return ASME_OK;
    ScorePlan p;
    int rc = make_plan(R, Kp, Vloc, &p, false, n_live ? plan_rows : 0);
    if (rc) return rc;
    ASME_REQUIRE(!(n_live && p.pair), "tc_score_ce_partial: a device row count cannot be combined with CTA pairs");
    const size_t need = (size_t)p.splits * MAX_EPI_WGS * R * 8;
    if (ws_bytes < need) {
        asme_set_error("tc_score_ce_partial: workspace too small (%zu < %zu)", ws_bytes, need);
        return ASME_ERR_WORKSPACE;
    }
    CUtensorMap tmA, tmB;
    rc = asme_tc_make_tmap_bf16(&tmA, Hb, R, Kp, Kp, BM);
    if (rc) return rc;
    rc = asme_tc_make_tmap_bf16(&tmB, Wb, Vloc, Kp, Kp, p.pair ? BN / 2 : BN);
    if (rc) return rc;
    rc = make_tail_maps(&p, Hb, R, Wb, Vloc, Kp);
    if (rc) return rc;
    ScoreTcArgs a{};
    a.R = R; a.Vloc = Vloc; a.v0 = v0; a.k = 0; a.kch = p.kch; a.stages = p.stages; a.last_ksteps = p.last_ksteps; a.tail16 = p.tail16; a.pend_cap = g_pend_cap;
    a.n_tiles = p.n_tiles; a.tiles_per_split = p.tiles_per_split;
    a.bias = bias; a.target = target; a.target_score = nullptr; a.thr_floor = -INFINITY;
    a.pv = (float*)ws;
    a.ps = a.pv + (size_t)p.parts * R;
    a.tile_lo = 0; a.tile_hi = p.tiles_per_split; a.part0 = 0; a.thr_init = nullptr;
    a.captured = target_logit;      // the caller zero-fills: only the shard that owns the target column writes
    a.n_live = n_live;
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_score<EPI_CE, 0, false>(tmA, tmB, a, p, st);
    if (rc) return rc;
    tc_ce_merge_kernel<<<ceil_div(R, 128), 128, 0, st>>>(a.pv, a.ps, p.parts, R, row_max, row_sumexp, n_live);
    ASME_LAUNCH_OK();
    return ASME_OK;
}
