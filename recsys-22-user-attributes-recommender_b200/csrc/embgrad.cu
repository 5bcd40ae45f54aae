// Deterministic embedding-table gradient (K21): sort token ids, then segmented-reduce the per-token
// gradient rows so that every distinct table row is written exactly once, by exactly one warp, in a
// fixed summation order (token order inside a run).  Replaces the atomic scatter-add of
// embedding_dense_backward, which is non-deterministic on CUDA.
//
//   1. keys = ids (int32; skipped / out-of-range ids -> V), vals = token index
//   2. stable LSD radix sort of (key, val) pairs (cub::DeviceRadixSort, key bits = ceil(log2 V))
//   3. one warp per chunk of 32 sorted entries: runs that live entirely inside the chunk are summed
//      and written directly; the pieces of runs crossing a chunk boundary go to carry buffers
//   4. one warp per run that crosses chunk boundaries sums its pieces in chunk order
// Long runs (PAD = 45 % and MASK = 20 % of the tokens of a cloze batch) are thereby reduced by many
// warps in parallel instead of serialising on one.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <limits.h>

#define CHUNK 32
#define MAXQ 4   // H <= 512 : up to 4 float4 per lane

__global__ void embgrad_keys_kernel(const int64_t* __restrict__ ids, int T, int V, int64_t skip_id, int row_divisor,
                                    int* __restrict__ keys, int* __restrict__ vals) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int64_t id = ids[t];
    keys[t] = (id == skip_id || id < 0 || id >= V) ? V : (int)id;   // V = 'skipped' sentinel, sorts last
    vals[t] = t / row_divisor;   // index of the gradient row this entry reads
}

struct Acc {
    float4 v[MAXQ];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int q = 0; q < MAXQ; ++q) v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ void add_row(const float* __restrict__ row, int H, int lane) {
#pragma unroll
        for (int q = 0; q < MAXQ; ++q) {
            const int c = lane * 4 + 128 * q;
            if (c < H) add4(v[q], ldg4(row + c));
        }
    }
    __device__ __forceinline__ void add_row_plain(const float* row, int H, int lane) {      // generic loads (shared memory)
#pragma unroll
        for (int q = 0; q < MAXQ; ++q) {
            const int c = lane * 4 + 128 * q;
            if (c < H) add4(v[q], *reinterpret_cast<const float4*>(row + c));
        }
    }
    __device__ __forceinline__ void store(float* __restrict__ row, int H, int lane) const {
#pragma unroll
        for (int q = 0; q < MAXQ; ++q) {
            const int c = lane * 4 + 128 * q;
            if (c < H) *reinterpret_cast<float4*>(row + c) = v[q];
        }
    }
    __device__ __forceinline__ void add_to(float* __restrict__ row, int H, int lane) const {
#pragma unroll
        for (int q = 0; q < MAXQ; ++q) {
            const int c = lane * 4 + 128 * q;
            if (c < H) {
                float4 cur = *reinterpret_cast<float4*>(row + c);
                add4(cur, v[q]);
                *reinterpret_cast<float4*>(row + c) = cur;
            }
        }
    }
};

// flags per chunk: bit0 head valid (run started before the chunk), bit1 head run continues past the chunk,
//                  bit2 tail valid (run starts in this chunk and continues into the next)
__global__ void __launch_bounds__(128) embgrad_chunk_kernel(const int* __restrict__ keys, const int* __restrict__ vals, int T, int V,
                                                            const float* __restrict__ d_rows, int H, float* __restrict__ d_table,
                                                            float* __restrict__ head, float* __restrict__ tail,
                                                            int* __restrict__ tail_key, int* __restrict__ flags, int chunks) {
    const int c = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (c >= chunks) return;
    const int base = c * CHUNK;
    const int key = base + lane < T ? keys[base + lane] : V;
    const int val = base + lane < T ? vals[base + lane] : 0;
    const int prev_key = base > 0 ? keys[base - 1] : -1;
    const int next_key = base + CHUNK < T ? keys[base + CHUNK] : V;
    int fl = 0;
    Acc acc;
    acc.zero();
    int run_start = 0;
    // rows are fetched eight at a time before any of them is consumed: the chunk used to pay one dependent L2 latency per entry
    const bool narrow = H <= 128;              // one float4 per lane covers the row
    bool done = false;
    for (int rb = 0; rb < CHUNK && !done; rb += 8) {
        float4 buf[8];
        if (narrow) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int v = __shfl_sync(0xffffffffu, val, rb + j);
                buf[j] = lane * 4 < H ? ldg4(d_rows + (size_t)v * H + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int r = rb + j;
            const int k = __shfl_sync(0xffffffffu, key, r);
            if (k >= V) { done = true; break; }   // sorted: only skipped entries follow (warp-uniform)
            if (narrow) {
                add4(acc.v[0], buf[j]);
            } else {
                const int v = __shfl_sync(0xffffffffu, val, r);
                acc.add_row(d_rows + (size_t)v * H, H, lane);
            }
            const int k_next = r + 1 < CHUNK ? __shfl_sync(0xffffffffu, key, r + 1) : next_key;
            if (k_next != k) {
                const bool started_before = run_start == 0 && prev_key == k;
                if (started_before) { acc.store(head + (size_t)c * H, H, lane); fl |= 1; }
                else acc.add_to(d_table + (size_t)k * H, H, lane);
                acc.zero();
                run_start = r + 1;
            } else if (r + 1 == CHUNK) {   // run continues into the next chunk
                const bool started_before = run_start == 0 && prev_key == k;
                if (started_before) { acc.store(head + (size_t)c * H, H, lane); fl |= 3; }
                else {
                    acc.store(tail + (size_t)c * H, H, lane);
                    fl |= 4;
                    if (lane == 0) tail_key[c] = k;
                }
            }
        }
    }
    if (lane == 0) flags[c] = fl;
}

// One CTA per chunk whose tail starts a run that continues into later chunks.  Long runs (PAD / MASK ids cover thousands
// of tokens = hundreds of chunks) used to be summed by ONE warp, one dependent 256-byte load after another (~370 us per
// call); now warp 0 finds the length of the run with ballots over the chunk flags, the CARRY_WARPS warps sum a strided
// subset of the head partials each, and the warp partials are combined in a fixed order (deterministic).
#define CARRY_WARPS 16
__global__ void __launch_bounds__(CARRY_WARPS * 32) embgrad_carry_kernel(const float* __restrict__ head, const float* __restrict__ tail,
                                                                         const int* __restrict__ tail_key,
                                                                         const int* __restrict__ flags, int chunks, int H,
                                                                         float* __restrict__ d_table) {
    extern __shared__ float carry_smem[];      // [CARRY_WARPS][H]
    __shared__ int run_len;
    const int c = blockIdx.x;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (!(flags[c] & 4)) return;               // CTA-uniform
    if (warp == 0) {
        // number of following chunks that hold a head piece of this run
        int len = 0;
        for (int base = c + 1; base < chunks; base += 32) {
            const int cc = base + lane;
            const int f = cc < chunks ? flags[cc] : 0;
            const unsigned full = __ballot_sync(0xffffffffu, (f & 3) == 3);     // head piece, run goes on
            const unsigned any = __ballot_sync(0xffffffffu, (f & 1) != 0);      // head piece
            const int n_full = full == 0xffffffffu ? 32 : __ffs(~full) - 1;     // leading chunks that continue the run
            if (n_full == 32) { len += 32; continue; }
            len += n_full + (((any >> n_full) & 1u) ? 1 : 0);                   // the last piece ends inside its chunk
            break;
        }
        if (lane == 0) run_len = len;
    }
    __syncthreads();
    const int len = run_len;
    Acc acc;
    acc.zero();
    if (warp == 0) acc.add_row(tail + (size_t)c * H, H, lane);
    {   // four independent partial sums per warp keep four row loads in flight; combined in a fixed order
        Acc a1, a2, a3;
        a1.zero(); a2.zero(); a3.zero();
        int i = warp;
        for (; i + 3 * CARRY_WARPS < len; i += 4 * CARRY_WARPS) {
            acc.add_row(head + (size_t)(c + 1 + i) * H, H, lane);
            a1.add_row(head + (size_t)(c + 1 + i + CARRY_WARPS) * H, H, lane);
            a2.add_row(head + (size_t)(c + 1 + i + 2 * CARRY_WARPS) * H, H, lane);
            a3.add_row(head + (size_t)(c + 1 + i + 3 * CARRY_WARPS) * H, H, lane);
        }
        for (; i < len; i += CARRY_WARPS) acc.add_row(head + (size_t)(c + 1 + i) * H, H, lane);
#pragma unroll
        for (int q = 0; q < MAXQ; ++q) {
            add4(a2.v[q], a3.v[q]);
            add4(a1.v[q], a2.v[q]);
            add4(acc.v[q], a1.v[q]);
        }
    }
    acc.store(carry_smem + (size_t)warp * H, H, lane);
    __syncthreads();
    if (warp == 0) {
        acc.zero();
        for (int w = 0; w < CARRY_WARPS; ++w) acc.add_row_plain(carry_smem + (size_t)w * H, H, lane);
        acc.add_to(d_table + (size_t)tail_key[c] * H, H, lane);
    }
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

static size_t cub_temp_bytes(int T) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr, T);
    return bytes;
}

extern "C" size_t asme_b200_embgrad_workspace_bytes(int T, int H) {
    if (T <= 0) return 0;
    const size_t chunks = (size_t)ceil_div(T, CHUNK);
    return 4 * align256((size_t)T * sizeof(int)) + align256(cub_temp_bytes(T)) + 2 * align256(chunks * H * sizeof(float)) +
           2 * align256(chunks * sizeof(int));
}

struct EmbgradLayout {
    int *keys_in, *vals_in, *keys_out, *vals_out;
    void* temp;
    size_t temp_bytes;
    float *head, *tail;
    int *tail_key, *flags;
    int chunks;
};
static EmbgradLayout embgrad_layout(void* ws, int T, int H) {
    EmbgradLayout L;
    L.chunks = ceil_div(T, CHUNK);
    char* p = (char*)ws;
    L.keys_in = (int*)p;  p += align256((size_t)T * sizeof(int));
    L.vals_in = (int*)p;  p += align256((size_t)T * sizeof(int));
    L.keys_out = (int*)p; p += align256((size_t)T * sizeof(int));
    L.vals_out = (int*)p; p += align256((size_t)T * sizeof(int));
    L.temp_bytes = cub_temp_bytes(T);
    L.temp = p;           p += align256(L.temp_bytes);
    L.head = (float*)p;   p += align256((size_t)L.chunks * H * sizeof(float));
    L.tail = (float*)p;   p += align256((size_t)L.chunks * H * sizeof(float));
    L.tail_key = (int*)p; p += align256((size_t)L.chunks * sizeof(int));
    L.flags = (int*)p;
    return L;
}

// Phase 1: (id, token) pairs sorted by id into the workspace.  Depends on the ids only, not on any gradient -- a training step
// can run it at the start of the backward pass on a second stream (asme_b200/models.py) and keep the workspace until phase 2.
extern "C" int asme_b200_embgrad_sort(const int64_t* ids, int T, int row_divisor, int H, int V, int64_t skip_id, void* ws,
                                      size_t ws_bytes, asme_stream_t stream) {
    ASME_REQUIRE(row_divisor >= 1, "embgrad: row_divisor=%d", row_divisor);
    ASME_REQUIRE(ids, "embgrad: null argument");
    ASME_REQUIRE(H % 4 == 0 && H >= 4 && H <= 512, "embgrad: H=%d unsupported (4..512, multiple of 4)", H);
    ASME_REQUIRE(V >= 1, "embgrad: V=%d", V);
    if (T == 0) return ASME_OK;
    if (ws_bytes < asme_b200_embgrad_workspace_bytes(T, H)) {
        asme_set_error("embgrad: workspace too small");
        return ASME_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const EmbgradLayout L = embgrad_layout(ws, T, H);
    embgrad_keys_kernel<<<ceil_div(T, 256), 256, 0, st>>>(ids, T, V, skip_id, row_divisor, L.keys_in, L.vals_in);
    ASME_LAUNCH_OK();
    int bits = 1;
    while ((1LL << bits) <= (long long)V) ++bits;   // keys are in [0, V]
    size_t temp_bytes = L.temp_bytes;
    ASME_CUDA_OK(cub::DeviceRadixSort::SortPairs(L.temp, temp_bytes, L.keys_in, L.keys_out, L.vals_in, L.vals_out, T, 0, bits, st));
    return ASME_OK;
}

// Phase 2: segmented sums of d_rows over the sorted pairs phase 1 left in the workspace; d_table[id] += ...
extern "C" int asme_b200_embgrad_reduce_sorted(int T, const float* d_rows, int H, float* d_table, int V, void* ws, size_t ws_bytes,
                                               asme_stream_t stream) {
    ASME_REQUIRE(d_rows && d_table, "embgrad: null argument");
    ASME_REQUIRE(H % 4 == 0 && H >= 4 && H <= 512, "embgrad: H=%d unsupported (4..512, multiple of 4)", H);
    if (T == 0) return ASME_OK;
    if (ws_bytes < asme_b200_embgrad_workspace_bytes(T, H)) {
        asme_set_error("embgrad: workspace too small");
        return ASME_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const EmbgradLayout L = embgrad_layout(ws, T, H);
    embgrad_chunk_kernel<<<ceil_div(L.chunks, 4), 128, 0, st>>>(L.keys_out, L.vals_out, T, V, d_rows, H, d_table, L.head, L.tail,
                                                                L.tail_key, L.flags, L.chunks);
    ASME_LAUNCH_OK();
    embgrad_carry_kernel<<<L.chunks, CARRY_WARPS * 32, (size_t)CARRY_WARPS * H * sizeof(float), st>>>(L.head, L.tail, L.tail_key, L.flags,
                                                                                                     L.chunks, H, d_table);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

extern "C" int asme_b200_embgrad_sorted_reduce(const int64_t* ids, int T, const float* d_rows, int row_divisor, int H,
                                               float* d_table, int V, int64_t skip_id, void* ws, size_t ws_bytes,
                                               asme_stream_t stream) {
    ASME_REQUIRE(ids && d_rows && d_table, "embgrad: null argument");
    int rc = asme_b200_embgrad_sort(ids, T, row_divisor, H, V, skip_id, ws, ws_bytes, stream);
    if (rc) return rc;
    return asme_b200_embgrad_reduce_sorted(T, d_rows, H, d_table, V, ws, ws_bytes, stream);
}
