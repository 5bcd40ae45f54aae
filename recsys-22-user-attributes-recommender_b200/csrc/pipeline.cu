// Input pipeline on the GPU: the per-sample Python of the reference's dataset processors as batched kernels.
//   cloze masking            data/datasets/processors/cloze_mask.py:50-92   (per item: uniform draw, 80 / 10 / 10 mask / random / keep)
//   positive / negative      data/datasets/processors/pos_neg_sampler.py:41-114  (x = seq[:-1], pos = seq[1:], negatives uniform over
//   sampling                 the vocabulary minus special tokens minus the tokens of the sequence, with replacement)
// Inputs are the right-padded (B,S) id tensors the reference's collate produces (data/collate.py:42-110).  The reference draws from
// Python's / torch's global generators, so its streams cannot be reproduced; these kernels draw from the library's counter-based
// generator (a pure function of seed, sequence, position) and are tested for the same DISTRIBUTION and the same invariants.
#include "common.cuh"

// uniform in [0,1) with 24 bits, stream `stream` of element (b, s)
__device__ __forceinline__ float pipe_uniform(uint64_t seed, uint32_t stream, uint32_t b, uint32_t s) {
    const uint32_t key = asme_mix32((uint32_t)seed ^ (stream * 0x9E3779B9u)) ^ (uint32_t)(seed >> 32);
    const uint32_t x = asme_mix32(asme_mix32(b ^ key) + s * 0x85EBCA6Bu);
    return (float)(x >> 8) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ uint32_t pipe_bits(uint64_t seed, uint32_t stream, uint32_t b, uint32_t s, uint32_t try_no) {
    const uint32_t key = asme_mix32((uint32_t)seed ^ (stream * 0x9E3779B9u)) ^ (uint32_t)(seed >> 32);
    return asme_mix32(asme_mix32(asme_mix32(b ^ key) + s * 0x85EBCA6Bu) + try_no * 0xC2B2AE35u);
}

#define PIPE_MAX_FEATURES 8
struct ClozeArgs {
    int B, S, n_feat;                        // feature 0 is the item sequence
    const int64_t* in[PIPE_MAX_FEATURES];    // (B,S) each, right-padded with pad_id (feature 0 decides the length)
    int64_t* out[PIPE_MAX_FEATURES];         // masked copies
    int64_t mask_id[PIPE_MAX_FEATURES];
    int64_t vocab[PIPE_MAX_FEATURES];        // len(tokenizer): random replacements are uniform in [0, vocab - 2], as the reference draws them
    int64_t* target;                         // (B,S): original item at the selected positions, pad_id elsewhere
    int64_t pad_id;
    float mask_prob, only_last_prob;
    uint64_t seed;
};

// one warp per sequence
__global__ void cloze_mask_kernel(const ClozeArgs a) {
    const int b = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (b >= a.B) return;
    const uint64_t seed = asme_seed(a.seed);
    const int64_t* items = a.in[0] + (size_t)b * a.S;
    int len = 0;                             // number of non-pad tokens (right padding)
    for (int s0 = 0; s0 < a.S; s0 += 32) {
        const int s = s0 + lane;
        len += __popc(__ballot_sync(0xffffffffu, s < a.S && items[s] != a.pad_id));
    }
    const bool last_only = pipe_uniform(seed, 1u, (uint32_t)b, 0xffffffffu) <= a.only_last_prob;     // cloze_mask.py:61-62
    for (int s = lane; s < a.S; s += 32) {
        const size_t o = (size_t)b * a.S + s;
        const int64_t item = items[s];
        int action = 0;                      // 0 = untouched (no target), 1 = MASK, 2 = random token, 3 = keep (target only)
        if (s < len) {
            if (last_only) {
                action = s == len - 1 ? 1 : 0;
            } else {
                float u = pipe_uniform(seed, 2u, (uint32_t)b, (uint32_t)s);
                if (u < a.mask_prob) {       // :74-85
                    u = u / a.mask_prob;
                    action = u < 0.8f ? 1 : (u < 0.9f ? 2 : 3);
                }
            }
        }
        a.target[o] = action != 0 ? item : a.pad_id;
        for (int f = 0; f < a.n_feat; ++f) {
            int64_t v = a.in[f][o];
            if (action == 1) v = a.mask_id[f];
            else if (action == 2)      // random_(0, len - 1) draws from [0, len - 2]: Tensor.random_ excludes its upper end (utils.py:41-50)
                v = (int64_t)(pipe_bits(seed, 3u + (uint32_t)f, (uint32_t)b, (uint32_t)s, 0u) % (uint64_t)(a.vocab[f] > 1 ? a.vocab[f] - 1 : 1));
            a.out[f][o] = v;
        }
    }
}

extern "C" int asme_b200_cloze_mask(int B, int S, int n_feat, const int64_t* const* in, int64_t* const* out, const int64_t* mask_id,
                                    const int64_t* vocab, int64_t* target, int64_t pad_id, float mask_prob, float only_last_prob,
                                    uint64_t seed, asme_stream_t stream) {
    ASME_REQUIRE(n_feat >= 1 && n_feat <= PIPE_MAX_FEATURES, "cloze_mask: n_feat=%d (1..%d)", n_feat, PIPE_MAX_FEATURES);
    ASME_REQUIRE(in && out && mask_id && vocab && target, "cloze_mask: null argument");
    ASME_REQUIRE(mask_prob >= 0.f && mask_prob <= 1.f && only_last_prob >= 0.f && only_last_prob <= 1.f, "cloze_mask: probabilities out of range");
    if (B == 0 || S == 0) return ASME_OK;
    ClozeArgs a{};
    a.B = B; a.S = S; a.n_feat = n_feat; a.target = target; a.pad_id = pad_id; a.mask_prob = mask_prob; a.only_last_prob = only_last_prob;
    a.seed = seed;
    for (int f = 0; f < n_feat; ++f) {
        ASME_REQUIRE(in[f] && out[f] && vocab[f] >= 1, "cloze_mask: feature %d", f);
        a.in[f] = in[f]; a.out[f] = out[f]; a.mask_id[f] = mask_id[f]; a.vocab[f] = vocab[f];
    }
    cloze_mask_kernel<<<ceil_div(B, 8), 256, 0, (cudaStream_t)stream>>>(a);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// positive / negative sampling: seq (B,S1) right-padded -> x, pos, neg (B,S1-1)
// negatives: uniform over [0,V) minus the n_special first ids (special tokens occupy the lowest ids) minus the tokens of the
// sequence; rejection sampling against the sequence held in shared memory (S1 <= 1024)
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void pos_neg_kernel(const int64_t* __restrict__ seq, int B, int S1, int64_t V, int n_special, int64_t pad_id, uint64_t seed_arg,
                               int64_t* __restrict__ x, int64_t* __restrict__ pos, int64_t* __restrict__ neg) {
    extern __shared__ int64_t sh_seq[];
    const int b = blockIdx.x;
    const uint64_t seed = asme_seed(seed_arg);
    int cnt = 0;
    for (int s = threadIdx.x; s < S1; s += blockDim.x) {
        const int64_t v = seq[(size_t)b * S1 + s];
        sh_seq[s] = v;
        cnt += v != pad_id;
    }
    __shared__ int len_sh;
    if (threadIdx.x == 0) len_sh = 0;
    __syncthreads();
    atomicAdd(&len_sh, cnt);
    __syncthreads();
    const int len = len_sh;
    const int S = S1 - 1;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const size_t o = (size_t)b * S + s;
        if (s >= len - 1) {                  // beyond the shortened sequence: padding
            x[o] = pad_id; pos[o] = pad_id; neg[o] = pad_id;
            continue;
        }
        x[o] = sh_seq[s];
        pos[o] = sh_seq[s + 1];
        int64_t cand = pad_id;
        for (uint32_t t = 0; t < 4096u; ++t) {      // expected tries ~ V / (V - len - n_special): a handful at most
            cand = (int64_t)n_special + (int64_t)(pipe_bits(seed, 11u, (uint32_t)b, (uint32_t)s, t) % (uint64_t)(V - n_special));
            bool used = false;
            for (int j = 0; j < len; ++j) used |= sh_seq[j] == cand;
            if (!used) break;
        }
        neg[o] = cand;
    }
}

extern "C" int asme_b200_pos_neg_sample(const int64_t* seq, int B, int S1, int64_t V, int n_special, int64_t pad_id, uint64_t seed,
                                        int64_t* x, int64_t* pos, int64_t* neg, asme_stream_t stream) {
    ASME_REQUIRE(seq && x && pos && neg, "pos_neg_sample: null argument");
    ASME_REQUIRE(S1 >= 2 && S1 <= 1024, "pos_neg_sample: S=%d unsupported (2..1024)", S1);
    ASME_REQUIRE(n_special >= 0 && V > (int64_t)n_special + S1, "pos_neg_sample: vocabulary of %lld ids is too small for sequences of %d", (long long)V, S1);
    if (B == 0) return ASME_OK;
    pos_neg_kernel<<<B, 128, (size_t)S1 * sizeof(int64_t), (cudaStream_t)stream>>>(seq, B, S1, V, n_special, pad_id, seed, x, pos, neg);
    ASME_LAUNCH_OK();
    return ASME_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Negative sampling for sampled ranking metrics (metrics/container/metrics_sampler.py:140-204): per user n_samples distinct items
// drawn from the item weights (popularity) with the user's target and every item of the input sequence excluded -- sequential
// draws without replacement from the renormalised weights == rejection sampling of an i.i.d. candidate stream (candidates that are
// excluded or already drawn are skipped).  One warp per user: 32 candidates per round by inversion of the cumulative weights
// (binary search), accepted in lane order.
// ---------------------------------------------------------------------------------------------------------------------------
#define NEG_MAX_SAMPLES 512
__global__ void weighted_negatives_kernel(const double* __restrict__ cdf, int V, const int64_t* __restrict__ input_seq, int S,
                                          const int64_t* __restrict__ targets, int B, int n_samples, uint64_t seed_arg,
                                          int64_t* __restrict__ out, int* __restrict__ failed) {
    extern __shared__ int64_t neg_sh[];                 // per warp: [S input items][1 target][n_samples accepted]
    const int wib = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int b = blockIdx.x * (blockDim.x / 32) + wib;
    if (b >= B) return;
    int64_t* mine = neg_sh + (size_t)wib * (S + 1 + n_samples);
    const uint64_t seed = asme_seed(seed_arg);
    for (int s = lane; s < S; s += 32) mine[s] = input_seq[(size_t)b * S + s];
    if (lane == 0) mine[S] = targets[b];
    __syncwarp();
    const double total = cdf[V - 1];
    int n_acc = 0;
    for (uint32_t round = 0; n_acc < n_samples && round < 4096u; ++round) {
        // candidate of this lane: inversion of the cumulative weights
        const uint32_t r0 = pipe_bits(seed, 21u, (uint32_t)b, round * 32u + (uint32_t)lane, 0u);
        const uint32_t r1 = pipe_bits(seed, 22u, (uint32_t)b, round * 32u + (uint32_t)lane, 0u);
        const double u = ((double)r0 * 4294967296.0 + (double)r1) * (1.0 / 18446744073709551616.0) * total;
        int lo = 0, hi = V - 1;                         // first index with cdf[i] > u
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cdf[mid] > u) hi = mid; else lo = mid + 1;
        }
        const int64_t cand = lo;
        bool bad = false;
        for (int j = 0; j < S + 1 + n_acc; ++j) bad |= mine[j] == cand;
        // accept in lane order; a lane's candidate is also rejected when an earlier lane of this round drew the same item
        for (int l = 0; l < 32 && n_acc < n_samples; ++l) {
            const int64_t c = __shfl_sync(0xffffffffu, cand, l);
            const bool c_bad = __shfl_sync(0xffffffffu, (int)bad, l) != 0;
            bool dup = false;
            if (!c_bad) {
                // compare with what was accepted earlier in THIS round (positions >= the count at the start of the round are new)
                for (int j = S + 1 + lane; j < S + 1 + n_acc; j += 32) dup |= mine[j] == c;
                dup = __any_sync(0xffffffffu, dup);
            }
            if (!c_bad && !dup) {
                if (lane == 0) mine[S + 1 + n_acc] = c;
                ++n_acc;
                __syncwarp();
            }
        }
    }
    if (n_acc < n_samples && lane == 0) atomicExch(failed, 1);
    for (int j = lane; j < n_samples; j += 32) out[(size_t)b * n_samples + j] = j < n_acc ? mine[S + 1 + j] : 0;
}

extern "C" int asme_b200_weighted_negatives(const double* cdf, int V, const int64_t* input_seq, int S, const int64_t* targets, int B,
                                            int n_samples, uint64_t seed, int64_t* out, int* failed, asme_stream_t stream) {
    ASME_REQUIRE(cdf && input_seq && targets && out && failed, "weighted_negatives: null argument");
    ASME_REQUIRE(n_samples >= 1 && n_samples <= NEG_MAX_SAMPLES, "weighted_negatives: n_samples=%d (1..%d)", n_samples, NEG_MAX_SAMPLES);
    ASME_REQUIRE(V >= 1 && S >= 0, "weighted_negatives: bad shape");
    if (B == 0) return ASME_OK;
    const int warps = 4;
    const size_t smem = (size_t)warps * (S + 1 + n_samples) * sizeof(int64_t);
    ASME_REQUIRE(smem <= 200 * 1024, "weighted_negatives: S=%d with %d samples needs too much shared memory", S, n_samples);
    { const int _rc = asme_ensure_max_smem((const void*)weighted_negatives_kernel); if (_rc) return _rc; }
    weighted_negatives_kernel<<<ceil_div(B, warps), warps * 32, smem, (cudaStream_t)stream>>>(cdf, V, input_seq, S, targets, B, n_samples,
                                                                                           seed, out, failed);
    ASME_LAUNCH_OK();
    return ASME_OK;
}
