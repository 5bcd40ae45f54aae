/* asme_b200.h -- C ABI of libasme_b200.so: the sm_100a (B200) kernels behind the ASME
 * sequential-recommender hot path.
 *
 * The reference (LSX-UniWue/recsys-22-user-attributes-recommender) has NO native code and no
 * FFI: its plug-in surface is Python (modules/registry.py:19-22, init/factories/include/
 * import_factory.py:51-81).  Each entry point below therefore cites the reference *call site*
 * (file:line under /root/reference/src/asme/core) whose stock-ATen arithmetic it replaces; the
 * Python binding a maintainer adds (ctypes, no torch types crossing the boundary) is shown in
 * INTEGRATION.md and implemented in asme_b200/_lib.py.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; kernels never allocate: scratch is
 *     passed as (ws, ws_bytes) and sized by the matching *_workspace_bytes() query;
 *   - all launches go to the caller's `stream`; entry points are re-entrant across streams, use
 *     no host threads and never synchronise the host;
 *   - ids are int64 (as delivered by the reference collate, data/collate.py:42-110), catalog
 *     indices returned as int32, activations / weights fp32 unless a `dt` argument says otherwise
 *     (ASME_DT_F32 = 0, ASME_DT_BF16 = 1);
 *   - return value: 0 on success, negative ASME_ERR_* otherwise, message via asme_b200_last_error().
 */
#ifndef ASME_B200_H
#define ASME_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* asme_stream_t; /* == cudaStream_t */

#define ASME_OK 0
#define ASME_ERR_INVALID (-1)   /* bad argument / unsupported shape */
#define ASME_ERR_CUDA (-2)      /* CUDA runtime error (launch, attribute, ...) */
#define ASME_ERR_WORKSPACE (-3) /* workspace too small */

#define ASME_DT_F32 0
#define ASME_DT_BF16 1

#define ASME_MAX_ATTR 8

const char* asme_b200_last_error(void);
int asme_b200_abi_version(void);
/* number of CUDA kernels this library has launched in this process (diagnostics / bench.py "gpu_launches") */
long long asme_b200_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * K1-K4  fused embedding gather-and-sum (+LayerNorm +dropout)
 * replaces: nn.Embedding item lookup            models/common/layers/sequence_embedding.py:87
 *           positional lookup + add             models/common/layers/transformer_layers.py:68,75
 *           attribute lookups / LinearUpscaler  models/kebert4rec/components.py:57-60, layers.py:24-27
 *           LayerNorm + Dropout                 transformer_layers.py:76-78, kebert4rec/components.py:61-62
 *   x = E[item] (+ P[t mod S]);  [x = drop_a(LN1(x))];  x += sum_a A_a[attr_a] + sum_b (sum_{j: id!=0} Wt_b[bag_b[j]] + bias_b);
 *   [x = drop_b(LN2(x))]
 * No branch on PAD/MASK ids (quirk Q9).  BERT4Rec: P = NULL (Q1), LN1 only.  KeBERT4Rec: LN2 only.
 * SASRec: LN1 and LN2 (Q2).  Bag tables are passed TRANSPOSED, Wt = Linear.weight^T, (Va,H).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const int64_t* item_ids; /* (T) = (B*S) */
    const float* item_table; /* (V,H) */
    const float* pos_table;  /* (>=S,H) or NULL */
    int n_attr;
    const int64_t* attr_ids[ASME_MAX_ATTR];  /* (T) each */
    const float* attr_table[ASME_MAX_ATTR];  /* (Va,H) */
    int n_bag;
    const int64_t* bag_ids[ASME_MAX_ATTR];   /* (T,width) each, id 0 contributes nothing */
    int bag_width[ASME_MAX_ATTR];
    const float* bag_table_t[ASME_MAX_ATTR]; /* (Va,H) = Linear.weight transposed */
    const float* bag_bias[ASME_MAX_ATTR];    /* (H), always added */
    const float* ln1_gamma;                  /* (H) or NULL */
    const float* ln1_beta;
    const float* ln2_gamma;                  /* (H) or NULL */
    const float* ln2_beta;
    float p_drop;                            /* 0 => identity (eval) */
    uint64_t seed;
    uint32_t site_a, site_b;                 /* dropout site ids of drop_a / drop_b */
    /* user prefix (models/ubert4rec/components.py:96-133): with n_user > 0 every sequence has S positions of which position 0
     * is sum_u U_u[user_ids_u[b]] (no positional row, no LN1, no item attributes) and positions 1..S-1 are the item tokens
     * b*(S-1) .. b*(S-1)+S-2 (item_ids / attr_ids / bag_ids stay (B*(S-1)) long, positional rows 0..S-2).  seg_table (2,H) or
     * NULL: row 0 is added to the user position, row 1 to the item positions, before LN2. */
    int n_user;
    const int64_t* user_ids[ASME_MAX_ATTR];  /* (B) each */
    const float* user_table[ASME_MAX_ATTR];  /* (Vu,H) */
    const float* seg_table;
    /* forward only: the first encoder block's input LayerNorm (transformer_layers.py:120-130) applied to the output rows in the
     * same pass: next_out (T,H) bf16 = LN(out; next_gamma, next_beta), next_stats (2,T) = row mean / rstd or NULL */
    const float* next_gamma;                 /* Basket inputs (N,S,BS) -- dense fallback (models/common/layers/sequence_embedding.py:9-45, :83-93): the item embeddings of a step
 * pooled over its basket, out (T,H) = sum / mean / max over j of table[ids[t, j]] (mode 0 / 1 / 2; PAD slots take part like any id);
 * arg (T,H) uint8 = slot of the maximum (max only; first slot on ties).  The pooled rows enter asme_b200_embed_fwd as a table indexed by
 * the token number.  Backward: d_rows (T*BS, H) = the gradient every (token, slot) receives -- feed it to
 * asme_b200_embgrad_sorted_reduce with the flattened ids. */
int asme_b200_embed_pool_fwd(const int64_t* ids, const float* table, int T, int BS, int H, int mode, float* out, uint8_t* arg,
                             asme_stream_t stream);
int asme_b200_embed_pool_bwd(const float* d_out, const uint8_t* arg, int T, int BS, int H, int mode, float* d_rows,
                             asme_stream_t stream);
/* (H) or NULL */
    const float* next_beta;
    void* next_out;
    float* next_stats;
} asme_embed_desc;

int asme_b200_embed_fwd(const asme_embed_desc* d, int T, int S, int H, float* out /*T,H*/,
                        float* stats /* (4,T): mean1,rstd1,mean2,rstd2; may be NULL in eval */, asme_stream_t stream);
/* backward: dOut (T,H) -> d_item_rows (T,H) = gradient w.r.t. (E[item]+P) rows, d_attr_rows (T,H) = gradient w.r.t.
 * the attribute sum (== d_item_rows when LN1 is absent; may alias); LayerNorm parameter gradients are ACCUMULATED
 * into dln (4,H): dgamma1,dbeta1,dgamma2,dbeta2 via deterministic two-stage column sums. */
size_t asme_b200_embed_bwd_workspace_bytes(int T, int H);
/* with a user prefix d_item_rows / d_attr_rows of a user position hold the gradient w.r.t. the user-embedding sum */
int asme_b200_embed_bwd(const asme_embed_desc* d, int T, int S, int H, const float* d_out, const float* stats,
                        float* d_item_rows, float* d_attr_rows, float* dln, void* ws, size_t ws_bytes,
                        asme_stream_t stream);

/* K21  deterministic embedding-table gradient: sort (id) -> segmented reduce -> one write per distinct row.
 * replaces: autograd embedding_dense_backward (atomic scatter-add) of K1-K3/K13.
 * d_table[ids[t], :] += d_rows[t / row_divisor, :] for all t with ids[t] != skip_id (skip_id = -1 keeps everything;
 * row_divisor = bag width for (T,width) id bags that share one gradient row per token, else 1). */
size_t asme_b200_embgrad_workspace_bytes(int T, int H);
int asme_b200_embgrad_sorted_reduce(const int64_t* ids, int T, const float* d_rows, int row_divisor, int H, float* d_table,
                                    int V, int64_t skip_id, void* ws, size_t ws_bytes, asme_stream_t stream);
/* the two phases of the call above, separately: the sort depends on the ids only (not on any gradient), so a training step can run
 * it early / on another stream; the reduce consumes the sorted pairs the sort left in the SAME workspace. */
int asme_b200_embgrad_sort(const int64_t* ids, int T, int row_divisor, int H, int V, int64_t skip_id, void* ws, size_t ws_bytes,
                           asme_stream_t stream);
int asme_b200_embgrad_reduce_sorted(int T, const float* d_rows, int H, float* d_table, int V, void* ws, size_t ws_bytes,
                                    asme_stream_t stream);
/* d_pos[s,:] += sum_b d_rows[b*S+s,:]   (positions are generated, t mod S; transformer_layers.py:68) */
int asme_b200_posgrad_reduce(const float* d_rows, int B, int S, int H, float* d_pos, asme_stream_t stream);
/* the same for sequences that are seq_stride_rows rows apart (user prefix: d_rows points at position 1 of sequence 0) */
int asme_b200_posgrad_reduce_strided(const float* d_rows, int B, int S, int seq_stride_rows, int H, float* d_pos,
                                     asme_stream_t stream);
/* bag tables: d_table_t[id,:] += d_rows[t,:] for every bag entry id != 0; d_bias += column sums of d_rows */
int asme_b200_colsum_accumulate(const float* x, int M, int N, float* out /*N, +=*/, void* ws, size_t ws_bytes,
                                asme_stream_t stream);
/* the same over the first *m_live rows only (device row count of a row selection, see asme_b200_select_rows; NULL = all M) */
int asme_b200_colsum_accumulate_live(const float* x, int M, int N, float* out /*N, +=*/, void* ws, size_t ws_bytes,
                                     const int32_t* m_live, asme_stream_t stream);
size_t asme_b200_colsum_workspace_bytes(int M, int N);

/* diagnostic: knob 0 = 1 (default): the forward row kernels (embedding gather, LayerNorm) give a row to H/16 lanes with four
 * 128-bit chunks each; 0: H/4 lanes with one chunk (the first layout, ALU-bound at H = 128).  Results differ only in the summation
 * order of the LayerNorm statistics. */
int asme_b200_rowwise_tune(int knob, int value);
/* ------------------------------------------------------------------------------------------
 * K4/K6/K11  LayerNorm (eps 1e-5, biased variance) forward / backward
 * replaces: nn.LayerNorm in SublayerConnection (transformer_layers.py:120-130) and the FFN modifier
 *           (models/common/components/representation_modifier/ffn_modifier.py:18-26)
 * ------------------------------------------------------------------------------------------ */
int asme_b200_layernorm_fwd(const float* x, const float* gamma, const float* beta, int M, int H, float* y,
                            float* stats /* (2,M) mean,rstd or NULL */, const int32_t* n_live /* device row count or NULL */,
                            asme_stream_t stream);
/* LayerNorm whose output feeds a tensor-core GEMM: bf16 copy of y (y_f32 optional) */
int asme_b200_layernorm_fwd_bf16(const float* x, const float* gamma, const float* beta, int M, int H, float* y_f32 /*NULL ok*/,
                                 void* y_bf16, float* stats, asme_stream_t stream);
/* y_f32 = x * mask(site_a); y_bf16 = bf16(y_f32 * mask(site_b)); a site of 0 (or p = 0) means no mask; y_f32 may be NULL */
int asme_b200_dropout_cast(const float* x, long long n, float p, uint64_t seed, uint32_t site_a, uint32_t site_b,
                           float* y_f32, void* y_bf16, asme_stream_t stream);
/* dx = (d_residual ? d_residual : 0) + LN'(dy); dgamma/dbeta ACCUMULATED (dgb = (2,H)) */
size_t asme_b200_layernorm_bwd_workspace_bytes(int M, int H);
int asme_b200_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* stats, int M, int H,
                            const float* d_residual, float* dx, float* dgb, void* ws, size_t ws_bytes,
                            const int32_t* n_live /* device row count or NULL */, asme_stream_t stream);
/* the same, also emitting what the next stage of the backward pass consumes (replaces an asme_b200_dropout_cast launch):
 * dx (fp32) = (LayerNorm gradient + d_residual) * mask(site_a), dx_bf16 (M,H) = bf16(dx * mask(site_b)); site 0 = no mask */
int asme_b200_layernorm_bwd_drop(const float* dy, const float* x, const float* gamma, const float* stats, int M, int H,
                                 const float* d_residual, float* dx, float* dgb, void* ws, size_t ws_bytes, float p_drop,
                                 uint64_t seed, uint32_t site_a, uint32_t site_b, void* dx_bf16, asme_stream_t stream);
/* dgb == NULL in the two calls above: the (gamma, beta) partials stay in ws as [asme_b200_layernorm_bwd_chunks(M, H)][2H] and
 * asme_b200_rows_reduce(ws, chunks, 2H, dgb, 1) finishes them -- a leaf of the backward pass that can run on another stream. */
int asme_b200_layernorm_bwd_chunks(int M, int H);
int asme_b200_rows_reduce(const float* partial, int chunks, int N, float* out, int accumulate, asme_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K7/K9/K10/K11/K12  dense layers: C = epilogue(A x op(B)), fp32 SIMT path (strict 1e-5 parity mode)
 * replaces: nn.Linear q/k/v/o (transformer_layers.py:190-199), FFN (transformer_layers.py:220),
 *           modifier Linear (ffn_modifier.py:19), projections (layers.py:109,142-143), and their autograd.
 *   trans_b = 1: B is (N,K) row-major (nn.Linear weight), C = A B^T       (forward)
 *   trans_b = 0: B is (K,N) row-major,                     C = A B         (dX = dY W)
 * epilogue, in order:  v = acc + bias[n];  if pre_act: pre_act[m,n] = v;  act;  v *= gelu'(mul_gelu_grad_of[m,n]);
 *                      dropout(site);  v += residual[m,n];  C[m,n] = v
 * ------------------------------------------------------------------------------------------ */
#define ASME_ACT_NONE 0
#define ASME_ACT_GELU 1
typedef struct {
    const float* bias;             /* (N) or NULL */
    float* pre_act;                /* (M,N) or NULL */
    int act;                       /* ASME_ACT_* */
    const float* mul_gelu_grad_of; /* (M,N) or NULL: multiply by gelu'(.) (backward through GELU) */
    float p_drop;                  /* epilogue dropout, 0 = none */
    uint64_t seed;
    uint32_t site;
    const float* residual;         /* (M,N) or NULL */
    const int32_t* m_live;         /* device count of live rows of A / C (row selections, asme_b200_select_rows) or NULL = all M */
} asme_gemm_epilogue;

int asme_b200_gemm(const float* A, const float* B, float* C, int M, int N, int K, int trans_b,
                   const asme_gemm_epilogue* epi /* may be NULL */, asme_stream_t stream);
/* weight / bias gradient: dW[N,K] (+)= dY[M,N]^T X[M,K], dbias[N] (+)= colsum(dY); split over M with a
 * deterministic second-stage reduction. accumulate = 0 overwrites. */
size_t asme_b200_gemm_wgrad_workspace_bytes(int M, int N, int K);
int asme_b200_gemm_wgrad(const float* dY, const float* X, int M, int N, int K, float* dW, float* dbias /*NULL ok*/,
                         int accumulate, void* ws, size_t ws_bytes, const int32_t* m_live /* device row count or NULL */,
                         asme_stream_t stream);

/* elementwise dropout with the same (seed, site, index) masks the fused epilogues use */
int asme_b200_dropout(const float* x, float* y, long long n, float p, uint64_t seed, uint32_t site,
                      asme_stream_t stream);
/* dz = dy * gelu'(z) (exact erf GELU, ffn_modifier.py:20) */
/* n_live != NULL: only the first *n_live rows of row_width elements */
int asme_b200_gelu_bwd(const float* dy, const float* z, float* dz, long long n, int row_width, const int32_t* n_live,
                       asme_stream_t stream);
/* y = a (*|+) b : post-fusion merge (kebert4rec/components.py:110-113); op 0 = add, 1 = multiply */
int asme_b200_binary(const float* a, const float* b, float* y, long long n, int op, asme_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K5+K8  masked attention for short sequences (S <= 256, head dim <= 64), mask generated in-kernel
 * replaces: mask materialisation (models/transformer/sequence_representation.py:33-48) and
 *           Attention.forward (transformer_layers.py:145-155): QK^T/sqrt(d) -> masked_fill(mask==0,-1e9)
 *           -> softmax -> dropout -> .V ; a fully masked row attends uniformly to all S keys (Q4).
 * qkv: (B*S, 3H) = [q | k | v] per token, head h at columns h*d..h*d+d-1 of each third.
 * key_valid: (B,S) uint8 padding mask or NULL (bidirectional without mask). ctx: (B*S, H).
 * stats: (2, B*heads*S) row max and row sum of exp (needed by backward), may be NULL in eval.
 * ------------------------------------------------------------------------------------------ */
int asme_b200_attn_fwd(const float* qkv, const uint8_t* key_valid, int B, int S, int heads, int d, int causal,
                       float p_drop, uint64_t seed, uint32_t site, float* ctx, float* stats, asme_stream_t stream);
int asme_b200_attn_bwd(const float* qkv, const uint8_t* key_valid, int B, int S, int heads, int d, int causal,
                       float p_drop, uint64_t seed, uint32_t site, const float* ctx, const float* d_ctx,
                       const float* stats, float* d_qkv /* (B*S,3H) */, void* ws, size_t ws_bytes,
                       asme_stream_t stream);
size_t asme_b200_attn_bwd_workspace_bytes(int B, int S, int heads);

/* ------------------------------------------------------------------------------------------
 * K12+K15+K18-K20  full-catalog scoring fused with top-k + exact target rank (logits never reach HBM)
 * replaces: ItemEmbeddingProjectionLayer / LinearProjectionLayer (layers.py:105-143) on the selected row
 *           (masked_training_module.py:80-91, next_item_prediction_training_module.py:226-244),
 *           AllItemsSampler multi-hot (metrics/container/metrics_sampler.py:45-71) and the per-metric
 *           full argsort (metrics/common.py:18-27).
 * Scores rows Hrows (R,H) against catalog slice W (Vloc,H) (+bias), global ids v0..v0+Vloc-1.
 * Per row: top-k (score desc, id asc), n_greater = #{s_j > s_t}, n_tie_lower = #{j < t : s_j == s_t} given the
 * target score s_t (target_score (R), computed by asme_b200_score_targets; shards all-reduce it first).
 * ------------------------------------------------------------------------------------------ */
int asme_b200_score_targets(const float* Hrows, int R, int H, const float* W, const float* bias, int v0, int Vloc,
                            const int64_t* target, float* target_score /* (R) +=, owner shard only */,
                            asme_stream_t stream);
size_t asme_b200_score_topk_workspace_bytes(int R, int Vloc, int k);
int asme_b200_score_topk_rank(const float* Hrows, int R, int H, const float* W, const float* bias, int v0, int Vloc,
                              const int64_t* target, const float* target_score, int k, float* topk_val /*R,k*/,
                              int32_t* topk_idx /*R,k*/, int32_t* n_greater /*R*/, int32_t* n_tie_lower /*R*/,
                              void* ws, size_t ws_bytes, asme_stream_t stream);
/* k-way merge of G partial top-k lists (vals/idx: (G,R,k)) with (score desc, id asc) order */
int asme_b200_topk_merge(const float* vals, const int32_t* idx, int G, int R, int k, float* out_val,
                         int32_t* out_idx, asme_stream_t stream);
/* K20: sums over the batch of recall/NDCG/MRR/precision @ ks from the 1-based target rank
 * (metrics/common.py:66-175, metrics/mrr.py:26-37). out: (4, n_k) sums, ACCUMULATED. */
int asme_b200_ranking_metrics(const int32_t* rank, int R, const int32_t* ks, int n_k, float* out, asme_stream_t stream);

/* dense-signature compatibility path of RankingMetric.update(predictions (N,I), positive_item_mask (N,I) int64,
 * metric_mask (N,I) int64 or NULL) (metrics/metric.py:57-83, metrics/common.py:4-175): per row, O(I) scan with
 * (score desc, id asc) order. out (7,N): recall, precision, DCG, NDCG, MRR, F1 @k and the full-sort rank of the
 * worst relevant item (metrics/common.py:30-46). */
int asme_b200_dense_ranking(const float* pred, const int64_t* pos_mask, const int64_t* metric_mask, int N, int I, int k,
                            float* out, asme_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K12+K16  scoring fused with log-softmax cross-entropy (ignore_index rows are skipped by the caller:
 * rows = positions with target != pad).  replaces layers.py:105-143 + nn.CrossEntropyLoss
 * (masked_training_module.py:107-111, losses/sasrec/sas_rec_losses.py:16-32).
 * partial: per row (max, sumexp, target logit) over the slice [v0, v0+Vloc) -- shards combine with
 * all-reduce(MAX)/(SUM).  loss_sum += sum_r (max_r + log(sumexp_r) - target_logit_r).
 * Device row counts: the rows may be a row selection of CAPACITY R whose live count is only known on the device
 * (asme_b200_select_rows); every entry point below then takes ``n_live`` (NULL = all R rows live), works on the live rows only,
 * and the mean's 1/n is applied on the device: ce_loss_from_partials writes loss_mean = loss_sum / *n_live, the backward entry
 * points use scale / *n_live.  One CUDA graph then serves every batch of a shape and the host never waits for the count.
 * ------------------------------------------------------------------------------------------ */
size_t asme_b200_score_ce_workspace_bytes(int R, int Vloc);
int asme_b200_score_ce_partial(const float* Hrows, int R, int H, const float* W, const float* bias, int v0, int Vloc,
                               const int64_t* target, float* row_max, float* row_sumexp, float* target_logit,
                               void* ws, size_t ws_bytes, const int32_t* n_live, asme_stream_t stream);
int asme_b200_ce_loss_from_partials(const float* row_max, const float* row_sumexp, const float* target_logit, int R,
                                    float* lse /*R*/, float* loss_sum /*1, +=*/, const int32_t* n_live,
                                    float* loss_mean /*1 or NULL: loss_sum / live rows*/, asme_stream_t stream);
/* vocab-sharded scoring: merge of the shards' softmax statistics -- pm / ps (G,R) per-shard row maxima / sum-exps -> those of the
 * whole catalog; and a shard's sum-exp re-expressed against the all-reduced row maximum (row_sumexp * exp(row_max - global_max)) */
int asme_b200_ce_combine(const float* pm, const float* ps, int G, int R, float* row_max, float* row_sumexp, asme_stream_t stream);
int asme_b200_ce_rescale(const float* row_sumexp, const float* row_max, const float* global_max, int R, float* out,
                         asme_stream_t stream);
/* backward: dlogit = (softmax - onehot) * scale; dH (R,H) = dlogit W (overwritten; shards all-reduce),
 * dW (Vloc,H) += dlogit^T H, dbias (Vloc) += colsum(dlogit). */
size_t asme_b200_score_ce_bwd_workspace_bytes(int R, int H, int Vloc);
int asme_b200_score_ce_bwd(const float* Hrows, int R, int H, const float* W, const float* bias, int v0, int Vloc,
                           const int64_t* target, const float* lse, float scale, float* dH, float* dW, float* dbias,
                           void* ws, size_t ws_bytes, const int32_t* n_live, asme_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Tensor-core (tcgen05 + TMEM + TMA) scoring path, bf16 operands / fp32 accumulation.
 * Same reference call sites as the fp32 entry points above (layers.py:105-143 + metrics/common.py:18-27 /
 * masked_training_module.py:107-111); the (R x V) logits exist only as 128x256 accumulator tiles in tensor memory.
 * Operands are prepared with asme_b200_cast_bf16: Hb (R,Kp) and Wb (Vloc,Kp) bf16 row-major, Kp = hidden size
 * zero-padded to a multiple of 16 (<= 272; a multiple of 64 for the CE backward).
 * ------------------------------------------------------------------------------------------ */
/* y[r, 0..ld_out) = bf16(x[r, 0..cols)) zero padded; ld_out % 4 == 0 */
/* n_live != NULL: rows past *n_live are written as ZEROS (the result is a tensor-core operand: no stale NaNs) */
int asme_b200_cast_bf16(const float* x, void* y, long long rows, int cols, int ld_in, int ld_out, const int32_t* n_live,
                        asme_stream_t stream);
/* as above plus two extra columns right after `cols` that fold a per-row bias into the contraction (h.w + b == [h,1,1].[w,b_hi,b_lo]):
 * mode 1 appends (1, 1) (activations), mode 2 appends bf16 hi / lo parts of bias[r] (weights); ld_out >= cols + 2, multiple of 16
 * for the scoring kernels, which then take bias = NULL and Kp = ld_out */
int asme_b200_cast_bf16_ext(const float* x, const float* bias, void* y, long long rows, int cols, int ld_in, int ld_out, int mode,
                            asme_stream_t stream);
/* One sweep over the catalog slice [v0, v0+Vloc):
 *   k > 0                      -> per row top-k (score desc, id asc): topk_val / topk_idx (R,k)
 *   target_score_out != NULL   -> score of the row's target column exactly as this kernel computes it (written by the
 *                                 shard that owns the column; caller zero-fills, shards all-reduce(SUM))
 *   target_score_in != NULL    -> n_greater = #{s_j > s_t}, n_tie_lower = #{j < t : s_j == s_t}   (exact full rank)
 * When only @k metrics are needed the rank is the target's position in the top-k list and no count pass is run. */
size_t asme_b200_tc_score_topk_workspace_bytes(int R, int Kp, int Vloc, int k);
int asme_b200_tc_score_topk(const void* Hb, int R, int Kp, const void* Wb, const float* bias, int v0, int Vloc,
                            const int64_t* target, const float* target_score_in, int k, float* topk_val,
                            int32_t* topk_idx, float* target_score_out, int32_t* n_greater, int32_t* n_tie_lower,
                            void* ws, size_t ws_bytes, asme_stream_t stream);
/* ------------------------------------------------------------------------------------------
 * EXACT top-k on the tensor-core path (metrics/common.py:18-27 sorts the fp32 logits).  The bf16 sweep above is the candidate
 * generator (KC > k best items by bf16-operand score); asme_b200_topk_rescore re-scores the candidates from the fp32 hidden rows
 * and the fp32 table slice [v0, v0+V) with the arithmetic of the fp32 path (sequential fmaf chain, + bias: bit-identical to
 * asme_b200_score_topk_rank / asme_b200_score_targets), orders them (score desc, id asc) and CERTIFIES each row: with
 * E = ||h-h~|| max||w_v|| + ||h~|| max||w_v-w~_v|| + 2^-15 max|b_v| + 2^-16 ||h|| max||w_v|| (h~, w~ the bf16 operands) bounding
 * |exact - bf16| for every item, no item outside the list can
 * enter the top k when (bf16 score of anything outside the list) + E < (exact k-th best candidate score).  Rows that cannot be certified get
 * row_flag = 1 (n_flagged counts them) and are re-run exactly by asme_b200_score_topk_flagged -- all decided on the device.
 *   norm_bound (3) = {max_v ||w_v||_2, max_v |b_v|, max_v ||w_v - bf16(w_v)||_2} from asme_b200_table_norm_bound (cache it until
 *   the weights change)
 *   target (R) optional: target_score (R) = exact score, written by the slice that owns the target's row (caller zero-fills);
 *   rank (R) = 1-based position of the target among the exact top k, k+1 when it is not among them
 * ------------------------------------------------------------------------------------------ */
int asme_b200_table_norm_bound(const float* W, int V, int H, const float* bias /*NULL ok*/, float* out3, asme_stream_t stream);
/* candidates for the exact top-k: cand_val / cand_idx (R,k_out) = the k_out best by bf16 score among what the sweep's lists kept,
 * bound (R) = upper bound of the bf16 score of every item in NONE of the row's lists (-inf: the list is the true bf16 top k_out) */
/* bias_bounds (ceil(Vloc/32), 2) = {max, min} of bias over every 32-column chunk of the slice (asme_b200_bias_chunk_bounds, cache it
 * until the weights change): with it the sweep keeps the bias OUT of the contraction (Kp = hidden size instead of hidden + 16 folded
 * columns, layers.py:138-143) and out of the hot epilogue -- the exact fp32 bias is added only to chunks whose best raw score plus
 * the chunk's largest bias could pass the row's threshold. */
int asme_b200_bias_chunk_bounds(const float* bias, int V, float* bounds, asme_stream_t stream);
size_t asme_b200_tc_score_candidates_workspace_bytes(int R, int Kp, int Vloc, int k, int k_out);
int asme_b200_tc_score_candidates(const void* Hb, int R, int Kp, const void* Wb, const float* bias, const float* bias_bounds /*NULL ok*/,
                                  int v0, int Vloc,
                                  const int64_t* target /*NULL ok*/, int k, int k_out, float* cand_val, int32_t* cand_idx, float* bound,
                                  float* target_score_out /*NULL ok: bf16 score of the target, owner shard writes*/, void* ws,
                                  size_t ws_bytes, asme_stream_t stream);
int asme_b200_topk_rescore(const float* Hrows, int R, int H, const float* W, const float* bias, int v0, int V,
                           const int32_t* cand_idx /*R,KC global ids, -1 = empty*/, const float* cand_val /*R,KC bf16-sweep scores*/,
                           const float* cand_bound /*R or NULL*/, int KC, int k, const float* norm_bound, const int64_t* target /*NULL ok*/, float* topk_val /*R,k*/,
                           int32_t* topk_idx /*R,k*/, float* target_score /*NULL ok*/, int32_t* rank /*NULL ok*/,
                           int32_t* row_flag /*R*/, int32_t* n_flagged /*1*/, asme_stream_t stream);
/* the exact fp32 sweep (asme_b200_score_topk_rank) restricted to the rows with row_flag != 0: only row tiles holding a flagged row
 * run, only flagged rows of topk_val / topk_idx / rank (= exact FULL rank, needs target + target_score) are overwritten */
size_t asme_b200_score_topk_flagged_workspace_bytes(int R, int Vloc);
int asme_b200_score_topk_flagged(const float* Hrows, int R, int H, const float* W, const float* bias, int v0, int Vloc,
                                 const int64_t* target, const float* target_score, int k, const int32_t* row_flag,
                                 float* topk_val, int32_t* topk_idx, int32_t* rank /*NULL ok*/, void* ws, size_t ws_bytes,
                                 asme_stream_t stream);
/* diagnostic: the same sweep with an empty epilogue -- the ceiling of the TMA -> tcgen05.mma -> TMEM pipeline for this shape */
int asme_b200_tc_score_pipeline_probe(const void* Hb, int R, int Kp, const void* Wb, int Vloc, int read_tmem /* also read (and
    discard) every accumulator: TMEM read throughput */, asme_stream_t stream);
/* diagnostic tuning knobs of the scoring sweep (results never depend on them): knob 0 / 3 = epilogue warpgroups (2 | 4) of
 * the top-k / the CE and count-only sweeps, knob 4 = CTA pairs, knob 5 = programmatic dependent launch, knob 6 = 16-column K tail
 * staged with the 32-byte swizzle (default on), knob 1 = sample-sweep divisor (0: no sample sweep; default 16), knob 2 = reject every top-k candidate (cost of the insertion-free sweep),
 * knob 7 = candidate FIFO depth (8..16), knob 8 = candidate sweeps over many row tiles: 1 (default) sequential parts per CTA with
 * unfolded per-thread lists, 0 the catalog cut into 16 splits per row tile */
int asme_b200_tc_score_tune(int knob, int value);
/* cross-entropy partials over the slice: row_max, row_sumexp (natural units, combine across shards as for the fp32
 * entry point) and target_logit (owner shard writes; caller zero-fills) */
/* n_live / plan_rows: device row count of a row selection and the host's guess of it (0 = R), from which only the split of the
 * catalog over CTAs is chosen -- a wrong guess costs balance, never correctness; the workspace query takes the same guess */
size_t asme_b200_tc_score_ce_workspace_bytes(int R, int Kp, int Vloc, int plan_rows);
int asme_b200_tc_score_ce_partial(const void* Hb, int R, int Kp, const void* Wb, const float* bias, int v0, int Vloc,
                                  const int64_t* target, float* row_max, float* row_sumexp, float* target_logit,
                                  void* ws, size_t ws_bytes, const int32_t* n_live, int plan_rows, asme_stream_t stream);

/* backward of the fused scoring + cross-entropy layer on the tensor cores: dlogit = (softmax - onehot) * scale is recomputed tile
 * by tile in tensor memory; dH (R,H) fp32 is overwritten, dW (Vloc,H) and dbias (Vloc) are accumulated (+=); any of the three
 * may be NULL (dbias needs dW).  lse = row_max + log(row_sumexp) over the WHOLE catalog (shards combine first). */
size_t asme_b200_tc_score_ce_bwd_workspace_bytes(int R, int H, int Kp, int Vloc, int plan_rows /* as for the forward; 0 = R */);
int asme_b200_tc_score_ce_bwd(const void* Hb, int R, int H, int Kp, const void* Wb, const float* bias, int v0, int Vloc,
                              const int64_t* target, const float* lse, float scale, float* dH, float* dW, float* dbias,
                              void* ws, size_t ws_bytes, const int32_t* n_live, int plan_rows, asme_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Tensor-core dense layers (tcgen05 + TMEM + TMA), bf16 operands / fp32 accumulation.  Same call sites as asme_b200_gemm.
 *   b_is_kn = 0: B is (N,K) row-major (nn.Linear weight), C = A B^T      forward
 *   b_is_kn = 1: B is (K,N) row-major,                     C = A B        dX = dY W   (MN-major operand, no transpose copy)
 * A (M,K) bf16; N in 32..256 (multiple of 32), K in {64,128,192,256}.  Epilogue, in order: + bias[n]; pre_act_bf16 = v;
 * GELU; v *= gelu'(gelu_grad_of[m,n]); Philox dropout(seed, site, m*N+n); + residual[m,n]; second dropout(post_site, 0 = none);
 * out_f32 and/or out_bf16 (row stride ld_bf16).
 * ------------------------------------------------------------------------------------------ */
int asme_b200_tc_gemm(const void* A, const void* B, int M, int N, int K, int b_is_kn, const float* bias, int act,
                      const void* gelu_grad_of, float p_drop, unsigned long long seed, unsigned int site,
                      unsigned int post_site, const float* residual, float* out_f32, void* out_bf16, int ld_bf16, void* pre_act_bf16,
                      asme_stream_t stream);
/* the same with a fused LayerNorm of the fp32 output rows (transformer_layers.py:120-130: the rows are the next sublayer's
 * LayerNorm input): ln_out (M,N) bf16 = LN(out_f32; ln_gamma, ln_beta), ln_stats (2,M) = row mean / rstd or NULL.
 * Needs out_f32 and N <= 128 (one column tile owns whole rows). */
int asme_b200_tc_gemm_ln(const void* A, const void* B, int M, int N, int K, int b_is_kn, const float* bias, int act,
                         const void* gelu_grad_of, float p_drop, unsigned long long seed, unsigned int site,
                         unsigned int post_site, const float* residual, float* out_f32, void* out_bf16, int ld_bf16,
                         void* pre_act_bf16, const float* ln_gamma, const float* ln_beta, void* ln_out, float* ln_stats,
                         asme_stream_t stream);
/* Position-wise feed-forward block as ONE kernel (inference; models/common/layers/transformer_layers.py:217-220 + the residual of
 * SublayerConnection :120-130): out = residual + W2 gelu(W1 Y + b1) + b2, the (M, FF) intermediate stays in shared / tensor memory.
 * Y (M,H) bf16 = the LayerNorm'ed input, W1 (FF,H) / W2 (H,FF) bf16 (nn.Linear layouts), residual (M,H) fp32; H in {64, 128},
 * FF % 64 == 0.  out_f32 (M,H) and / or ln_out (M,H) bf16 = LayerNorm(out; ln_gamma, ln_beta) (the next block's input).  The fp32
 * output equals asme_b200_tc_gemm(act = GELU, bf16 out) followed by asme_b200_tc_gemm(residual) bit for bit. */
int asme_b200_tc_ffn_fused(const void* Y, const void* W1, const float* b1, const void* W2, const float* b2, const float* residual,
                           int M, int H, int FF, float* out_f32, const float* ln_gamma, const float* ln_beta, void* ln_out,
                           asme_stream_t stream);
/* The whole tail of an encoder block as ONE kernel (inference; transformer_layers.py:181-199 output_linear + :120-130 residual +
 * LayerNorm of the output sublayer + :217-220 feed-forward + its residual [+ the next block's LayerNorm]):
 *   x2 = residual + ctx Wo^T + bo;  y = LayerNorm(x2; pro_gamma, pro_beta);  out = x2 + W2 gelu(W1 y + b1) + b2
 * ctx (M,H) bf16 = attention output, Wo (H,H) bf16, residual (M,H) fp32 = the block's input rows.  x2, y and the (M, FF) intermediate
 * stay in tensor / shared memory.  Outputs as asme_b200_tc_ffn_fused. */
int asme_b200_tc_block_tail_fused(const void* ctx, const void* Wo, const float* bo, const float* residual, const float* pro_gamma,
                                  const float* pro_beta, const void* W1, const float* b1, const void* W2, const float* b2, int M, int H,
                                  int FF, float* out_f32, const float* ln_gamma, const float* ln_beta, void* ln_out, asme_stream_t stream);
/* diagnostic: knob 0 = GELU warpgroups of asme_b200_tc_ffn_fused (0 = automatic, default; 2; 4) */
int asme_b200_tc_ffn_tune(int knob, int value);
/* diagnostic: knob 0 selects the tall kernel (1 = persistent CTAs with a resident weight tile, default; 0 = one CTA per tile) */
int asme_b200_tc_gemm_tune(int knob, int value);
/* dW (N,K) fp32 (+)= dY(M,N)^T X(M,K), dbias (N) (+)= colsum(dY); dY, X bf16; token contraction split over the SMs with a
 * deterministic second-stage reduction */
size_t asme_b200_tc_wgrad_workspace_bytes(int M, int N, int K);
int asme_b200_tc_wgrad(const void* dY, const void* X, int M, int N, int K, float* dW, float* dbias, int accumulate,
                       void* ws, size_t ws_bytes, asme_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K5+K8 on the tensor cores: masked attention for S <= 256, head dim 16/32/64, hidden % 64 == 0.
 * qkv: (B*S, 3H) bf16 = [q | k | v] per token (the bf16 output of the QKV projection, read by TMA); ctx: (B*S, H) bf16.
 * Same semantics, statistics layout and Philox dropout stream as asme_b200_attn_fwd.
 * ------------------------------------------------------------------------------------------ */
int asme_b200_tc_attn_fwd(const void* qkv, const uint8_t* key_valid, int B, int S, int heads, int d, int causal,
                          float p_drop, unsigned long long seed, unsigned int site, void* ctx, float* stats,
                          uint32_t* keep_bits /* (B*heads*S, 8) dropout keep bits for the backward pass, or NULL */,
                          asme_stream_t stream);
/* evaluation of selected positions (the last encoder layer only needs the hidden state of one position per sequence): only the
 * 128-query tile that holds flat row only_row[b] (= b*S + position) of sequence b is computed; the other rows of ctx are left
 * untouched.  Rows that are computed are bit-identical to asme_b200_tc_attn_fwd. */
int asme_b200_tc_attn_fwd_rows(const void* qkv, const uint8_t* key_valid, int B, int S, int heads, int d, int causal,
                               const int64_t* only_row, void* ctx, asme_stream_t stream);
/* one query position per sequence as two matrix-vector products (HBM-bound, fp32 arithmetic on the bf16 q/k/v): ctx_rows (B, H)
 * bf16, row b = attention output of flat row only_row[b] (= b*S + position) -- Attention.forward (transformer_layers.py:145-155)
 * restricted to the rows the last encoder layer needs (models/.../evaluation reads one position per sequence).  d: power of
 * two 8..256, H/8 must divide 256. */
int asme_b200_attn_row_fwd(const void* qkv, const uint8_t* key_valid, int B, int S, int heads, int d, int causal,
                           const int64_t* only_row, void* ctx_rows, asme_stream_t stream);
/* d_qkv (B*S, 3H) bf16 from d_ctx (B*S, H) bf16; scores and dP are recomputed in tensor memory, transposed quantities come
 * from transposed MMAs; needs the forward's ctx, stats and (when p_drop > 0) keep_bits */
int asme_b200_tc_attn_bwd(const void* qkv, const uint8_t* key_valid, int B, int S, int heads, int d, int causal,
                          float p_drop, const void* ctx, const void* d_ctx, const float* stats,
                          const uint32_t* keep_bits, void* d_qkv, asme_stream_t stream);
/* diagnostic: knob 0 selects the backward kernel (1 = single sweep, default; 0 = two sweeps), knob 1 the epilogue warpgroups of the
 * single-sweep kernel (2 or 4), knob 2 the forward kernel (2 = probabilities in tensor memory as the A operand of the second MMA,
 * two CTAs per SM, default when the grid fills the machine; 3 = the same for every grid size; 1 = probabilities through shared memory);
 * results agree to rounding */
int asme_b200_tc_attn_tune(int knob, int value);

/* ------------------------------------------------------------------------------------------
 * K13+K17  SASRec positive/negative dot products fused with the BCE loss
 * replaces: SASRecProjectionComponent.forward train branch (models/sasrec/components.py:35-44) and
 *           sas_rec_binary_cross_entropy (losses/sasrec/sas_rec_losses.py:47-75).
 * sums[0] += sum_t mask_t (-log(sigmoid(p_t)+1e-24) - log(1-sigmoid(n_t)+1e-24)); sums[1] += sum_t mask_t
 * ------------------------------------------------------------------------------------------ */
int asme_b200_posneg_bce_fwd(const float* Hseq /*T,H*/, const float* E, const int64_t* pos, const int64_t* neg,
                             const uint8_t* mask, int T, int H, float* pos_logit, float* neg_logit, float* sums,
                             asme_stream_t stream);
/* dH (T,H) overwritten; d_pos_rows / d_neg_rows (T,H) = per-token gradients of E[pos], E[neg] (feed embgrad). */
int asme_b200_posneg_bce_bwd(const float* Hseq, const float* E, const int64_t* pos, const int64_t* neg,
                             const uint8_t* mask, int T, int H, const float* pos_logit, const float* neg_logit,
                             const float* sums, float dloss, float* dH, float* d_pos_rows, float* d_neg_rows,
                             asme_stream_t stream);

/* rows gather: out[r,:] = x[row_index[r],:]  (K15 row select) and its transpose scatter (rows are distinct); negative indices
 * and rows past *n_live (device row count or NULL) are skipped */
int asme_b200_gather_rows(const float* x, const int64_t* row_index, int R, int H, float* out, const int32_t* n_live,
                          asme_stream_t stream);
int asme_b200_scatter_rows(const float* rows, const int64_t* row_index, int R, int H, float* out /*T,H*/, const int32_t* n_live,
                           asme_stream_t stream);
/* Row selection ON THE DEVICE: rows[] = the flat positions t (ascending) with target[t] != ignore_id -- the rows
 * nn.CrossEntropyLoss(ignore_index=pad) sees (modules/masked_training_module.py:93-111, losses/sasrec/sas_rec_losses.py:16-32) --
 * row_targets[] = their targets, n_rows[0] = how many.  rows / row_targets have capacity T; slots past n_rows carry -1 / ignore_id.
 * Replaces torch.nonzero + index_select (a host synchronisation per batch, and a different launch geometry per batch). */
size_t asme_b200_select_rows_workspace_bytes(long long T);
int asme_b200_select_rows(const int64_t* target, long long T, int64_t ignore_id, int64_t* rows, int64_t* row_targets,
                          int32_t* n_rows, void* ws, size_t ws_bytes, asme_stream_t stream);
/* Gradient clipping by global L2 norm over the flat gradient arena: grad *= min(1, max_norm / (||grad|| + 1e-6))
 * (pl.Trainer(gradient_clip_val=...) -> torch.nn.utils.clip_grad_norm_); norm_out (1) optional; deterministic reduction */
size_t asme_b200_clip_grad_norm_workspace_bytes(void);
int asme_b200_clip_grad_norm(float* grad, long long n, float max_norm, float* norm_out, void* ws, size_t ws_bytes,
                             asme_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K22  fused Adam over the flat parameter arena (torch.optim.Adam semantics: L2 decay added to the gradient;
 * reference defaults beta = (0.99, 0.998), masked_training_module.py:165-168)
 * ------------------------------------------------------------------------------------------ */
int asme_b200_adam_step(float* param, const float* grad, float* m, float* v, long long n, double lr, double beta1,
                        double beta2, double eps, double weight_decay, int step, asme_stream_t stream);
int asme_b200_fill(float* x, long long n, float value, asme_stream_t stream);
/* Device-resident step state for CUDA-graph replays of a whole training step: struct {uint64 seed; double lr; int64 adam_step}.
 * asme_b200_step_state_advance (first node of the graph) bumps seed and adam_step; every dropout `seed` argument may be passed
 * as (1<<63 | device pointer to the state) instead of a value; asme_b200_adam_step_dev reads step and lr from the state. */
int asme_b200_step_state_advance(void* state, asme_stream_t stream);
int asme_b200_adam_step_dev(float* param, const float* grad, float* m, float* v, long long n, const void* state, double beta1,
                            double beta2, double eps, double weight_decay, asme_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Input pipeline on the GPU.  Ids are int64 (B,S), right-padded with pad_id, as the reference's collate
 * delivers them (data/collate.py:42-110).  Streams are counter-based (seed, sequence, position): same distribution and
 * invariants as the reference's processors, not the same draws (those come from Python's global generator).
 *   cloze_mask      data/datasets/processors/cloze_mask.py:50-92: with probability only_last_prob only the last item of a sequence is
 *                   masked; otherwise every item is selected with probability mask_prob and then replaced by the MASK id (80 %), a
 *                   uniform random id in [0, vocab-2] (10 %; Tensor.random_ excludes its upper end) or kept (10 %); target = original item at selected positions, pad_id
 *                   elsewhere.  Feature 0 is the item sequence; further sequence features are masked at the same positions.
 *   pos_neg_sample  data/datasets/processors/pos_neg_sampler.py:41-114: x = seq[:-1], pos = seq[1:], neg = uniform over the
 *                   vocabulary minus the n_special lowest ids minus the tokens of the sequence (with replacement); outputs (B,S-1).
 * ------------------------------------------------------------------------------------------ */
int asme_b200_cloze_mask(int B, int S, int n_feat, const int64_t* const* in, int64_t* const* out, const int64_t* mask_id,
                         const int64_t* vocab, int64_t* target, int64_t pad_id, float mask_prob, float only_last_prob,
                         uint64_t seed, asme_stream_t stream);
int asme_b200_pos_neg_sample(const int64_t* seq, int B, int S, int64_t V, int n_special, int64_t pad_id, uint64_t seed,
                             int64_t* x, int64_t* pos, int64_t* neg, asme_stream_t stream);

/* Negatives for sampled ranking metrics (metrics/container/metrics_sampler.py:140-204): for every user n_samples DISTINCT items drawn
 * from the item weights (cdf = cumulative sum of the weights, float64, (V)) with the user's target and all items of the input
 * sequence (B,S) excluded.  *failed is set to 1 when a user has fewer than n_samples admissible items. */
int asme_b200_weighted_negatives(const double* cdf, int V, const int64_t* input_seq, int S, const int64_t* targets, int B,
                                 int n_samples, uint64_t seed, int64_t* out, int* failed, asme_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ASME_B200_H */
