/*
 * arxiv_rag_b200 — C ABI of the B200-native retrieval hot path.
 *
 * This is the drop-in boundary for the ONE data-parallel path of matiasrodlo/arxiv-rag that
 * BASELINE.json's north_star names:
 *   stage A  batched all-mpnet-base-v2 encode: token ids -> MPNet encoder -> masked mean-pool
 *            -> L2 normalise.  Replaces what `model.encode(batch, normalize_embeddings=True,
 *            convert_to_numpy=True)` computes at
 *            4-embed/generation/generate_embeddings_parallel.py:146-153 (and :160-165), i.e.
 *            sentence-transformers -> transformers MPNetModel.forward (modeling_mpnet.py:403-455).
 *   stage B  exact cosine top-k of query embeddings against the chunk-embedding matrix.  The
 *            reference has no such routine; the definition generalised is
 *            TextChunker._cosine_similarity, 3-chunks/pipeline/src/processors/text_processor.py:1601-1605,
 *            with top_k from 3-chunks/pipeline/config.yaml:62-64.
 *
 * The reference is pure Python, so the binding a maintainer adds is a ctypes stub.
 * Conventions: every function returns 0 on success or a negative ARB_ERR_* code; the message is
 * available from arb_last_error().  No exceptions cross the boundary.  The caller owns every
 * input/output buffer; "dev" pointers are CUDA device pointers on the current device, "host"
 * pointers are ordinary host memory.  All work is enqueued on the given stream (a cudaStream_t
 * passed as void*, NULL = legacy default stream) and the call returns after enqueue.
 * There is no CPU fallback: without an sm_100a device every compute entry point fails.
 */
#ifndef ARXIV_RAG_B200_H
#define ARXIV_RAG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARB_OK 0
#define ARB_ERR_INVALID (-1)     /* bad shape / dtype / alignment / null pointer */
#define ARB_ERR_CUDA (-2)        /* a CUDA call failed */
#define ARB_ERR_WORKSPACE (-3)   /* caller workspace too small */
#define ARB_ERR_UNSUPPORTED (-4)

#define ARB_DTYPE_F32 0
#define ARB_DTYPE_BF16 1
#define ARB_DTYPE_F16 2
#define ARB_DTYPE_BF16_PURE 3 /* encoder handles only: bf16 without the short-batch fp16 path */

#define ARB_EPI_BIAS 0
#define ARB_EPI_BIAS_GELU 1
#define ARB_EPI_BIAS_RESIDUAL 2

/* Thread-local message of the last failing call on this thread. */
const char* arb_last_error(void);
/* ABI version of this header (bumped on any signature change). */
int arb_abi_version(void);

/* ------------------------------------------------------------------------------------------
 * Stage A — MPNet sentence encoder (replaces SentenceTransformer('all-mpnet-base-v2').encode,
 * generate_embeddings_parallel.py:47,146-153; math: modeling_mpnet.py:57-96,116-186,189-273,
 * 284-360,403-455 + sentence-transformers Pooling(mean) + Normalize).
 * ------------------------------------------------------------------------------------------ */
typedef struct ArbMpnetConfig {
    int32_t vocab_size;                     /* 30527 */
    int32_t max_position_embeddings;        /* 514 */
    int32_t hidden_size;                    /* 768 */
    int32_t num_layers;                     /* 12 */
    int32_t num_heads;                      /* 12 */
    int32_t intermediate_size;              /* 3072 */
    int32_t relative_attention_num_buckets; /* 32 */
    int32_t pad_token_id;                   /* 1 (MPNetEmbeddings.padding_idx, modeling_mpnet.py:60) */
    float layer_norm_eps;                   /* 1e-5 for all-mpnet-base-v2 */
    int32_t compute_dtype;                  /* 16-bit format of weights + activations in HBM and of
                                               the tensor-core operands (accumulation, softmax and
                                               LayerNorm statistics are always fp32; one format
                                               per tcgen05 MMA — mixed A/B formats fault on B200):
                                               ARB_DTYPE_F16   fp16: the shipped default. Cosine vs
                                                 the fp32 reference >= 0.9999 on every row (~0.999997);
                                               ARB_DTYPE_BF16  what BASELINE configs[1] names: bf16
                                                 for batches of >= arb_mpnet_short_seq() tokens,
                                                 an fp16 copy of the weights for shorter batches
                                                 (8-bit mantissas cannot hold 0.9999 on rows of a
                                                 few tokens). ~0.99995 on full rows of well-behaved
                                                 weights, below 0.9999 on heavy-tailed ones;
                                               ARB_DTYPE_BF16_PURE  bf16 for every batch (A/B) */
    int32_t position_mode;                  /* 0 = MPNet: padding-aware position ids (modeling_mpnet.py:889-897);
                                               1 = BERT: absolute index (all-MiniLM-L6-v2; token-type row 0 is
                                               folded into the position table by the caller). A BERT-style
                                               encoder also sets relative_attention_num_buckets = 0. */
} ArbMpnetConfig;

/* All pointers are HOST fp32 arrays in the nn.Module layouts ([out,in] for Linear weights). */
typedef struct ArbMpnetLayerWeights {
    const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b; /* attention.attn.{q,k,v,o} */
    const float *attn_ln_g, *attn_ln_b;                         /* attention.LayerNorm */
    const float *ffn_in_w, *ffn_in_b;                           /* intermediate.dense [I,H] */
    const float *ffn_out_w, *ffn_out_b;                         /* output.dense [H,I] */
    const float *out_ln_g, *out_ln_b;                           /* output.LayerNorm */
} ArbMpnetLayerWeights;

typedef struct ArbMpnetWeights {
    const float* word_embeddings;         /* [vocab, H] */
    const float* position_embeddings;     /* [max_pos, H] */
    const float *emb_ln_g, *emb_ln_b;     /* embeddings.LayerNorm */
    const float* relative_attention_bias; /* encoder.relative_attention_bias.weight [buckets, heads] */
    const ArbMpnetLayerWeights* layers;   /* [num_layers] */
} ArbMpnetWeights;

/* Uploads and packs the weights (16-bit, fused QKV) and allocates activations for up to
 * max_tokens = B*S tokens per call on CUDA device `device`. */
int arb_mpnet_create(const ArbMpnetConfig* cfg, const ArbMpnetWeights* weights, int64_t max_tokens,
                     int32_t max_seq, int32_t device, void** handle);
int arb_mpnet_destroy(void* handle);
/* Bytes of device memory held by the handle (weights + activations). */
int64_t arb_mpnet_device_bytes(void* handle);
/* ids/mask: dev int32 [B,S] (mask 1 = token, 0 = padding; ids of padding = pad_token_id);
 * out: dev fp32 [B,hidden] unit-norm rows in input order. Rows whose mask is all zero -> 0. */
int arb_mpnet_encode(void* handle, const int32_t* ids_dev, const int32_t* mask_dev, int32_t B,
                     int32_t S, float* out_dev, void* stream);
/* Data errors found by the kernels of earlier arb_mpnet_encode calls (a token id outside
 * [0, vocab_size) — torch's embedding raises on it; the kernel clamps the gather and records the
 * id). Call after the stream has been synchronised: ARB_OK, or ARB_ERR_INVALID (message in
 * arb_last_error(); the status is cleared). */
int arb_mpnet_status(void* handle);
/* ARB_DTYPE_BF16 handles: batches with S below this run in fp16; hosts that sort rows by length
 * should cut their batches at this length. 0 for the other dtypes. */
int arb_mpnet_short_seq(void* handle);
/* Number of kernel launches one arb_mpnet_encode call enqueues (for launch accounting). */
int arb_mpnet_launches_per_encode(void* handle);
/* MPNetEncoder.relative_position_bucket (modeling_mpnet.py:343-360) for relative_position = j-i. */
int arb_mpnet_relative_bucket(int32_t relative_position, int32_t num_buckets, int32_t max_distance);

/* ------------------------------------------------------------------------------------------
 * Stage B — exact cosine top-k search over unit-norm rows.
 * scores[q, r] = <queries[q], corpus[id]>, rows ordered by (score desc, id asc);
 * out_ids = local row + id_offset (int64), unused slots (k > N) hold score -inf / id -1.
 * dtype ARB_DTYPE_BF16: operands bf16, scores accumulate bf16 products exactly in fp32.
 * dtype ARB_DTYPE_F32 : operands fp32, scored in one tf32 tensor-core pass over the stored rows; the
 *                       k + 22 best are re-scored with exact fp32 FMAs and re-ranked (returned scores are
 *                       true fp32 dot products). See arb_topk_search_f32 for the exactness verdict.
 * ------------------------------------------------------------------------------------------ */
size_t arb_topk_search_workspace_bytes(int32_t dtype, int64_t Q, int64_t N, int32_t D, int32_t k);
int arb_topk_search(const void* queries_dev, const void* corpus_dev, int32_t dtype, int64_t Q,
                    int64_t N, int32_t D, int32_t k, float* out_scores_dev, int64_t* out_ids_dev,
                    int64_t id_offset, void* workspace_dev, size_t workspace_bytes, void* stream);
/* fp32 search with its exactness verdict.
 * mode 0: tf32 pass + exact re-score (what arb_topk_search does for ARB_DTYPE_F32). tf32 scoring is off by
 *   at most 2^-9 |q| |c| for any row, so the result is provably the exact top-k whenever the exact k-th
 *   best score clears the worst candidate's approximate score by that bound; unverified_dev[q] (int32 [Q],
 *   may be NULL) receives 0 when it does, 1 when it does not (dense near-ties around rank k).
 * mode 1: 3-term bf16 hi/lo split (scoring error ~4e-7) + exact re-score: 3x the tensor-core work and a
 *   1.5x copy of the corpus in the workspace — the fallback for the queries mode 0 left unverified.
 * corpus_max_norm: an upper bound of the corpus rows' L2 norms (1 for the unit rows of `encode`). */
size_t arb_topk_search_f32_workspace_bytes(int64_t Q, int64_t N, int32_t D, int32_t k, int32_t mode);
int arb_topk_search_f32(const float* queries_dev, const float* corpus_dev, int64_t Q, int64_t N, int32_t D, int32_t k,
                        float corpus_max_norm, float* out_scores_dev, int64_t* out_ids_dev, int64_t id_offset,
                        int32_t* unverified_dev, int32_t mode, void* workspace_dev, size_t workspace_bytes, void* stream);
/* Merge G sorted per-shard lists (e.g. the all-gathered [G,Q,k] of a row-sharded corpus). */
int arb_topk_merge(const float* scores_dev, const int64_t* ids_dev, int32_t G, int64_t Q, int32_t k,
                   float* out_scores_dev, int64_t* out_ids_dev, void* stream);
/* Schedule of arb_topk_search (process-wide; for tests and benchmarks): 0 = auto, 1 = one CTA per
 * 128-query tile, 2 = CTA pairs (cta_group::2) scoring 256 queries per corpus chunk. Auto uses
 * pairs whenever there is more than one query tile. The workspace size depends on the mode: query it
 * after setting the mode. */
int arb_set_search_mode(int32_t mode);
/* Pacing of the work units that walk the same corpus split (they keep within a few chunks of each other
 * so the split's chunks are shared through L2; process-wide, initial value 1 unless ARB_SEARCH_PACE=0).
 * Results do not depend on it; the GPU tests compare them. */
int arb_set_search_pace(int32_t on);
/* Row-sharded search moves each rank's result in ONE all-gather: a record is the rank's [Q,k]
 * float32 scores followed, at arb_topk_record_ids_offset (8-byte aligned), by its [Q,k] int64 ids —
 * pass those two addresses to arb_topk_search as out_scores_dev / out_ids_dev. arb_topk_merge_records
 * merges G records laid end to end (the all-gather output), same order rule as arb_topk_merge. */
size_t arb_topk_record_bytes(int64_t Q, int32_t k);
size_t arb_topk_record_ids_offset(int64_t Q, int32_t k);
int arb_topk_merge_records(const void* records_dev, int32_t G, int64_t Q, int32_t k, float* out_scores_dev,
                           int64_t* out_ids_dev, void* stream);
/* Peer-memory exchange for ranks of one node (NVLink / NVSwitch), replacing all-gather + merge by ONE
 * kernel per rank. Each rank allocates an exchange buffer (arb_exchange_alloc, zero-filled, a CUDA
 * allocation of its own), exports its IPC handle (64 bytes), imports its peers' and uploads the G
 * buffer addresses (its own at index `rank`) as a device array of pointers. arb_topk_exchange_merge
 * — called collectively, in the same order, by every rank — stores the local record into every
 * rank's buffer, signals, waits for the G records of this round and merges them (same order rule as
 * arb_topk_merge). slot_bytes (multiple of 8) is the per-record capacity the buffers were sized with
 * (arb_topk_exchange_bytes); records larger than a slot must take the all-gather path. */
size_t arb_topk_exchange_bytes(int32_t G, size_t slot_bytes);
int arb_exchange_alloc(size_t bytes, void** dev_ptr_out);
int arb_exchange_free(void* dev_ptr);
int arb_ipc_export(const void* dev_ptr, void* handle_out_64);
int arb_ipc_import(const void* handle_64, void** dev_ptr_out);
int arb_ipc_close(void* dev_ptr);
int arb_topk_exchange_merge(const void* local_record_dev, const void* peer_bufs_dev, int32_t rank, int32_t G, int64_t Q,
                            int32_t k, size_t slot_bytes, float* out_scores_dev, int64_t* out_ids_dev, void* stream);
/* The exchange kernel waits for its peers' records for at most ARB_EXCHANGE_TIMEOUT_MS (default 10 s).
 * If a record never arrives (a rank died or skipped the call) it returns -inf / -1 rows and marks this
 * rank's buffer; this call (which synchronises the device) then returns ARB_ERR_CUDA with the peer's
 * rank in arb_last_error(). ARB_OK otherwise. */
int arb_topk_exchange_status(const void* own_buf_dev);
/* out[i] = cos(emb[i], emb[i-1]) for fp32 rows [n, D] (out[0] = 1): the adjacent-sentence similarity
 * TextChunker._chunk_semantic computes with _cosine_similarity (text_processor.py:1547-1561, :1601-1605). */
int arb_adjacent_cosine(const float* emb_dev, int64_t n, int32_t D, float* out_dev, void* stream);
/* ------------------------------------------------------------------------------------------
 * Host-side WordPiece tokenizer (no GPU work): the string half of SentenceTransformer.encode
 * (generate_embeddings_parallel.py:146-153 hands List[str] to model.encode; the tokenizer that
 * sentence-transformers calls first is BertNormalizer + BertPreTokenizer + WordPiece + the
 * <s>..</s> / [CLS]..[SEP] template). Specification and checker: arxiv_rag_b200/tokenizer.py.
 *
 * create : the vocabulary as UTF-8 token bytes (token i = bytes[offsets[i], offsets[i+1])) with its
 *          ids; a repeated token keeps the last id.
 * encode : n UTF-8 texts (text r = bytes[offsets[r], offsets[r+1])) -> out_ids[r, 0..out_lens[r])
 *          = cls, pieces (at most max_length-2), sep; the rest of the row (out_stride >=
 *          max(max_length, 2) int32) is pad. Rows whose normalisation depends on context
 *          (U+03A3, the 26 non-Mn combining marks with a non-zero class) or that are not valid
 *          UTF-8 get out_fallback[r] = 1 and must be re-done by the caller (the Python class does).
 *          num_threads <= 0: all hardware threads. Thread-safe: a handle is read-only after create. */
int arb_tokenizer_create(const char* token_bytes, const int64_t* token_offsets, const int32_t* token_ids, int32_t n_tokens,
                         int32_t cls_id, int32_t sep_id, int32_t pad_id, int32_t unk_id, int32_t do_lower_case,
                         void** handle_out);
int arb_tokenizer_destroy(void* handle);
int arb_tokenizer_encode(void* handle, const char* text_bytes, const int64_t* text_offsets, int64_t n_texts,
                         int32_t max_length, int32_t num_threads, int32_t* out_ids, int64_t out_stride,
                         int32_t* out_lens, uint8_t* out_fallback);

/* Upper bound on the kernel launches one arb_topk_search call enqueues (the query pad copy is skipped
 * when Q is a whole number of query tiles). */
int arb_topk_search_launches(int32_t dtype);

/* ------------------------------------------------------------------------------------------
 * Kernel-level entry points (each is one launch); used by the parity tests and available to
 * callers that want to compose the encoder themselves. All pointers are device pointers;
 * `dtype` is the 16-bit format of every operand (ARB_DTYPE_BF16 or ARB_DTYPE_F16).
 * ------------------------------------------------------------------------------------------ */
/* GEMM with a LayerNorm folded into its epilogue (how the encoder avoids separate LayerNorm passes).
 * epilogue 3: C = rstd (A.B'^T - mean colsum) + bias          A rows are PRE-LayerNorm; B' = B diag(gamma),
 *          4: same, then GELU                                   colsum[n] = sum_k B'[n,k], bias = b + B beta
 *          5: C = A.B^T + bias + LayerNorm(R) (gamma, beta), and row partials of C -> stats_out
 *          6: C = A.B^T + bias + R,                          and row partials of C -> stats_out
 * Row statistics are partial sums: stats[p * M + r] = (sum, sum of squares) of 128 columns of row r as
 * float2; stats_in has parts_in parts covering width_in columns (of A's rows for 3/4, of R's for 5);
 * stats_out receives N / 128 parts (N % 128 == 0). */
int arb_gemm16_lnfold(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* bias,
                      const void* R, int64_t ldr, const float* colsum, const float* gamma, const float* beta,
                      const float* stats_in, int32_t parts_in, int32_t width_in, float* stats_out, float eps, int64_t M,
                      int32_t N, int32_t K, int32_t epilogue, int32_t dtype, void* stream);
/* Tile schedule of the arb_gemm16* kernels (process-wide; for tests and benchmarks): 0 = auto,
 * 1 = one CTA per 128x256 tile (tcgen05 cta_group::1), 2 = CTA pairs sharing a 256x256 tile
 * (cta_group::2, thread-block clusters of two), 3 = one CTA per 128x128 tile (16-bit outputs; other
 * outputs fall back to 1). Auto: pairs from 4096 rows up; below that single
 * CTAs, with 128x128 tiles when the 128x256 tiles would fill less than half of the SMs (query-time
 * batches; ARB_GEMM_NARROW=0 disables). Kernel chains of small calls are launched programmatically
 * dependent (ARB_PDL=0 disables). */
int arb_set_gemm_mode(int32_t mode);
/* Programmatic dependent launch of kernel chains (process-wide; initial value from ARB_PDL): 0 = never,
 * 1 = calls the library considers latency-bound (an encode of <= 16384 tokens), 2 = every chain.
 * Results are bit-identical in every mode; the GPU tests compare them. */
int arb_set_pdl_mode(int32_t mode);
/* C[M,N] = epi(A[M,K] . B[N,K]^T + bias[N]) (+ R[M,N]); 16-bit operands, fp32 accumulate. */
int arb_gemm16(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
               const float* bias, const void* R, int64_t ldr, int64_t M, int32_t N, int32_t K,
               int32_t epilogue, int32_t dtype, void* stream);
int arb_gemm16_f32out(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                      int64_t M, int32_t N, int32_t K, int32_t dtype, void* stream);
int arb_embed_layernorm(const int32_t* ids, const float* word_emb, const float* pos_emb,
                        const float* gamma, const float* beta, void* out16, int32_t B, int32_t S,
                        int32_t H, int32_t vocab, int32_t max_pos, int32_t pad_id, int32_t position_mode,
                        float eps, int32_t dtype, void* stream);
int arb_layernorm16(const void* x, const float* gamma, const float* beta, void* out, int64_t rows,
                    int32_t H, float eps, int32_t dtype, void* stream);
int arb_attention16(const void* qkv, const float* rel_bias, int32_t max_rel, const int32_t* mask,
                    void* ctx, int32_t B, int32_t S, int32_t heads, int32_t head_dim, int32_t dtype,
                    int32_t impl /* 0 auto, 1 mma.sync kernel, 2 tcgen05 kernel, 8 softmax warps (head dim 64, S <= 384; auto: 64 <= S <= 192), 3 experimental (ARB_ERR_UNSUPPORTED in default builds), 4 tcgen05 kernel, 16 softmax warps (auto: 192 < S <= 384) */,
                    void* stream);
int arb_pool_normalize(const void* hidden16, const int32_t* mask, float* out, int32_t B, int32_t S,
                       int32_t H, int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ARXIV_RAG_B200_H */
