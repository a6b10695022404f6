// Internal (C++) launch interface between the kernels and the C-ABI layer in capi.cu.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace arb {

// Encoder activations/weights are 16-bit in HBM: bf16 or fp16 (`fp16` flag of each launch). Both
// operands of a tcgen05 kind::f16 MMA must share the format: the instruction descriptor has separate
// A/B format fields, but a bf16 x fp16 instruction faults as illegal on B200 (tried in round 2).
typedef uint16_t h16;

enum GemmEpilogue : int {
    EPI_BIAS = 0,           // C = A.B^T + bias
    EPI_BIAS_GELU = 1,      // C = gelu_erf(A.B^T + bias)
    EPI_BIAS_RESIDUAL = 2,  // C = A.B^T + bias + R
    // LayerNorm folded into the neighbouring GEMMs (no separate LayerNorm pass):
    // the producer writes the PRE-LayerNorm rows x plus per-row (sum, sum of squares) partials; a
    // consumer GEMM whose weights carry gamma (W' = W diag(gamma), colsum c = W' 1, bias b' = b + W beta)
    // recovers LN(x) W^T + b = rstd (x W'^T - mu c) + b' in its epilogue.
    EPI_LNIN_BIAS = 3,          // C = rstd (A.B'^T - mu c) + b'                (A rows are pre-LN)
    EPI_LNIN_BIAS_GELU = 4,     // C = gelu(rstd (A.B'^T - mu c) + b')
    EPI_BIAS_LNRES_STATS = 5,   // C = A.B^T + bias + LN(R); row partials of C -> stats_out
    EPI_BIAS_RES_STATS = 6,     // C = A.B^T + bias + R;     row partials of C -> stats_out
};

// Extra operands of the folded-LayerNorm epilogues. Row statistics travel as partial sums:
// part p of row r at stats[p * M + r] = (sum, sum of squares) over 128 columns of that row.
struct LnFoldArgs {
    const float* colsum = nullptr;     // c[N]                      (EPI_LNIN_*)
    const float* gamma = nullptr;      // LayerNorm weight of R     (EPI_BIAS_LNRES_STATS)
    const float* beta = nullptr;       // LayerNorm bias of R       (EPI_BIAS_LNRES_STATS)
    const float2* stats_in = nullptr;  // partials of the rows being normalised (A rows / R rows)
    float2* stats_out = nullptr;       // partials of the output rows, N / 128 parts (EPI_*_STATS)
    int parts_in = 0;
    float inv_width_in = 0.f;          // 1 / (columns the input partials cover)
    float eps = 0.f;
};

// C[M,N] = epi(A[M,K] . B[N,K]^T); A, B, C, R 16-bit row-major; bias fp32 [N] (may be null).
int launch_gemm16(const h16* A, int64_t lda, const h16* B, int64_t ldb, h16* C, int64_t ldc,
                  const float* bias, const h16* R, int64_t ldr, int64_t M, int N, int K,
                  int epilogue, bool fp16, cudaStream_t stream);
// The folded-LayerNorm epilogues (3..6); N % 128 == 0 for the *_STATS ones.
int launch_gemm16_fold(const h16* A, int64_t lda, const h16* B, int64_t ldb, h16* C, int64_t ldc,
                       const float* bias, const h16* R, int64_t ldr, int64_t M, int N, int K,
                       int epilogue, const LnFoldArgs& fold, bool fp16, cudaStream_t stream);
// Same main loop, fp32 output, no epilogue math (used by the kernel parity tests).
int launch_gemm16_f32out(const h16* A, int64_t lda, const h16* B, int64_t ldb, float* C,
                         int64_t ldc, int64_t M, int N, int K, bool fp16, cudaStream_t stream);

// word_emb[ids] + pos_emb[position_ids(ids)] -> LayerNorm -> 16-bit hidden [B*S, H]
int launch_embed_ln(const int32_t* ids, const float* word_emb, const float* pos_emb,
                    const float* gamma, const float* beta, h16* out, int B, int S, int H, int vocab,
                    int max_pos, int pad_id, int pos_mode /* 0 MPNet pad-aware, 1 absolute */, float eps, bool fp16,
                    int* err_flag_dev /* nullable: [0]=1, [1]=id, [2]=token index on an out-of-range id */,
                    cudaStream_t stream);
// out = LayerNorm(x) row-wise, x 16-bit [rows, H] (already holds GEMM output + residual).
int launch_layernorm(const h16* x, const float* gamma, const float* beta, h16* out, int64_t rows,
                     int H, float eps, bool fp16, cudaStream_t stream);
// masked mean over tokens then L2 normalise: hidden [B,S,H], mask int32 [B,S] -> fp32 [B,H]
// pooling + L2 normalisation with the last LayerNorm applied on the fly from the pre-LN rows and their
// row partials ([parts][B*S] (sum, sum of squares)): the folded forward's final step
int launch_pool_ln_normalize(const h16* pre, const float2* stats, int parts, const float* gamma, const float* beta,
                             float eps, const int32_t* mask, float* out, int B, int S, int H, bool fp16,
                             cudaStream_t stream);
int launch_pool_normalize(const h16* hidden, const int32_t* mask, float* out, int B, int S, int H,
                          bool fp16, cudaStream_t stream);
// out[i] = cos(emb[i], emb[i-1]) (out[0] = 1): the adjacent-sentence similarity of semantic chunking
int launch_adjacent_cosine(const float* emb, float* out, int64_t n, int D, cudaStream_t stream);
// softmax(q.k^T/sqrt(dh) + rel_bias[h][j-i] + mask) . v for every (batch, head);
// qkv [B*S, 3*H] (q | k | v column blocks), rel_bias fp32 [heads, 2*max_rel-1]
// (entry r <-> j-i = r-(max_rel-1)), mask int32 [B,S]; ctx [B*S, H].
int launch_attention(const h16* qkv, const float* rel_bias, int max_rel, const int32_t* mask,
                     h16* ctx, int B, int S, int heads, int dh, bool fp16, int impl, cudaStream_t stream);
// the implementations behind it: mma.sync flash kernel (any S <= 768; head dim 32) and the tcgen05/TMEM
// kernels (attention_tc2.cu: sub-block pipelined, the default; attention_tc.cu: the round-1 schedule)
int launch_attention_mma(const h16* qkv, const float* rel_bias, int max_rel, const int32_t* mask,
                         h16* ctx, int B, int S, int heads, int dh, bool fp16, cudaStream_t stream);
bool attention_tc2_supported(int S, int dh);
int launch_attention_tc2(const h16* qkv, const float* rel_bias, int max_rel, const int32_t* mask,
                         h16* ctx, int B, int S, int heads, int dh, bool fp16, cudaStream_t stream);
bool attention_tc_supported(int S, int dh);
// attention_tc3.cu: the same schedule with 16 softmax warps (two column threads per query row and key half)
bool attention_tc3_supported(int S, int dh);
int launch_attention_tc3(const h16* qkv, const float* rel_bias, int max_rel, const int32_t* mask,
                         h16* ctx, int B, int S, int heads, int dh, bool fp16, cudaStream_t stream);
int launch_attention_tc(const h16* qkv, const float* rel_bias, int max_rel, const int32_t* mask,
                        h16* ctx, int B, int S, int heads, int dh, bool fp16, cudaStream_t stream);

// Search: fused score GEMM + per-query running top-k over a bf16 corpus.
size_t search_workspace_bytes(int64_t Q, int64_t N, int D, int k);
int launch_search_bf16(const __nv_bfloat16* q, const __nv_bfloat16* corpus, int64_t Q, int64_t N,
                       int D, int k, float* out_scores, int64_t* out_ids, int64_t id_offset,
                       void* workspace, size_t ws_bytes, cudaStream_t stream);
// fp32 operands. mode 0: one kind::tf32 pass over the stored rows, k + 22 candidates re-scored in exact
// fp32, and a per-query verdict whether the result is provably the exact top-k (unverified[q] = 1 -> the
// caller re-runs that query with mode 1). mode 1: 3-term bf16 hi/lo split (error ~4e-7) + re-score.
// corpus_max_norm: upper bound of the corpus rows' L2 norms (1 for unit rows), used by the verdict.
size_t search_f32_workspace_bytes(int64_t Q, int64_t N, int D, int k, int mode);
int launch_search_f32(const float* q, const float* corpus, int64_t Q, int64_t N, int D, int k,
                      float corpus_max_norm, float* out_scores, int64_t* out_ids, int64_t id_offset,
                      int32_t* unverified, int mode, void* workspace, size_t ws_bytes, cudaStream_t stream);
// k-way merge of G per-shard top-k lists: [G,Q,k] -> [Q,k], order (score desc, id asc).
int launch_topk_merge(const float* scores, const int64_t* ids, int G, int64_t Q, int k,
                      float* out_scores, int64_t* out_ids, cudaStream_t stream);
// Search schedule override: 0 auto, 1 one CTA per 128-query tile, 2 CTA pairs per 256 queries.
void set_search_mode(int mode);
bool& search_pace_ref();
// GEMM schedule override for tests/benchmarks: 0 auto, 1 one CTA per 128x256 tile, 2 CTA pairs per 256x256 tile.
void set_gemm_mode(int mode);
// One rank's [Q,k] result as a single buffer (fp32 scores, then 8-byte aligned int64 ids), and the
// merge of G such records laid end to end (what one all-gather of the records produces).
size_t topk_record_ids_offset(int64_t Q, int k);
size_t topk_record_bytes(int64_t Q, int k);
int launch_topk_merge_records(const void* records, int G, int64_t Q, int k, float* out_scores,
                              int64_t* out_ids, cudaStream_t stream);
// Peer-memory exchange (one kernel: push records to all ranks' mapped buffers, flag, wait, merge).
size_t topk_exchange_bytes(int G, size_t slot_bytes);
int topk_exchange_status(const void* own_buf_dev);
int launch_topk_exchange_merge(const void* local_record, void* const* peer_bufs_dev, int rank, int G, int64_t Q,
                               int k, size_t slot_bytes, float* out_scores, int64_t* out_ids, cudaStream_t stream);

}  // namespace arb
