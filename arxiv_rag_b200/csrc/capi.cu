// extern "C" surface declared in include/arxiv_rag_b200.h: the MPNet encoder handle, the
// search entry points and thin per-kernel wrappers. No torch types, no exceptions.
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/arxiv_rag_b200.h"
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace arb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

// public 16-bit dtype code of the kernel-level entry points -> fp16 flag
static int dtype16(int32_t dtype, bool* fp16) {
    ARB_REQUIRE(dtype == ARB_DTYPE_BF16 || dtype == ARB_DTYPE_F16, "dtype %d must be ARB_DTYPE_BF16 or ARB_DTYPE_F16", dtype);
    *fp16 = dtype == ARB_DTYPE_F16;
    return ARB_OK;
}

// ---- MPNetEncoder.relative_position_bucket (modeling_mpnet.py:343-360), float32 like torch.
static int relative_bucket(int relative_position, int num_buckets, int max_distance) {
    int n = -relative_position;
    const int nb = num_buckets / 2;
    int ret = 0;
    if (n < 0) {
        ret += nb;
        n = -n;
    }
    const int max_exact = nb / 2;
    if (n < max_exact) return ret + n;
    const float a = logf(static_cast<float>(n) / static_cast<float>(max_exact));
    const float b = a / static_cast<float>(log(static_cast<double>(max_distance) / max_exact));
    const float c = b * static_cast<float>(nb - max_exact);
    int v = max_exact + static_cast<int>(c);
    if (v > nb - 1) v = nb - 1;
    return ret + v;
}

struct LayerDev {
    h16 *w_qkv, *w_o, *w_in, *w_out;  // [3H,H] [H,H] [I,H] [H,I]
    float *b_qkv, *b_o, *b_in, *b_out;
    float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    // folded-LayerNorm path: column sums of the gamma-scaled weights (w_qkv of layers >= 1, w_in)
    float *c_qkv = nullptr, *c_in = nullptr;
};

struct Mpnet {
    ArbMpnetConfig cfg;
    int device = 0;
    int64_t max_tokens = 0;
    int max_seq = 0;
    int64_t bytes = 0;
    std::vector<void*> allocs;
    float *word_emb = nullptr, *pos_emb = nullptr, *emb_g = nullptr, *emb_b = nullptr;
    float* rel_bias = nullptr;  // [heads, 2*max_seq-1], entry r <-> j-i = r-(max_seq-1)
    std::vector<LayerDev> layers;     // weights in the handle's 16-bit format
    std::vector<LayerDev> layers_f16; // bf16 handles: an fp16 copy for short batches (see short_f16)
    h16 *h = nullptr, *h1 = nullptr, *tmp = nullptr, *ctx = nullptr, *qkv = nullptr, *ffn = nullptr;
    bool fp16 = false;     // 16-bit format of weights + activations (one format per tcgen05 MMA)
    // compute_dtype ARB_DTYPE_BF16: batches padded to fewer than kShortSeq tokens run in fp16 (an
    // fp16 copy of the weights, the same activation buffers). A row of a few tokens has no
    // mean-pool averaging over its bf16 rounding noise and cannot reach cosine 0.9999 otherwise
    // (tools/rounding_budget.py). ARB_DTYPE_BF16_PURE switches this off (A/B baseline).
    bool short_f16 = false;
    int* status_host = nullptr;  // host-mapped status word [0]=error, [1]=token id, [2]=token index
    int* status_dev = nullptr;
    // LayerNorm folded into the neighbouring GEMMs: producers write pre-LN rows + row partials,
    // consumers carry gamma in their weights and finish the normalisation in the epilogue.
    bool fold_ln = false;
    float2 *stats_x = nullptr, *stats_y = nullptr;  // [H/128][max_tokens] (sum, sum of squares)

    template <typename T>
    int alloc(T** p, size_t count) {
        void* d = nullptr;
        ARB_CHECK_CUDA(cudaMalloc(&d, count * sizeof(T)));
        allocs.push_back(d);
        bytes += static_cast<int64_t>(count * sizeof(T));
        *p = static_cast<T*>(d);
        return ARB_OK;
    }
    int upload_f32(float** dst, const float* src, size_t count) {
        ARB_REQUIRE(src != nullptr, "mpnet_create: missing weight array");
        if (int rc = alloc(dst, count)) return rc;
        ARB_CHECK_CUDA(cudaMemcpy(*dst, src, count * sizeof(float), cudaMemcpyHostToDevice));
        return ARB_OK;
    }
    // fp32 host values -> 16-bit (round-to-nearest-even) device
    int upload16(h16* dst, const float* src, size_t count, bool f16) {
        ARB_REQUIRE(src != nullptr, "mpnet_create: missing weight matrix");
        std::vector<h16> tmpv(count);
        if (f16) {
            for (size_t i = 0; i < count; ++i) tmpv[i] = __half_as_ushort(__float2half_rn(src[i]));
        } else {
            for (size_t i = 0; i < count; ++i) tmpv[i] = __bfloat16_as_ushort(__float2bfloat16_rn(src[i]));
        }
        ARB_CHECK_CUDA(cudaMemcpy(dst, tmpv.data(), count * sizeof(h16), cudaMemcpyHostToDevice));
        return ARB_OK;
    }
    // W'[n,k] = W[n,k] * gamma[k] rounded to 16 bit -> dst; colsum[n] = sum_k W'[n,k] (of the ROUNDED
    // values, so that x W'^T - mean * colsum cancels exactly); bias_out[n] = bias[n] + sum_k W[n,k] beta[k]
    int upload16_folded(h16* dst, float* colsum_dev, float* bias_dev, const float* w, const float* bias,
                        const float* gamma, const float* beta, size_t N, size_t K, bool f16) {
        ARB_REQUIRE(w && bias && gamma && beta, "mpnet_create: missing weight array");
        std::vector<h16> t16(N * K);
        std::vector<float> cs(N), bo(N);
        for (size_t n = 0; n < N; ++n) {
            double c = 0.0, b = bias[n];
            for (size_t k = 0; k < K; ++k) {
                const float wf = w[n * K + k] * gamma[k];
                float back;
                if (f16) {
                    const __half hv = __float2half_rn(wf);
                    t16[n * K + k] = __half_as_ushort(hv);
                    back = __half2float(hv);
                } else {
                    const __nv_bfloat16 bv = __float2bfloat16_rn(wf);
                    t16[n * K + k] = __bfloat16_as_ushort(bv);
                    back = __bfloat162float(bv);
                }
                c += back;
                b += static_cast<double>(w[n * K + k]) * beta[k];
            }
            cs[n] = static_cast<float>(c);
            bo[n] = static_cast<float>(b);
        }
        ARB_CHECK_CUDA(cudaMemcpy(dst, t16.data(), N * K * sizeof(h16), cudaMemcpyHostToDevice));
        ARB_CHECK_CUDA(cudaMemcpy(colsum_dev, cs.data(), N * 4, cudaMemcpyHostToDevice));
        ARB_CHECK_CUDA(cudaMemcpy(bias_dev, bo.data(), N * 4, cudaMemcpyHostToDevice));
        return ARB_OK;
    }
    ~Mpnet() {
        for (void* p : allocs) cudaFree(p);
        if (status_host) cudaFreeHost(status_host);
    }
};

static int mpnet_build_layers(Mpnet* m, const ArbMpnetWeights* w, std::vector<LayerDev>& layers, bool f16);

static int mpnet_build(Mpnet* m, const ArbMpnetWeights* w) {
    const ArbMpnetConfig& c = m->cfg;
    const size_t H = c.hidden_size, I = c.intermediate_size;
    if (int rc = m->upload_f32(&m->word_emb, w->word_embeddings, static_cast<size_t>(c.vocab_size) * H)) return rc;
    if (int rc = m->upload_f32(&m->pos_emb, w->position_embeddings, static_cast<size_t>(c.max_position_embeddings) * H)) return rc;
    if (int rc = m->upload_f32(&m->emb_g, w->emb_ln_g, H)) return rc;
    if (int rc = m->upload_f32(&m->emb_b, w->emb_ln_b, H)) return rc;
    // expand the bucketed relative bias once: it depends only on j-i (modeling_mpnet.py:324-341)
    // (a BERT-style encoder has none: relative_attention_num_buckets == 0 -> attention without bias)
    if (c.relative_attention_num_buckets > 0) {
        ARB_REQUIRE(w->relative_attention_bias != nullptr, "mpnet_create: missing relative_attention_bias");
        const int P = m->max_seq, W = 2 * P - 1;
        std::vector<float> tbl(static_cast<size_t>(c.num_heads) * W);
        for (int r = 0; r < W; ++r) {
            const int bucket = relative_bucket(r - (P - 1), c.relative_attention_num_buckets, 128);
            for (int hd = 0; hd < c.num_heads; ++hd)
                tbl[static_cast<size_t>(hd) * W + r] = w->relative_attention_bias[static_cast<size_t>(bucket) * c.num_heads + hd];
        }
        if (int rc = m->upload_f32(&m->rel_bias, tbl.data(), tbl.size())) return rc;
    }
    if (int rc = mpnet_build_layers(m, w, m->layers, m->fp16)) return rc;
    if (m->short_f16)
        if (int rc = mpnet_build_layers(m, w, m->layers_f16, true)) return rc;
    const size_t T = static_cast<size_t>(m->max_tokens);
    if (int rc = m->alloc(&m->h, T * H)) return rc;
    if (int rc = m->alloc(&m->h1, T * H)) return rc;
    if (int rc = m->alloc(&m->tmp, T * H)) return rc;
    if (int rc = m->alloc(&m->ctx, T * H)) return rc;
    if (int rc = m->alloc(&m->qkv, T * 3 * H)) return rc;
    if (int rc = m->alloc(&m->ffn, T * I)) return rc;
    if (m->fold_ln) {
        if (int rc = m->alloc(&m->stats_x, T * (H / 128))) return rc;
        if (int rc = m->alloc(&m->stats_y, T * (H / 128))) return rc;
    }
    return ARB_OK;
}

// One set of per-layer device weights in the fp16 (f16 = true) or bf16 format.
static int mpnet_build_layers(Mpnet* m, const ArbMpnetWeights* w, std::vector<LayerDev>& layers, bool f16) {
    const ArbMpnetConfig& c = m->cfg;
    const size_t H = c.hidden_size, I = c.intermediate_size;
    layers.resize(c.num_layers);
    for (int l = 0; l < c.num_layers; ++l) {
        const ArbMpnetLayerWeights& lw = w->layers[l];
        LayerDev& d = layers[l];
        if (int rc = m->alloc(&d.w_qkv, 3 * H * H)) return rc;
        if (int rc = m->alloc(&d.b_qkv, 3 * H)) return rc;
        ARB_REQUIRE(lw.q_b && lw.k_b && lw.v_b, "mpnet_create: missing q/k/v bias (layer %d)", l);
        if (m->fold_ln && l > 0) {
            // the q/k/v projections of layer l read LN2 of layer l-1: carry its gamma/beta
            const ArbMpnetLayerWeights& pw = w->layers[l - 1];
            if (int rc = m->alloc(&d.c_qkv, 3 * H)) return rc;
            const float* ws[3] = {lw.q_w, lw.k_w, lw.v_w};
            const float* bs[3] = {lw.q_b, lw.k_b, lw.v_b};
            for (int t = 0; t < 3; ++t)
                if (int rc = m->upload16_folded(d.w_qkv + t * H * H, d.c_qkv + t * H, d.b_qkv + t * H, ws[t], bs[t],
                                                pw.out_ln_g, pw.out_ln_b, H, H, f16))
                    return rc;
        } else {
            if (int rc = m->upload16(d.w_qkv, lw.q_w, H * H, f16)) return rc;
            if (int rc = m->upload16(d.w_qkv + H * H, lw.k_w, H * H, f16)) return rc;
            if (int rc = m->upload16(d.w_qkv + 2 * H * H, lw.v_w, H * H, f16)) return rc;
            ARB_CHECK_CUDA(cudaMemcpy(d.b_qkv, lw.q_b, H * 4, cudaMemcpyHostToDevice));
            ARB_CHECK_CUDA(cudaMemcpy(d.b_qkv + H, lw.k_b, H * 4, cudaMemcpyHostToDevice));
            ARB_CHECK_CUDA(cudaMemcpy(d.b_qkv + 2 * H, lw.v_b, H * 4, cudaMemcpyHostToDevice));
        }
        if (int rc = m->alloc(&d.w_o, H * H)) return rc;
        if (int rc = m->upload16(d.w_o, lw.o_w, H * H, f16)) return rc;
        if (int rc = m->upload_f32(&d.b_o, lw.o_b, H)) return rc;
        if (int rc = m->upload_f32(&d.ln1_g, lw.attn_ln_g, H)) return rc;
        if (int rc = m->upload_f32(&d.ln1_b, lw.attn_ln_b, H)) return rc;
        if (int rc = m->alloc(&d.w_in, I * H)) return rc;
        if (m->fold_ln) {  // the FFN up-projection reads LN1 of this layer
            if (int rc = m->alloc(&d.c_in, I)) return rc;
            if (int rc = m->alloc(&d.b_in, I)) return rc;
            if (int rc = m->upload16_folded(d.w_in, d.c_in, d.b_in, lw.ffn_in_w, lw.ffn_in_b, lw.attn_ln_g, lw.attn_ln_b, I, H, f16))
                return rc;
        } else {
            if (int rc = m->upload16(d.w_in, lw.ffn_in_w, I * H, f16)) return rc;
            if (int rc = m->upload_f32(&d.b_in, lw.ffn_in_b, I)) return rc;
        }
        if (int rc = m->alloc(&d.w_out, H * I)) return rc;
        if (int rc = m->upload16(d.w_out, lw.ffn_out_w, H * I, f16)) return rc;
        if (int rc = m->upload_f32(&d.b_out, lw.ffn_out_b, H)) return rc;
        if (int rc = m->upload_f32(&d.ln2_g, lw.out_ln_g, H)) return rc;
        if (int rc = m->upload_f32(&d.ln2_b, lw.out_ln_b, H)) return rc;
    }
    return ARB_OK;
}

constexpr int kShortSeq = 32;

constexpr int64_t kPdlMaxTokens = 16384;

int& pdl_mode_ref() {
    static int mode = []() {
        const char* e = getenv("ARB_PDL");
        return e && e[0] >= '0' && e[0] <= '2' ? e[0] - '0' : 1;
    }();
    return mode;
}

static int mpnet_encode(Mpnet* m, const int32_t* ids, const int32_t* mask, int B, int S, float* out,
                        cudaStream_t st) {
    const ArbMpnetConfig& c = m->cfg;
    const int H = c.hidden_size, I = c.intermediate_size;
    const int64_t T = static_cast<int64_t>(B) * S;
    const bool use_f16_copy = m->short_f16 && S < kShortSeq;
    const bool a16 = m->fp16 || use_f16_copy;
    const bool fmt = a16;  // one 16-bit format for every operand of this call
    const std::vector<LayerDev>& layers = use_f16_copy ? m->layers_f16 : m->layers;
    // Query-time batches: the forward is 62 short kernels; launch them programmatically dependent so
    // each one's set-up and weight prefetch overlap its predecessor (common.cuh: pdl_scope).
    // ARB_PDL=0 never, 2 always, default: up to kPdlMaxTokens tokens.
    pdl_scope pdl(T <= kPdlMaxTokens);
    int rc;
    if ((rc = launch_embed_ln(ids, m->word_emb, m->pos_emb, m->emb_g, m->emb_b, m->h, B, S, H,
                              c.vocab_size, c.max_position_embeddings, c.pad_token_id, c.position_mode,
                              c.layer_norm_eps, a16, m->status_dev, st)))
        return rc;
    if (m->fold_ln) {
        // No LayerNorm passes between the GEMMs: x (pre-LN, in `tmp`) and y (pre-LN, in `h1`) travel
        // with their row partials; every consumer finishes the normalisation in its epilogue.
        LnFoldArgs f;
        f.parts_in = H / 128;
        f.inv_width_in = 1.0f / static_cast<float>(H);
        f.eps = c.layer_norm_eps;
        for (int l = 0; l < c.num_layers; ++l) {
            const LayerDev& d = layers[l];
            LnFoldArgs fq = f, fo = f, fu = f, fd = f;
            if (l == 0) {  // the embedding LayerNorm output is already normalised
                if ((rc = launch_gemm16(m->h, H, d.w_qkv, H, m->qkv, 3 * H, d.b_qkv, nullptr, 0, T, 3 * H, H, EPI_BIAS, fmt, st))) return rc;
            } else {
                fq.colsum = d.c_qkv;
                fq.stats_in = m->stats_x;
                if ((rc = launch_gemm16_fold(m->tmp, H, d.w_qkv, H, m->qkv, 3 * H, d.b_qkv, nullptr, 0, T, 3 * H, H, EPI_LNIN_BIAS, fq, fmt, st))) return rc;
            }
            if ((rc = launch_attention(m->qkv, m->rel_bias, m->max_seq, mask, m->ctx, B, S, c.num_heads, H / c.num_heads, a16, 0, st))) return rc;
            // y = ctx Wo^T + bo + LN2_{l-1}(x)   (layer 0: + h), row partials of y
            fo.stats_out = m->stats_y;
            if (l == 0) {
                if ((rc = launch_gemm16_fold(m->ctx, H, d.w_o, H, m->h1, H, d.b_o, m->h, H, T, H, H, EPI_BIAS_RES_STATS, fo, fmt, st))) return rc;
            } else {
                fo.gamma = layers[l - 1].ln2_g;
                fo.beta = layers[l - 1].ln2_b;
                fo.stats_in = m->stats_x;
                if ((rc = launch_gemm16_fold(m->ctx, H, d.w_o, H, m->h1, H, d.b_o, m->tmp, H, T, H, H, EPI_BIAS_LNRES_STATS, fo, fmt, st))) return rc;
            }
            // ffn = gelu(LN1(y) W1^T + b1)
            fu.colsum = d.c_in;
            fu.stats_in = m->stats_y;
            if ((rc = launch_gemm16_fold(m->h1, H, d.w_in, H, m->ffn, I, d.b_in, nullptr, 0, T, I, H, EPI_LNIN_BIAS_GELU, fu, fmt, st))) return rc;
            // x = ffn W2^T + b2 + LN1(y), row partials of x
            fd.gamma = d.ln1_g;
            fd.beta = d.ln1_b;
            fd.stats_in = m->stats_y;
            fd.stats_out = m->stats_x;
            if ((rc = launch_gemm16_fold(m->ffn, I, d.w_out, I, m->tmp, H, d.b_out, m->h1, H, T, H, I, EPI_BIAS_LNRES_STATS, fd, fmt, st))) return rc;
        }
        const LayerDev& last = layers[c.num_layers - 1];
        // the last LayerNorm rides in the pooling kernel (pre-LN rows + their row partials)
        return launch_pool_ln_normalize(m->tmp, m->stats_x, H / 128, last.ln2_g, last.ln2_b, c.layer_norm_eps, mask, out, B, S, H,
                                        a16, st);
    }
    for (int l = 0; l < c.num_layers; ++l) {
        const LayerDev& d = layers[l];
        // q,k,v projections as one [T,H] x [3H,H]^T GEMM (modeling_mpnet.py:145-159)
        if ((rc = launch_gemm16(m->h, H, d.w_qkv, H, m->qkv, 3 * H, d.b_qkv, nullptr, 0, T, 3 * H, H, EPI_BIAS, fmt, st))) return rc;
        // softmax(qk^T/8 + position_bias + mask) v (:162-177)
        if ((rc = launch_attention(m->qkv, m->rel_bias, m->max_seq, mask, m->ctx, B, S, c.num_heads, H / c.num_heads, a16, 0, st))) return rc;
        // o-projection + residual (:183), then the post-LN (:210) as a LayerNorm pass
        if ((rc = launch_gemm16(m->ctx, H, d.w_o, H, m->tmp, H, d.b_o, m->h, H, T, H, H, EPI_BIAS_RESIDUAL, fmt, st))) return rc;
        if ((rc = launch_layernorm(m->tmp, d.ln1_g, d.ln1_b, m->h1, T, H, c.layer_norm_eps, a16, st))) return rc;
        // FFN: GELU fused in the up-projection epilogue (:225-228), residual in the down-projection (:239-243)
        if ((rc = launch_gemm16(m->h1, H, d.w_in, H, m->ffn, I, d.b_in, nullptr, 0, T, I, H, EPI_BIAS_GELU, fmt, st))) return rc;
        if ((rc = launch_gemm16(m->ffn, I, d.w_out, I, m->tmp, H, d.b_out, m->h1, H, T, H, I, EPI_BIAS_RESIDUAL, fmt, st))) return rc;
        if ((rc = launch_layernorm(m->tmp, d.ln2_g, d.ln2_b, m->h, T, H, c.layer_norm_eps, a16, st))) return rc;
    }
    return launch_pool_normalize(m->h, mask, out, B, S, H, a16, st);
}

}  // namespace arb

using namespace arb;

extern "C" {

const char* arb_last_error(void) { return get_error(); }
int arb_abi_version(void) { return 2; }

int arb_mpnet_relative_bucket(int32_t relative_position, int32_t num_buckets, int32_t max_distance) {
    return relative_bucket(relative_position, num_buckets, max_distance);
}

int arb_mpnet_create(const ArbMpnetConfig* cfg, const ArbMpnetWeights* weights, int64_t max_tokens,
                     int32_t max_seq, int32_t device, void** handle) {
    ARB_REQUIRE(cfg && weights && handle, "mpnet_create: null argument");
    ARB_REQUIRE(weights->layers != nullptr, "mpnet_create: null layer array");
    ARB_REQUIRE(cfg->hidden_size > 0 && cfg->num_heads > 0 && cfg->hidden_size % cfg->num_heads == 0,
                "mpnet_create: bad hidden/heads %d/%d", cfg->hidden_size, cfg->num_heads);
    ARB_REQUIRE(cfg->hidden_size / cfg->num_heads == 64 || cfg->hidden_size / cfg->num_heads == 32,
                "mpnet_create: head dim %d unsupported (64 or 32)", cfg->hidden_size / cfg->num_heads);
    ARB_REQUIRE(cfg->position_mode == 0 || cfg->position_mode == 1, "mpnet_create: position_mode %d must be 0 or 1",
                cfg->position_mode);
    ARB_REQUIRE(cfg->hidden_size % 128 == 0 && cfg->hidden_size <= 1024, "mpnet_create: hidden size %d unsupported", cfg->hidden_size);
    ARB_REQUIRE(cfg->intermediate_size % 32 == 0, "mpnet_create: intermediate size %d unsupported", cfg->intermediate_size);
    ARB_REQUIRE(cfg->compute_dtype == ARB_DTYPE_BF16 || cfg->compute_dtype == ARB_DTYPE_F16 ||
                    cfg->compute_dtype == ARB_DTYPE_BF16_PURE,
                "mpnet_create: compute_dtype %d must be ARB_DTYPE_F16, ARB_DTYPE_BF16 or ARB_DTYPE_BF16_PURE", cfg->compute_dtype);
    ARB_REQUIRE(cfg->num_layers > 0 && cfg->vocab_size > 0 && cfg->max_position_embeddings > 2,
                "mpnet_create: bad layer/vocab/position counts");
    ARB_REQUIRE(max_tokens > 0 && max_seq > 0 && max_seq <= 768, "mpnet_create: bad max_tokens=%lld / max_seq=%d",
                (long long)max_tokens, max_seq);
    ARB_REQUIRE((cfg->position_mode == 1 ? max_seq - 1 : max_seq + cfg->pad_token_id) < cfg->max_position_embeddings,
                "mpnet_create: max_seq %d exceeds the position table (%d rows)", max_seq, cfg->max_position_embeddings);
    int ndev = 0;
    ARB_CHECK_CUDA(cudaGetDeviceCount(&ndev));
    ARB_REQUIRE(device >= 0 && device < ndev, "mpnet_create: no CUDA device %d (found %d) — this library has no CPU path", device, ndev);
    ARB_CHECK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    ARB_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("mpnet_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return ARB_ERR_UNSUPPORTED;
    }
    Mpnet* m = new (std::nothrow) Mpnet();
    ARB_REQUIRE(m != nullptr, "mpnet_create: out of host memory");
    m->cfg = *cfg;
    m->fp16 = cfg->compute_dtype == ARB_DTYPE_F16;
    m->short_f16 = cfg->compute_dtype == ARB_DTYPE_BF16;
    // The LayerNorms are folded into the neighbouring GEMM epilogues
    // (EPI_LNIN_* / EPI_*_STATS, kernels.h). ARB_FOLD_LN=0 keeps the GEMM + LayerNorm-pass path
    // (the A/B baseline; both are parity-tested).
    const char* fold_env = getenv("ARB_FOLD_LN");
    m->fold_ln = !(fold_env && fold_env[0] == '0');
    m->device = device;
    m->max_tokens = max_tokens;
    m->max_seq = max_seq;
    int rc = mpnet_build(m, weights);
    if (rc == ARB_OK) {
        // status word the kernels write on a data error (out-of-range token id): pinned, host-mapped
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&m->status_host), 4 * sizeof(int), cudaHostAllocMapped);
        if (e == cudaSuccess) e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&m->status_dev), m->status_host, 0);
        if (e != cudaSuccess) {
            set_error("mpnet_create: status word allocation failed: %s", cudaGetErrorString(e));
            rc = ARB_ERR_CUDA;
        } else {
            memset(m->status_host, 0, 4 * sizeof(int));
        }
    }
    if (rc) {
        delete m;
        return rc;
    }
    ARB_CHECK_CUDA(cudaDeviceSynchronize());
    *handle = m;
    return ARB_OK;
}

int arb_mpnet_destroy(void* handle) {
    if (handle) delete static_cast<Mpnet*>(handle);
    return ARB_OK;
}

int64_t arb_mpnet_device_bytes(void* handle) { return handle ? static_cast<Mpnet*>(handle)->bytes : 0; }

int arb_mpnet_launches_per_encode(void* handle) {
    if (!handle) return 0;
    const Mpnet* m = static_cast<Mpnet*>(handle);
    if (m->fold_ln) return 2 + 5 * m->cfg.num_layers;  // embed, 5 per layer, pool (with the last LayerNorm)
    return 2 + 7 * m->cfg.num_layers;
}

int arb_mpnet_encode(void* handle, const int32_t* ids_dev, const int32_t* mask_dev, int32_t B,
                     int32_t S, float* out_dev, void* stream) {
    ARB_REQUIRE(handle && ids_dev && mask_dev && out_dev, "mpnet_encode: null argument");
    Mpnet* m = static_cast<Mpnet*>(handle);
    ARB_REQUIRE(B > 0 && S > 0, "mpnet_encode: empty batch B=%d S=%d", B, S);
    ARB_REQUIRE(S <= m->max_seq, "mpnet_encode: S=%d exceeds max_seq=%d", S, m->max_seq);
    ARB_REQUIRE(static_cast<int64_t>(B) * S <= m->max_tokens, "mpnet_encode: B*S=%lld exceeds max_tokens=%lld",
                (long long)B * S, (long long)m->max_tokens);
    int cur = -1;
    ARB_CHECK_CUDA(cudaGetDevice(&cur));
    ARB_REQUIRE(cur == m->device, "mpnet_encode: the handle lives on device %d but the current device is %d", m->device, cur);
    return mpnet_encode(m, ids_dev, mask_dev, B, S, out_dev, static_cast<cudaStream_t>(stream));
}

int arb_mpnet_short_seq(void* handle) {
    return handle && static_cast<Mpnet*>(handle)->short_f16 ? kShortSeq : 0;
}

int arb_mpnet_status(void* handle) {
    ARB_REQUIRE(handle != nullptr, "mpnet_status: null handle");
    Mpnet* m = static_cast<Mpnet*>(handle);
    volatile int* st = m->status_host;
    if (st[0] == 0) return ARB_OK;
    set_error("mpnet_encode: token id %d at token index %d is outside the vocabulary [0, %d)", st[1], st[2], m->cfg.vocab_size);
    st[0] = 0;
    return ARB_ERR_INVALID;
}

size_t arb_topk_search_workspace_bytes(int32_t dtype, int64_t Q, int64_t N, int32_t D, int32_t k) {
    if (dtype == ARB_DTYPE_BF16) return search_workspace_bytes(Q, N, D, k);
    if (dtype == ARB_DTYPE_F32) return search_f32_workspace_bytes(Q, N, D, k, 0);
    return 0;
}

size_t arb_topk_search_f32_workspace_bytes(int64_t Q, int64_t N, int32_t D, int32_t k, int32_t mode) {
    return search_f32_workspace_bytes(Q, N, D, k, mode);
}

int arb_topk_search_f32(const float* queries_dev, const float* corpus_dev, int64_t Q, int64_t N, int32_t D, int32_t k,
                        float corpus_max_norm, float* out_scores_dev, int64_t* out_ids_dev, int64_t id_offset,
                        int32_t* unverified_dev, int32_t mode, void* workspace_dev, size_t workspace_bytes, void* stream) {
    ARB_REQUIRE(corpus_max_norm > 0.f && corpus_max_norm < 1e30f, "search_f32: corpus_max_norm must be a positive bound of the row norms");
    return launch_search_f32(queries_dev, corpus_dev, Q, N, D, k, corpus_max_norm, out_scores_dev, out_ids_dev, id_offset,
                             unverified_dev, mode, workspace_dev, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int arb_topk_search(const void* queries_dev, const void* corpus_dev, int32_t dtype, int64_t Q,
                    int64_t N, int32_t D, int32_t k, float* out_scores_dev, int64_t* out_ids_dev,
                    int64_t id_offset, void* workspace_dev, size_t workspace_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == ARB_DTYPE_BF16)
        return launch_search_bf16(static_cast<const __nv_bfloat16*>(queries_dev),
                                  static_cast<const __nv_bfloat16*>(corpus_dev), Q, N, D, k,
                                  out_scores_dev, out_ids_dev, id_offset, workspace_dev, workspace_bytes, st);
    if (dtype == ARB_DTYPE_F32)
        return launch_search_f32(static_cast<const float*>(queries_dev), static_cast<const float*>(corpus_dev),
                                 Q, N, D, k, 1.0f, out_scores_dev, out_ids_dev, id_offset, nullptr, 0, workspace_dev,
                                 workspace_bytes, st);
    set_error("topk_search: unknown dtype %d", dtype);
    return ARB_ERR_INVALID;
}

// search kernel + split merge (+ the query pad copy when Q is not a whole number of tiles; + the re-score for fp32)
int arb_topk_search_launches(int32_t dtype) { return dtype == ARB_DTYPE_F32 ? 4 : 3; }

int arb_topk_merge(const float* scores_dev, const int64_t* ids_dev, int32_t G, int64_t Q, int32_t k,
                   float* out_scores_dev, int64_t* out_ids_dev, void* stream) {
    return launch_topk_merge(scores_dev, ids_dev, G, Q, k, out_scores_dev, out_ids_dev,
                             static_cast<cudaStream_t>(stream));
}



int arb_gemm16_lnfold(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* bias,
                      const void* R, int64_t ldr, const float* colsum, const float* gamma, const float* beta,
                      const float* stats_in, int32_t parts_in, int32_t width_in, float* stats_out, float eps, int64_t M,
                      int32_t N, int32_t K, int32_t epilogue, int32_t dtype, void* stream) {
    bool fp16;
    if (int rc = dtype16(dtype, &fp16)) return rc;
    LnFoldArgs f;
    f.colsum = colsum;
    f.gamma = gamma;
    f.beta = beta;
    f.stats_in = reinterpret_cast<const float2*>(stats_in);
    f.stats_out = reinterpret_cast<float2*>(stats_out);
    f.parts_in = parts_in;
    f.inv_width_in = width_in > 0 ? 1.0f / static_cast<float>(width_in) : 0.f;
    f.eps = eps;
    return launch_gemm16_fold(static_cast<const h16*>(A), lda, static_cast<const h16*>(B), ldb, static_cast<h16*>(C), ldc, bias,
                              static_cast<const h16*>(R), ldr, M, N, K, epilogue, f, fp16, static_cast<cudaStream_t>(stream));
}

size_t arb_topk_exchange_bytes(int32_t G, size_t slot_bytes) { return G > 0 ? topk_exchange_bytes(G, slot_bytes) : 0; }

int arb_exchange_alloc(size_t bytes, void** dev_ptr_out) {
    ARB_REQUIRE(dev_ptr_out != nullptr && bytes > 0, "exchange_alloc: bad arguments");
    void* p = nullptr;
    ARB_CHECK_CUDA(cudaMalloc(&p, bytes));  // a whole allocation of its own: its IPC handle maps exactly this buffer
    ARB_CHECK_CUDA(cudaMemset(p, 0, bytes));
    ARB_CHECK_CUDA(cudaDeviceSynchronize());
    *dev_ptr_out = p;
    return ARB_OK;
}

int arb_exchange_free(void* dev_ptr) {
    if (dev_ptr) ARB_CHECK_CUDA(cudaFree(dev_ptr));
    return ARB_OK;
}

int arb_ipc_export(const void* dev_ptr, void* handle_out_64) {
    ARB_REQUIRE(dev_ptr && handle_out_64, "ipc_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    ARB_CHECK_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
    memcpy(handle_out_64, &h, sizeof(h));
    return ARB_OK;
}

int arb_ipc_import(const void* handle_64, void** dev_ptr_out) {
    ARB_REQUIRE(handle_64 && dev_ptr_out, "ipc_import: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_64, sizeof(h));
    void* p = nullptr;
    ARB_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr_out = p;
    return ARB_OK;
}

int arb_ipc_close(void* dev_ptr) {
    if (dev_ptr) ARB_CHECK_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return ARB_OK;
}

int arb_topk_exchange_merge(const void* local_record_dev, const void* peer_bufs_dev, int32_t rank, int32_t G, int64_t Q,
                            int32_t k, size_t slot_bytes, float* out_scores_dev, int64_t* out_ids_dev, void* stream) {
    return launch_topk_exchange_merge(local_record_dev, static_cast<void* const*>(peer_bufs_dev), rank, G, Q, k,
                                      slot_bytes, out_scores_dev, out_ids_dev, static_cast<cudaStream_t>(stream));
}

int arb_topk_exchange_status(const void* own_buf_dev) { return topk_exchange_status(own_buf_dev); }

int arb_set_search_mode(int32_t mode) {
    ARB_REQUIRE(mode >= 0 && mode <= 2, "search mode %d must be 0 (auto), 1 (single CTA) or 2 (CTA pairs)", mode);
    set_search_mode(mode);
    return ARB_OK;
}

int arb_set_search_pace(int32_t on) {
    search_pace_ref() = on != 0;
    return ARB_OK;
}

int arb_set_pdl_mode(int32_t mode) {
    ARB_REQUIRE(mode >= 0 && mode <= 2, "pdl mode %d must be 0 (never), 1 (latency-bound calls) or 2 (always)", mode);
    pdl_mode_ref() = mode;
    return ARB_OK;
}

int arb_set_gemm_mode(int32_t mode) {
    ARB_REQUIRE(mode >= 0 && mode <= 3, "gemm mode %d must be 0 (auto), 1 (single CTA), 2 (CTA pairs) or 3 (single CTA, narrow tiles)", mode);
    set_gemm_mode(mode);
    return ARB_OK;
}

size_t arb_topk_record_bytes(int64_t Q, int32_t k) { return Q > 0 && k > 0 ? topk_record_bytes(Q, k) : 0; }
size_t arb_topk_record_ids_offset(int64_t Q, int32_t k) { return Q > 0 && k > 0 ? topk_record_ids_offset(Q, k) : 0; }

int arb_topk_merge_records(const void* records_dev, int32_t G, int64_t Q, int32_t k, float* out_scores_dev,
                           int64_t* out_ids_dev, void* stream) {
    return launch_topk_merge_records(records_dev, G, Q, k, out_scores_dev, out_ids_dev,
                                     static_cast<cudaStream_t>(stream));
}


int arb_gemm16(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
               const float* bias, const void* R, int64_t ldr, int64_t M, int32_t N, int32_t K,
               int32_t epilogue, int32_t dtype, void* stream) {
    bool f;
    if (int rc = dtype16(dtype, &f)) return rc;
    return launch_gemm16(static_cast<const h16*>(A), lda, static_cast<const h16*>(B), ldb, static_cast<h16*>(C),
                         ldc, bias, static_cast<const h16*>(R), ldr, M, N, K, epilogue, f,
                         static_cast<cudaStream_t>(stream));
}

int arb_gemm16_f32out(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                      int64_t M, int32_t N, int32_t K, int32_t dtype, void* stream) {
    bool f;
    if (int rc = dtype16(dtype, &f)) return rc;
    return launch_gemm16_f32out(static_cast<const h16*>(A), lda, static_cast<const h16*>(B), ldb, C, ldc, M, N,
                                K, f, static_cast<cudaStream_t>(stream));
}

int arb_embed_layernorm(const int32_t* ids, const float* word_emb, const float* pos_emb,
                        const float* gamma, const float* beta, void* out16, int32_t B, int32_t S,
                        int32_t H, int32_t vocab, int32_t max_pos, int32_t pad_id, int32_t position_mode,
                        float eps, int32_t dtype, void* stream) {
    bool f;
    if (int rc = dtype16(dtype, &f)) return rc;
    ARB_REQUIRE(position_mode == 0 || position_mode == 1, "embed_layernorm: position_mode %d must be 0 or 1", position_mode);
    return launch_embed_ln(ids, word_emb, pos_emb, gamma, beta, static_cast<h16*>(out16), B, S, H, vocab,
                           max_pos, pad_id, position_mode, eps, f, nullptr, static_cast<cudaStream_t>(stream));
}

int arb_adjacent_cosine(const float* emb_dev, int64_t n, int32_t D, float* out_dev, void* stream) {
    return launch_adjacent_cosine(emb_dev, out_dev, n, D, static_cast<cudaStream_t>(stream));
}

int arb_layernorm16(const void* x, const float* gamma, const float* beta, void* out, int64_t rows,
                    int32_t H, float eps, int32_t dtype, void* stream) {
    bool f;
    if (int rc = dtype16(dtype, &f)) return rc;
    return launch_layernorm(static_cast<const h16*>(x), gamma, beta, static_cast<h16*>(out), rows, H, eps, f,
                            static_cast<cudaStream_t>(stream));
}

int arb_attention16(const void* qkv, const float* rel_bias, int32_t max_rel, const int32_t* mask,
                    void* ctx, int32_t B, int32_t S, int32_t heads, int32_t head_dim, int32_t dtype,
                    int32_t impl, void* stream) {
    bool f;
    if (int rc = dtype16(dtype, &f)) return rc;
    return launch_attention(static_cast<const h16*>(qkv), rel_bias, max_rel, mask, static_cast<h16*>(ctx), B, S,
                            heads, head_dim, f, impl, static_cast<cudaStream_t>(stream));
}

int arb_pool_normalize(const void* hidden16, const int32_t* mask, float* out, int32_t B, int32_t S,
                       int32_t H, int32_t dtype, void* stream) {
    bool f;
    if (int rc = dtype16(dtype, &f)) return rc;
    return launch_pool_normalize(static_cast<const h16*>(hidden16), mask, out, B, S, H, f,
                                 static_cast<cudaStream_t>(stream));
}

}  // extern "C"
