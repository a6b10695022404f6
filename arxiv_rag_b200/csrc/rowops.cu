// HBM-bandwidth row kernels of the encode path: embedding gather + LayerNorm, LayerNorm,
// masked mean-pool + L2 normalise. All statistics are fp32; activations are bf16 in HBM.
//   embed_ln        <- MPNetEmbeddings.forward (modeling_mpnet.py:72-96) + position ids (:889-897)
//   layernorm       <- the post-LN of MPNetAttention (:210) and MPNetOutput (:242)
//   pool_normalize  <- sentence-transformers Pooling(mean) + Normalize, reached from
//                      generate_embeddings_parallel.py:146-153 (normalize_embeddings=True)
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace arb {

constexpr int kMaxVec = 8;  // H <= 1024, H % 128 == 0: each lane owns H/128 groups of 4 columns

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffff, v, o);
    return v;
}

// Normalise one row held as nvec float4 per lane and store it in the 16-bit activation format.
template <bool kF16>
__device__ __forceinline__ void ln_row_store(float4 (&x)[kMaxVec], int nvec, int H, float eps,
                                             const float* __restrict__ gamma,
                                             const float* __restrict__ beta,
                                             h16* __restrict__ out_row, int lane) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
        if (i < nvec) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    const float mean = warp_sum(s) / static_cast<float>(H);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
        if (i < nvec) {
            const float a = x[i].x - mean, b = x[i].y - mean, c = x[i].z - mean, d = x[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(H) + eps);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
        if (i < nvec) {
            const int col = (i * 32 + lane) * 4;
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
            uint2 o;
            o.x = pack16x2<kF16>((x[i].x - mean) * rstd * g.x + b.x, (x[i].y - mean) * rstd * g.y + b.y);
            o.y = pack16x2<kF16>((x[i].z - mean) * rstd * g.z + b.z, (x[i].w - mean) * rstd * g.w + b.w);
            *reinterpret_cast<uint2*>(out_row + col) = o;
        }
}

// One CTA per sequence: position ids by a prefix count of non-pad ids, then one warp per token.
template <bool kF16>
__global__ void __launch_bounds__(256)
embed_ln_kernel(const int32_t* __restrict__ ids, const float* __restrict__ word_emb,
                const float* __restrict__ pos_emb, const float* __restrict__ gamma,
                const float* __restrict__ beta, h16* __restrict__ out, int S, int H,
                int vocab, int max_pos, int pad_id, int pos_mode, float eps, int* __restrict__ err_flag) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ int smem_i[];
    int* s_pos = smem_i;              // [S] position id of each token
    int* s_chunk = smem_i + S;        // [ceil(S/32)] non-pad count per 32-token chunk
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int32_t* row_ids = ids + static_cast<int64_t>(b) * S;
    const int nchunks = (S + 31) / 32;

    for (int c = warp; c < nchunks; c += nwarps) {
        const int s = c * 32 + lane;
        const bool nonpad = s < S && row_ids[s] != pad_id;
        const unsigned bal = __ballot_sync(0xffffffff, nonpad);
        if (s < S) s_pos[s] = nonpad ? __popc(bal & (0xffffffffu >> (31 - lane))) : -1;
        if (lane == 0) s_chunk[c] = __popc(bal);
    }
    __syncthreads();
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        int p = s_pos[s];
        if (p < 0) {
            p = pad_id;  // padded token -> position row `padding_idx`
        } else {
            const int c = s >> 5;
            for (int j = 0; j < c; ++j) p += s_chunk[j];
            p += pad_id;
        }
        if (pos_mode == 1) p = s;  // BERT: absolute position = token index (BertEmbeddings.position_ids)
        s_pos[s] = min(p, max_pos - 1);
    }
    __syncthreads();

    // gridDim.y CTAs share a sequence (each recomputed the positions above): a token costs a warp ~3 us
    // of dependent gather latency, so a lone 64-token query wants 8 CTAs, not 8 tokens per warp in turn
    const int nvec = H / 128;
    for (int s = static_cast<int>(blockIdx.y) * nwarps + warp; s < S; s += nwarps * static_cast<int>(gridDim.y)) {
        int id = row_ids[s];
        if (id < 0 || id >= vocab) {
            // torch's embedding raises on such an id. Report it through the handle's status word
            // (host-mapped, read by arb_mpnet_status after the stream is synchronised) and clamp so
            // the gather itself stays in bounds.
            if (lane == 0 && err_flag != nullptr) {
                err_flag[1] = id;
                err_flag[2] = b * S + s;
                __threadfence_system();
                err_flag[0] = 1;
            }
            id = min(max(id, 0), vocab - 1);
        }
        const float* w = word_emb + static_cast<int64_t>(id) * H;
        const float* p = pos_emb + static_cast<int64_t>(s_pos[s]) * H;
        float4 x[kMaxVec];
#pragma unroll
        for (int i = 0; i < kMaxVec; ++i)
            if (i < nvec) {
                const int col = (i * 32 + lane) * 4;
                const float4 a = __ldg(reinterpret_cast<const float4*>(w + col));
                const float4 c = __ldg(reinterpret_cast<const float4*>(p + col));
                x[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
            }
        ln_row_store<kF16>(x, nvec, H, eps, gamma, beta, out + (static_cast<int64_t>(b) * S + s) * H, lane);
    }
}

// One warp per row; rows are independent so the grid is simply sized to cover them.
template <bool kF16>
__global__ void __launch_bounds__(256)
layernorm_kernel(const h16* __restrict__ x, const float* __restrict__ gamma,
                 const float* __restrict__ beta, h16* __restrict__ out, int64_t rows, int H,
                 float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int nvec = H / 128;
    const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
         row < rows; row += warps_total) {
        const h16* xr = x + row * H;
        float4 v[kMaxVec];
#pragma unroll
        for (int i = 0; i < kMaxVec; ++i)
            if (i < nvec) {
                const uint2 u = __ldg(reinterpret_cast<const uint2*>(xr + (i * 32 + lane) * 4));
                const float2 a = unpack16x2<kF16>(u.x), b = unpack16x2<kF16>(u.y);
                v[i] = make_float4(a.x, a.y, b.x, b.y);
            }
        ln_row_store<kF16>(v, nvec, H, eps, gamma, beta, out + row * H, lane);
    }
}

// One CTA per sequence, (H/4) x G threads: thread (g, c) sums columns 4c..4c+3 over the
// unmasked tokens s = g (mod G). Masked tokens are never read.
constexpr int kPoolGroups = 4;
template <bool kF16>
__global__ void __launch_bounds__(1024)
pool_normalize_kernel(const h16* __restrict__ hidden, const int32_t* __restrict__ mask,
                      float* __restrict__ out, int S, int H) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float smem_f[];
    float* s_part = smem_f;                     // [G][H]
    float* s_red = smem_f + kPoolGroups * H;    // [32] block-reduction scratch
    float* s_cnt = s_red + 32;                  // [G] per-group token counts
    const int b = blockIdx.x;
    const int tpg = H / 4;  // threads per group
    const int g = threadIdx.x / tpg, c = threadIdx.x % tpg;
    const h16* hb = hidden + static_cast<int64_t>(b) * S * H;
    const int32_t* mb = mask + static_cast<int64_t>(b) * S;

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float cnt_g = 0.f;
    for (int s = g; s < S; s += kPoolGroups) {
        const float m = static_cast<float>(mb[s]);
        cnt_g += m;
        if (m != 0.f) {
            const uint2 u = __ldg(reinterpret_cast<const uint2*>(hb + static_cast<int64_t>(s) * H + c * 4));
            const float2 a = unpack16x2<kF16>(u.x), d = unpack16x2<kF16>(u.y);
            acc.x += m * a.x;
            acc.y += m * a.y;
            acc.z += m * d.x;
            acc.w += m * d.y;
        }
    }
    *reinterpret_cast<float4*>(s_part + g * H + c * 4) = acc;
    if (c == 0) s_cnt[g] = cnt_g;
    __syncthreads();
    float cnt = 0.f;  // sum of the mask over the whole sequence
#pragma unroll
    for (int j = 0; j < kPoolGroups; ++j) cnt += s_cnt[j];

    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    float sq = 0.f;
    if (g == 0) {
#pragma unroll
        for (int j = 0; j < kPoolGroups; ++j) {
            const float4 p = *reinterpret_cast<const float4*>(s_part + j * H + c * 4);
            v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
        }
        const float inv = 1.0f / fmaxf(cnt, 1e-9f);  // Pooling: sum / clamp(sum_mask, min=1e-9)
        v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
        sq = (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sq;
    __syncthreads();
    float tot = 0.f;
    const int nw = blockDim.x >> 5;
    for (int j = 0; j < nw; ++j) tot += s_red[j];
    if (g == 0) {
        const float inv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);  // F.normalize(p=2, eps=1e-12)
        *reinterpret_cast<float4*>(out + static_cast<int64_t>(b) * H + c * 4) =
            make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
    }
}

// The last LayerNorm folded into the pooling: reads the PRE-LayerNorm rows and the per-row partials
// their producer wrote (gemm.cu EPI_BIAS_LNRES_STATS), so no LayerNorm pass is left anywhere in the
// folded forward. mean_t(LN(x_t)) = gamma * mean_t((x_t - mu_t) rstd_t) + beta, over unmasked tokens.
template <bool kF16>
__global__ void __launch_bounds__(1024)
pool_ln_normalize_kernel(const h16* __restrict__ pre, const float2* __restrict__ stats, int parts, int64_t total_rows,
                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                         const int32_t* __restrict__ mask, float* __restrict__ out, int S, int H) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float smem_f[];
    float* s_part = smem_f;                     // [G][H]
    float* s_red = smem_f + kPoolGroups * H;    // [32] block-reduction scratch
    float* s_cnt = s_red + 32;                  // [G] per-group token counts
    float* s_rs = s_cnt + kPoolGroups;          // [S] rstd of each token's row
    float* s_nm = s_rs + S;                     // [S] -mean * rstd
    const int b = blockIdx.x;
    const int tpg = H / 4;  // threads per group
    const int g = threadIdx.x / tpg, c = threadIdx.x % tpg;
    const h16* hb = pre + static_cast<int64_t>(b) * S * H;
    const int32_t* mb = mask + static_cast<int64_t>(b) * S;
    const float inv_h = 1.0f / static_cast<float>(H);
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        if (mb[s] == 0) continue;
        const int64_t row = static_cast<int64_t>(b) * S + s;
        float su = 0.f, sq = 0.f;
        for (int p = 0; p < parts; ++p) {
            const float2 t = __ldg(stats + static_cast<int64_t>(p) * total_rows + row);
            su += t.x;
            sq += t.y;
        }
        const float mean = su * inv_h;
        const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, sq * inv_h), 0.f) + eps);
        s_rs[s] = rstd;
        s_nm[s] = -mean * rstd;
    }
    __syncthreads();

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float cnt_g = 0.f;
    for (int s = g; s < S; s += kPoolGroups) {
        const float m = static_cast<float>(mb[s]);
        cnt_g += m;
        if (m != 0.f) {
            const uint2 u = __ldg(reinterpret_cast<const uint2*>(hb + static_cast<int64_t>(s) * H + c * 4));
            const float2 a = unpack16x2<kF16>(u.x), d = unpack16x2<kF16>(u.y);
            const float rs = s_rs[s], nm = s_nm[s];
            acc.x += m * fmaf(a.x, rs, nm);
            acc.y += m * fmaf(a.y, rs, nm);
            acc.z += m * fmaf(d.x, rs, nm);
            acc.w += m * fmaf(d.y, rs, nm);
        }
    }
    *reinterpret_cast<float4*>(s_part + g * H + c * 4) = acc;
    if (c == 0) s_cnt[g] = cnt_g;
    __syncthreads();
    float cnt = 0.f;
#pragma unroll
    for (int j = 0; j < kPoolGroups; ++j) cnt += s_cnt[j];

    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    float sq = 0.f;
    if (g == 0) {
#pragma unroll
        for (int j = 0; j < kPoolGroups; ++j) {
            const float4 p = *reinterpret_cast<const float4*>(s_part + j * H + c * 4);
            v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
        }
        const float inv = 1.0f / fmaxf(cnt, 1e-9f);  // Pooling: sum / clamp(sum_mask, min=1e-9)
        const float wb = cnt * inv;                  // 1 for any real row, 0 for an all-pad row (-> zero vector)
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c * 4));
        const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + c * 4));
        v.x = fmaf(gm.x, v.x * inv, bt.x * wb);
        v.y = fmaf(gm.y, v.y * inv, bt.y * wb);
        v.z = fmaf(gm.z, v.z * inv, bt.z * wb);
        v.w = fmaf(gm.w, v.w * inv, bt.w * wb);
        sq = (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sq;
    __syncthreads();
    float tot = 0.f;
    const int nw = blockDim.x >> 5;
    for (int j = 0; j < nw; ++j) tot += s_red[j];
    if (g == 0) {
        const float inv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);  // F.normalize(p=2, eps=1e-12)
        *reinterpret_cast<float4*>(out + static_cast<int64_t>(b) * H + c * 4) =
            make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
    }
}

// cos(e[i], e[i-1]) for consecutive rows — TextChunker._cosine_similarity
// (text_processor.py:1601-1605) over the adjacent sentence pairs of _chunk_semantic (:1547-1561).
// One warp per row; out[0] = 1.
__global__ void __launch_bounds__(256)
adjacent_cosine_kernel(const float* __restrict__ emb, float* __restrict__ out, int64_t n, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    if (i == 0) {
        if (lane == 0) out[0] = 1.0f;
        return;
    }
    const float* a = emb + i * D;
    const float* b = a - D;
    float dot = 0.f, na = 0.f, nb = 0.f;
    for (int d = lane * 4; d < D; d += 128) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(a + d));
        const float4 y = __ldg(reinterpret_cast<const float4*>(b + d));
        dot += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
        na += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        nb += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
    }
    dot = warp_sum(dot);
    na = warp_sum(na);
    nb = warp_sum(nb);
    if (lane == 0) out[i] = dot / (sqrtf(na) * sqrtf(nb));
}

int launch_adjacent_cosine(const float* emb, float* out, int64_t n, int D, cudaStream_t stream) {
    ARB_REQUIRE(emb && out, "adjacent_cosine: null pointer");
    ARB_REQUIRE(n > 0 && D > 0 && D % 4 == 0, "adjacent_cosine: bad shape n=%lld D=%d", (long long)n, D);
    const int64_t blocks = (n + 7) / 8;
    ARB_REQUIRE(blocks < (1ll << 31), "adjacent_cosine: n too large");
    adjacent_cosine_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(emb, out, n, D);
    ARB_CHECK_CUDA(cudaGetLastError());
    return ARB_OK;
}

static int check_h(int H) {
    ARB_REQUIRE(H > 0 && H % 128 == 0 && H <= 128 * kMaxVec, "hidden size %d unsupported (need H %% 128 == 0, H <= %d)", H, 128 * kMaxVec);
    return ARB_OK;
}

int launch_embed_ln(const int32_t* ids, const float* word_emb, const float* pos_emb,
                    const float* gamma, const float* beta, h16* out, int B, int S, int H, int vocab,
                    int max_pos, int pad_id, int pos_mode, float eps, bool fp16, int* err_flag_dev,
                    cudaStream_t stream) {
    ARB_REQUIRE(ids && word_emb && pos_emb && gamma && beta && out, "embed_ln: null pointer");
    ARB_REQUIRE(B > 0 && S > 0 && S <= 4096, "embed_ln: bad shape B=%d S=%d", B, S);
    if (int rc = check_h(H)) return rc;
    const size_t smem = (S + (S + 31) / 32) * sizeof(int);
    auto kern = fp16 ? embed_ln_kernel<true> : embed_ln_kernel<false>;
    int gy = (4 * num_sms()) / B;  // enough CTAs to cover the SMs a few times over, never more than one warp pass each
    if (gy > (S + 7) / 8) gy = (S + 7) / 8;
    if (gy < 1) gy = 1;
    ARB_CHECK_CUDA(launch_kernel(kern, dim3(B, gy), dim3(256), smem, stream, 1, ids, word_emb, pos_emb, gamma, beta, out, S, H,
                                 vocab, max_pos, pad_id, pos_mode, eps, err_flag_dev));
    return ARB_OK;
}

int launch_layernorm(const h16* x, const float* gamma, const float* beta, h16* out, int64_t rows,
                     int H, float eps, bool fp16, cudaStream_t stream) {
    ARB_REQUIRE(x && gamma && beta && out, "layernorm: null pointer");
    ARB_REQUIRE(rows > 0, "layernorm: rows=%lld", (long long)rows);
    if (int rc = check_h(H)) return rc;
    const int64_t blocks_needed = (rows + 7) / 8;
    const int64_t cap = static_cast<int64_t>(num_sms()) * 32;
    const int grid = static_cast<int>(blocks_needed < cap ? blocks_needed : cap);
    auto kern = fp16 ? layernorm_kernel<true> : layernorm_kernel<false>;
    ARB_CHECK_CUDA(launch_kernel(kern, dim3(grid), dim3(256), 0, stream, 1, x, gamma, beta, out, rows, H, eps));
    return ARB_OK;
}

int launch_pool_normalize(const h16* hidden, const int32_t* mask, float* out, int B, int S, int H,
                          bool fp16, cudaStream_t stream) {
    ARB_REQUIRE(hidden && mask && out, "pool_normalize: null pointer");
    ARB_REQUIRE(B > 0 && S > 0, "pool_normalize: bad shape B=%d S=%d", B, S);
    if (int rc = check_h(H)) return rc;
    const int threads = (H / 4) * kPoolGroups;
    ARB_REQUIRE(threads <= 1024 && threads % 32 == 0, "pool_normalize: H=%d unsupported", H);
    const size_t smem = (kPoolGroups * H + 32 + kPoolGroups) * sizeof(float);
    auto kern = fp16 ? pool_normalize_kernel<true> : pool_normalize_kernel<false>;
    ARB_CHECK_CUDA(launch_kernel(kern, dim3(B), dim3(threads), smem, stream, 1, hidden, mask, out, S, H));
    return ARB_OK;
}

int launch_pool_ln_normalize(const h16* pre, const float2* stats, int parts, const float* gamma, const float* beta,
                             float eps, const int32_t* mask, float* out, int B, int S, int H, bool fp16,
                             cudaStream_t stream) {
    ARB_REQUIRE(pre && stats && gamma && beta && mask && out, "pool_ln_normalize: null pointer");
    ARB_REQUIRE(B > 0 && S > 0 && parts > 0, "pool_ln_normalize: bad shape B=%d S=%d parts=%d", B, S, parts);
    if (int rc = check_h(H)) return rc;
    const int threads = (H / 4) * kPoolGroups;
    ARB_REQUIRE(threads <= 1024 && threads % 32 == 0, "pool_ln_normalize: H=%d unsupported", H);
    const size_t smem = (kPoolGroups * H + 32 + kPoolGroups + 2 * static_cast<size_t>(S)) * sizeof(float);
    auto kern = fp16 ? pool_ln_normalize_kernel<true> : pool_ln_normalize_kernel<false>;
    ARB_CHECK_CUDA(launch_kernel(kern, dim3(B), dim3(threads), smem, stream, 1, pre, stats, parts,
                                 static_cast<int64_t>(B) * S, gamma, beta, eps, mask, out, S, H));
    return ARB_OK;
}

}  // namespace arb
