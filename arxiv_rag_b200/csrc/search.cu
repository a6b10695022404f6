// Exact cosine top-k search: scores = Q . C^T on tcgen05 tensor cores with the per-query
// running top-k fused into the TMEM epilogue, so the [Q, N] score matrix never reaches HBM.
// The reference has no search routine; the semantic definition generalised here is
// TextChunker._cosine_similarity (3-chunks/pipeline/src/processors/text_processor.py:1601-1605)
// on unit-norm rows (generate_embeddings_parallel.py:149 normalize_embeddings=True), with
// top_k from 3-chunks/pipeline/config.yaml:62-64. Ties are ordered by ascending row id.
//
// Work decomposition: item = (corpus split, 128-query tile), split-major, dealt round-robin to a
// persistent grid so the CTAs running together sweep the same corpus region (L2 reuse). Each
// epilogue thread owns one query row (one TMEM lane): it scans the 256 scores of every chunk
// against its k-th best, appends the rare survivors to a small per-row candidate buffer in shared
// memory, and the buffers are folded into the sorted top-k lists in batches (k <= 16: lists in
// registers, each thread folds its own buffer; k <= 128: lists in shared memory, merged by the
// whole warp one row at a time). Per-item lists go to the workspace and a k-way merge kernel
// produces the final order.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tmap.cuh"
#include "umma_pipe.cuh"

namespace arb {

constexpr int kSBN = 256;     // corpus rows per chunk (MMA N)
constexpr int kMaxK = 128;    // k <= 16: lists in registers; else 128 rows x k x 8 B of shared memory
// smem ring depth: 4 x 48 KB stages for k <= 16, 3 while list + a >= 16-candidate buffer fit, else 2

struct SearchPlan {
    bool pair;     // CTA-pair schedule (two query tiles per work item)
    int nq;        // 128-query tiles, or pairs of them
    int nchunks;   // corpus chunks of kSBN rows
    int nsplit;    // corpus splits
    int cps;       // chunks per split (last split may be short)
    int grid;      // CTAs
};

constexpr int kSharedSplitMaxChunks = 1000;  // ~390 MB of bf16 rows: see make_plan

// 0 = auto (pairs as soon as there is more than one query tile), 1 = single CTA, 2 = pairs
static int g_search_mode = 0;
void set_search_mode(int mode) { g_search_mode = mode; }
// pacing of the units that share a corpus split (SearchTileIter): on unless ARB_SEARCH_PACE=0; arb_set_search_pace
bool& search_pace_ref() {
    static bool on = []() {
        const char* e = getenv("ARB_SEARCH_PACE");
        return !(e && e[0] == '0');
    }();
    return on;
}

static SearchPlan make_plan(int64_t Q, int64_t N, int k) {
    SearchPlan p;
    const int tiles = static_cast<int>((Q + kBM - 1) / kBM);
    p.pair = g_search_mode == 2 || (g_search_mode == 0 && tiles > 1);
    p.nq = p.pair ? (tiles + 1) / 2 : tiles;
    p.nchunks = static_cast<int>((N + kSBN - 1) / kSBN);
    const int G = p.pair ? num_sms() / 2 : num_sms();  // schedulable units: CTAs or CTA pairs
    // Every item (split x query tile) starts with empty lists and pays a warm-up while its k-th
    // best is still low: ~k (1 + ln(rows/k)) survivors per query row have to be merged, which for
    // the shared-memory lists (k > 16) costs about 1.4 k chunk-times of epilogue work per item,
    // for the register lists a few chunks. time ~ rounds * (cps + warm); ideal ~ nq * nchunks / G.
    // Pick the split count with the best ratio; prefer fewer splits on near-ties (less merge work).
    const double warm = k > 16 ? 1.4 * k : 3.0;
    int best = 1;
    double best_eff = -1.0;
    const int max_split = p.nchunks < 4 ? 1 : (p.nchunks / 4 < 4 * G ? p.nchunks / 4 : 4 * G);
    // Several query tiles stream the same split side by side and share its chunks through L2 only
    // while they stay within the cache of each other. Over a long split they drift apart (measured:
    // 2 171-chunk splits at Q = 4096 over 5 M rows re-read the corpus 2.7 times from DRAM), so splits
    // walked by 12 or more tile pairs are kept short: 2.5-6 % faster at Q = 4096 / 8192 over 5 M rows,
    // DRAM reads 2.7x -> 1.8x the corpus. With few tiles per split (Q = 700: +5 %, 2048: +2.5 %) and from
    // half of the pairs per split up (Q = 16 384: +2 %, 32 768: +8 %) it is the other way round, so no
    // limit there. ARB_SEARCH_MAX_CPS overrides the length (0 = no limit).
    static const int env_max_cps = []() {
        const char* e = getenv("ARB_SEARCH_MAX_CPS");
        return e ? atoi(e) : -1;
    }();
    const int max_cps = env_max_cps >= 0 ? env_max_cps : kSharedSplitMaxChunks;
    // (only where a fresh item warms up in a few chunks: the shared-memory lists of k > 16 need long splits)
    const int min_split = (p.nq >= 12 && 2 * p.nq <= G && max_cps > 0 && k <= 16) ? (p.nchunks + max_cps - 1) / max_cps : 1;
    for (int first = min_split < max_split ? min_split : max_split; best_eff < 0.0; first = 1) {
        for (int s = first; s <= max_split; ++s) {
            const int cps = (p.nchunks + s - 1) / s;
            const int s_eff = (p.nchunks + cps - 1) / cps;  // splits actually non-empty
            if (s_eff != s) continue;
            const int64_t items = static_cast<int64_t>(p.nq) * s;
            const int64_t rounds = (items + G - 1) / G;
            const double eff = (static_cast<double>(p.nq) * p.nchunks / G) / (static_cast<double>(rounds) * (cps + warm));
            if (eff > best_eff + 0.02) {
                best_eff = eff;
                best = s;
            }
            if (items >= 16ll * G && s >= first + 8) break;
        }
        if (first == 1) break;
    }
    p.nsplit = best;
    p.cps = (p.nchunks + best - 1) / best;
    const int64_t items = static_cast<int64_t>(p.nq) * p.nsplit;
    p.grid = static_cast<int>(items < G ? items : G) * (p.pair ? 2 : 1);
    return p;
}

__device__ __forceinline__ void st_relaxed_gpu(int* p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Yields the (query-tile row, corpus chunk row) sequence of this CTA's items.
// nq = query tiles (single CTA) or query-tile PAIRS (CTA-pair schedule, where this CTA takes tile
// 2 * pair + rank: qmul = 2, qadd = rank).
//
// Pacing (`progress` != nullptr: the TMA-issuing thread of a work unit only). The query tiles that
// walk one split at the same time share its chunks through L2 only while they stay close together;
// left alone they drift apart and each chunk comes from DRAM two or three times. So a unit publishes
// the chunk it is about to load every kPaceEvery chunks and looks at ONE peer per check, round-robin
// over the units that walk the same split in the same round (units are dealt items round-robin, so
// round = item / step, and every item of a round is started by some unit: a wait always ends; it is
// bounded all the same): it does not run more than kPaceWindow chunks ahead of a peer it has seen. The
// peer's slot is requested one check before it is judged, so the TMA-issuing thread never stalls on
// the load (a blocking scan of all peers cost 10-40 % of the kernel). progress[] lives in the caller's workspace and needs no
// initialisation: an item's slot is reset when the item starts, a finished item's slot holds its
// full length, and whatever else a peer's slot holds either ends the wait at once or makes the
// leader wait for a peer that is about to start.
constexpr int kPaceEvery = 4;
constexpr int kPaceWindow = 16;
constexpr int kPaceMaxSpins = 4000;  // x ~50 ns: give up after ~0.2 ms, never hang

struct SearchTileIter {
    int item, step, items, nq, cps, nchunks, qmul, qadd;
    int chunk = 0, chunk_end = 0, row_q = 0;
    int* progress = nullptr;
    int cur = -1, chunk_begin = 0, peer_lo = 0, peer_hi = 0, rot = 0, seen_peer = 0, seen = 1 << 30;
    __device__ __forceinline__ SearchTileIter(int first, int step_, int items_, int nq_, int cps_,
                                              int nchunks_, int qmul_ = 1, int qadd_ = 0)
        : item(first), step(step_), items(items_), nq(nq_), cps(cps_), nchunks(nchunks_), qmul(qmul_), qadd(qadd_) {}
    __device__ __forceinline__ bool next(int& row_a, int& row_b) {
        while (chunk >= chunk_end) {
            if (progress != nullptr && cur >= 0) st_relaxed_gpu(progress + cur, cps + kPaceWindow);  // done: never holds anyone back
            if (item >= items) {
                cur = -1;
                return false;
            }
            const int split = item / nq;
            row_q = ((item % nq) * qmul + qadd) * kBM;
            chunk = split * cps;
            chunk_end = min(chunk + cps, nchunks);
            if (progress != nullptr) {
                cur = item;
                chunk_begin = chunk;
                const int round = item / step;
                peer_lo = max(split * nq, round * step);
                peer_hi = min(min((split + 1) * nq, (round + 1) * step), items);
                seen = 1 << 30;
                rot = cur - peer_lo;
                st_relaxed_gpu(progress + cur, 0);
            }
            item += step;
        }
        if (progress != nullptr) {
            const int c = chunk - chunk_begin;
            if ((c & (kPaceEvery - 1)) == 0) {
                st_relaxed_gpu(progress + cur, c);
                // judge the peer whose slot was read at the previous check (the load has long landed: no
                // stall on the TMA-issuing thread), re-polling only if it really lags
                for (int spin = 0; seen + kPaceWindow + kPaceEvery < c && spin < kPaceMaxSpins; ++spin) {
                    __nanosleep(50);
                    seen = ld_relaxed_gpu(progress + seen_peer);
                }
                // and request the next peer's slot, round-robin over the units on this split
                const int n = peer_hi - peer_lo;
                if (n > 1) {
                    rot = rot + 1 < n ? rot + 1 : 0;
                    if (peer_lo + rot == cur) rot = rot + 1 < n ? rot + 1 : 0;
                    seen_peer = peer_lo + rot;
                    seen = ld_relaxed_gpu(progress + seen_peer);
                }
            }
        }
        row_a = row_q;
        row_b = chunk * kSBN;
        ++chunk;
        return true;
    }
};

constexpr int kSearchThreads = 192;  // TMA warp, MMA warp, 4 epilogue warps (one per TMEM lane quarter)

// k > 16: each row's sorted list lives in shared memory ([row][k], row-major) next to a small
// unsorted candidate buffer ([row][cbuf]) that the row's own thread appends to. When a buffer is
// full the whole warp merges it into the list in one pass: every lane owns list slots lane,
// lane+32, ... (k <= 128) and one candidate; a list entry moves down by the number of candidates
// that beat it, a candidate lands at (#list entries >= it) + (its rank among the candidates), so
// equal scores keep arrival (= ascending id) order. All reads precede all writes. Returns the
// row's new k-th best. Amortised cost per candidate is ~1/cbuf of a single-element insertion.
__device__ __noinline__ float warp_list_merge(float* ls, int* li, int k, const float* cs, const int* ci,
                                              int n, int lane) {
    __syncwarp();  // the owner's buffer stores are visible to the warp
    float s[4];
    int d[4], adv[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int i = lane + 32 * t;
        s[t] = i < k ? ls[i] : -INFINITY;
        d[t] = i < k ? li[i] : -1;
        adv[t] = 0;
    }
    const bool has = lane < n;
    const float cv = has ? cs[lane] : -INFINITY;
    const int cid = has ? ci[lane] : -1;
    int crank = 0;
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
        const float sj = cs[j];  // broadcast read
#pragma unroll
        for (int t = 0; t < 4; ++t) adv[t] += sj > s[t] ? 1 : 0;
        crank += (sj > cv || (sj == cv && j < lane)) ? 1 : 0;
    }
    int lo = 0, hi = k;  // first list index whose score is < cv (the list is descending)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ls[mid] >= cv) lo = mid + 1;
        else hi = mid;
    }
    __syncwarp();  // every slot has been read before any slot is overwritten
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int i = lane + 32 * t;
        const int p = i + adv[t];
        if (i < k && p < k) {
            ls[p] = s[t];
            li[p] = d[t];
        }
    }
    if (has && lo + crank < k) {
        ls[lo + crank] = cv;
        li[lo + crank] = cid;
    }
    __syncwarp();
    return ls[k - 1];
}

// Register-resident sorted list (k <= KR <= 16): a branch-free carry chain, 5 instructions per
// slot. A lane whose candidate does not beat its k-th best falls through unchanged, so the warp
// runs the chain only when __any lane has a survivor. Strict '>' keeps equal scores in arrival
// (= ascending id) order.
template <int KR>
__device__ __forceinline__ void reg_insert(float (&s)[KR], int (&id)[KR], float v, int nid) {
#pragma unroll
    for (int i = 0; i < KR; ++i) {
        const bool p = v > s[i];
        const float ts = s[i];
        const int ti = id[i];
        s[i] = p ? v : ts;
        id[i] = p ? nid : ti;
        v = p ? ts : v;
        nid = p ? ti : nid;
    }
}

// KR > 0: top-k lists in registers (k <= KR); KR == 0: lists in shared memory (k <= kMaxK).
// kPair: clusters of two CTAs (tcgen05 cta_group::2) score 256 queries against each corpus chunk;
// every CTA stages only half of the chunk (32 KB stages -> a deeper ring beside large lists, half
// the L2 -> SM corpus traffic) and keeps the lists of its own 128 query rows. `nq` then counts
// query-tile pairs.
// kTf32: the operands are fp32 rows (D = 16-bit words per row = 2 x the fp32 width) scored with
// kind::tf32 MMAs.
template <int kSStages, int KR, bool kPair, bool kTf32 = false>
__global__ void __launch_bounds__(kSearchThreads, 1)
search_topk_kernel(const __grid_constant__ CUtensorMap tmap_q,
                   const __grid_constant__ CUtensorMap tmap_c, float* __restrict__ part_scores,
                   int32_t* __restrict__ part_ids, int64_t Q, int64_t N, int D, int k, int nq, int cps,
                   int nchunks, int nsplit, int cbuf, int* __restrict__ progress) {
    using SM = PipeSmem<kSBN, kSStages, 0, kPair ? 2 : 1>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SM sm{smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u)};
    const int warp = __shfl_sync(0xffffffff, threadIdx.x / 32, 0);
    const int lane = threadIdx.x & 31;
    const int kblocks = (D + kBK - 1) / kBK;
    const int items = nq * nsplit;
    const int rank = kPair ? static_cast<int>(cluster_ctarank()) : 0;
    const int item0 = kPair ? static_cast<int>(blockIdx.x) / 2 : static_cast<int>(blockIdx.x);
    const int item_step = kPair ? static_cast<int>(gridDim.x) / 2 : static_cast<int>(gridDim.x);
    SearchTileIter it(item0, item_step, items, nq, cps, nchunks, kPair ? 2 : 1, rank);

    uint32_t tmem_base;
    if constexpr (kPair) tmem_base = pipe2_setup(sm, warp, &tmap_q, &tmap_c, 2 * 4);
    else tmem_base = pipe_setup(sm, warp, &tmap_q, &tmap_c, 128);

    if (warp == 0) {
        // queries are re-read by every chunk -> keep in L2; the corpus streams through once per split
        if (elect_one()) {
            SearchTileIter pit = it;
            if (rank == 0 && nq > 1) pit.progress = progress;  // pacing among the tiles that share a split
            if constexpr (kPair) pipe2_produce(sm, &tmap_q, &tmap_c, pit, kblocks, rank, kEvictLast, kEvictNormal);
            else pipe_produce(sm, &tmap_q, &tmap_c, pit, kblocks, kEvictLast, kEvictNormal);
        }
    } else if (warp == 1) {
        if (elect_one()) {
            if constexpr (kPair) {
                if (rank == 0)
                    pipe2_mma<SM, kTf32>(sm, tmem_base, it, kblocks,
                                         kTf32 ? umma_idesc_tf32(2 * kBM, SM::kBN) : umma_idesc_16bit(2 * kBM, SM::kBN, false));
            } else {
                pipe_mma<SM, kTf32>(sm, tmem_base, it, kblocks,
                                    kTf32 ? umma_idesc_tf32(kBM, SM::kBN) : umma_idesc_16bit(kBM, SM::kBN, false));
            }
        }
    } else {
        const int lane_grp = warp & 3;
        const int trow = lane_grp * 32 + lane;  // row of the 128-query tile owned by this thread
        // shared memory after the ring: [128][k] list scores, [128][k] list ids (KR == 0 only), then
        // the candidate buffers [128][cstride] scores and ids (odd stride: conflict-free appends)
        const int lk = KR > 0 ? 0 : k;
        const int cstride = cbuf > 1 ? cbuf + 1 : cbuf;
        float* ls_all = reinterpret_cast<float*>(sm.extra());
        int* li_all = reinterpret_cast<int*>(sm.extra() + static_cast<size_t>(lk) * kBM * 4);
        float* ls = ls_all + trow * lk;
        int* li = li_all + trow * lk;
        float* cbs_all = reinterpret_cast<float*>(sm.extra() + static_cast<size_t>(lk) * kBM * 8);
        int* cbi_all = reinterpret_cast<int*>(sm.extra() + static_cast<size_t>(lk) * kBM * 8 +
                                              static_cast<size_t>(cstride) * kBM * 4);
        float* cbs = cbs_all + trow * cstride;
        int* cbi = cbi_all + trow * cstride;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = item0; item < items; item += item_step) {
            const int split = item / nq;
            const int qt = kPair ? (item % nq) * 2 + rank : item % nq;
            const int c_begin = split * cps;
            const int c_end = min(c_begin + cps, nchunks);
            constexpr int KRA = KR > 0 ? KR : 1;
            float rs[KRA];
            int ri[KRA];
#pragma unroll
            for (int i = 0; i < KRA; ++i) {
                rs[i] = -INFINITY;
                ri[i] = -1;
            }
            if (KR == 0) {
                for (int i = 0; i < k; ++i) {
                    ls[i] = -INFINITY;
                    li[i] = -1;
                }
                __syncwarp();  // lists are touched by the whole warp from here on
            }
            // Survivors of the scan are appended to the row's buffer by its own thread; `thr` is the
            // row's k-th best as of the last merge (stale = lower, so nothing is ever missed).
            float thr = -INFINITY;
            int cnt = 0;
            // KR == 0: merge the buffers of the rows in `m` into their lists, one row at a time, whole warp
            auto flush = [&](unsigned m) {
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const int row = lane_grp * 32 + src;
                    const int n = __shfl_sync(0xffffffff, cnt, src);
                    const float nt = warp_list_merge(ls_all + row * k, li_all + row * k, k, cbs_all + row * cstride,
                                                     cbi_all + row * cstride, n, lane);
                    if (lane == src) {
                        thr = nt;
                        cnt = 0;
                    }
                }
            };
            // KR > 0: every thread folds its own buffer into its register list; the 32 rows of the
            // warp advance together, so the cost is the longest buffer, not the sum
            auto drain = [&]() {
                const int longest = __reduce_max_sync(0xffffffff, cnt);
                for (int i = 0; i < longest; ++i) {
                    const bool on = i < cnt;
                    reg_insert<KRA>(rs, ri, on ? cbs[i] : -INFINITY, on ? cbi[i] : -1);
                }
                cnt = 0;
                thr = rs[KRA - 1];
            };
            for (int chunk = c_begin; chunk < c_end; ++chunk) {
                mbar_wait(sm.tmem_full(acc), acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) +
                                       static_cast<uint32_t>(acc * kSBN);
                const int64_t n0 = static_cast<int64_t>(chunk) * kSBN;
                const int nvalid = static_cast<int>(N - n0 < kSBN ? N - n0 : kSBN);
                // scan 32 score columns of this thread's query row against its running k-th best
                auto scan = [&](const uint32_t (&r)[32], int c0) {
                    if (c0 >= nvalid) return;  // warp-uniform: columns past the corpus end
                    const bool full = c0 + 32 <= nvalid;
                    const int id0 = static_cast<int>(n0) + c0;
                    // cheap pre-filter: after warm-up almost no 32x32 block holds a survivor
                    float mx = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        mx = fmaxf(mx, (full || c0 + j < nvalid) ? __uint_as_float(r[j]) : -INFINITY);
                    if (!__any_sync(0xffffffff, mx > thr)) return;
                    // columns of this row that beat its (possibly stale) k-th best
                    unsigned pass = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) pass |= (__uint_as_float(r[j]) > thr ? 1u : 0u) << j;
                    if (!full) pass &= (1u << (nvalid - c0)) - 1u;
                    const int np = __popc(pass);
                    // make room first: rows whose buffer cannot take this block's survivors
                    if (KR > 0) {
                        if (__any_sync(0xffffffff, cnt + np > cbuf)) drain();  // cbuf == 32 >= np
                    } else {
                        const unsigned tight = __ballot_sync(0xffffffff, cnt > 0 && cnt + np > cbuf);
                        if (tight) flush(tight);
                    }
                    if (KR > 0 || !__any_sync(0xffffffff, np > cbuf)) {
                        // common case: append without any further warp synchronisation
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (pass & (1u << j)) {
                                cbs[cnt] = __uint_as_float(r[j]);
                                cbi[cnt] = id0 + j;
                                ++cnt;
                            }
                        }
                    } else {
                        // a row has more survivors than a whole buffer holds (list still filling up,
                        // or a small buffer beside a large k): append one column at a time
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if ((pass & (1u << j)) && __uint_as_float(r[j]) > thr) {
                                cbs[cnt] = __uint_as_float(r[j]);
                                cbi[cnt] = id0 + j;
                                ++cnt;
                            }
                            const unsigned m = __ballot_sync(0xffffffff, cnt == cbuf);
                            if (m) flush(m);
                        }
                    }
                };
                // (one instance of the scan code: duplicating it for a second register buffer
                // overflows the instruction cache and costs more than the TMEM latency it hides)
#pragma unroll 1
                for (int c0 = 0; c0 < kSBN; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + c0, r);
                    tmem_ld_wait();
                    if (c0 + 32 == kSBN) {  // the whole accumulator has been read: hand the TMEM buffer back
                        tc_fence_before();
                        if constexpr (kPair) {  // one arrival per warp on the leader CTA's barrier
                            __syncwarp();
                            if (lane == 0) mbar_arrive_remote_cta(map_to_cta(smem_u32(sm.tmem_empty(acc)), 0));
                        } else {
                            mbar_arrive(sm.tmem_empty(acc));
                        }
                    }
                    scan(r, c0);
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
            // fold in what is still buffered
            if (KR > 0) drain();
            else flush(__ballot_sync(0xffffffff, cnt > 0));
            // publish this item's list: part[split][query][k]
            const int64_t qrow = static_cast<int64_t>(qt) * kBM + trow;
            if (qrow < Q) {
                float* ps = part_scores + (static_cast<int64_t>(split) * Q + qrow) * k;
                int32_t* pi = part_ids + (static_cast<int64_t>(split) * Q + qrow) * k;
                if (KR > 0) {
#pragma unroll
                    for (int i = 0; i < KRA; ++i)
                        if (i < k) {
                            ps[i] = rs[i];
                            pi[i] = ri[i];
                        }
                } else {
                    for (int i = 0; i < k; ++i) {
                        ps[i] = ls[i];
                        pi[i] = li[i];
                    }
                }
            }
        }
    }
    if constexpr (kPair) pipe2_teardown(sm, warp, tmem_base);
    else pipe_teardown(sm, warp, tmem_base);
}

// ----------------------------------------------------------------------------- k-way merge
// Every input list is sorted by (score desc, id asc). List g of query q starts at
// scores + g * stride_s + q * k (ids likewise with stride_i), so the same kernels merge the search
// kernel's per-split lists and the per-rank records of an all-gather.
//
// One 128-thread block per query copies its G*k entries to shared memory in one sweep of
// independent loads. Then, normally, a pruned rank-by-counting merge: the k-th best of the lists'
// first ceil(k/G) entries is a lower bound tau of the global k-th best; the few entries not behind
// tau (>= k of them, usually < 2k) are compacted and each finds its output slot by counting the
// candidates ahead of it — no serial pick-the-best rounds. If the candidate set is large (skewed
// lists) or the lists are too short to give a bound, a pairwise tree merge takes over:
// ceil(log2 G) levels, ping-pong between two buffers, every entry locating its slot by a binary
// search in the partner list.
constexpr int kMergeThreads = 128;
constexpr int kMergeCandCap = 512;

template <typename IdT>
__device__ __forceinline__ bool sorts_ahead(float sa, IdT ia, float sb, IdT ib) {
    return sa > sb || (sa == sb && ia < ib);
}

// Merge the G lists of query q; called by all kMergeThreads threads of a block. Ends without a
// barrier: a caller that loops over queries must __syncthreads() between calls.
template <typename IdT>
__device__ __forceinline__ void merge_one_query(uint8_t* s_merge, const float* __restrict__ scores,
                                                const IdT* __restrict__ ids, int G, int64_t stride_s, int64_t stride_i,
                                                int64_t q, int k, int64_t id_offset, float* __restrict__ out_scores,
                                                int64_t* __restrict__ out_ids) {
    const int tid = threadIdx.x;
    const int n = G * k, half = ((G + 1) / 2) * k;
    // buffers: ids A[n], ids B[half], scores A[n], scores B[half], list lengths A[G], B[(G+1)/2]
    IdT* idA = reinterpret_cast<IdT*>(s_merge);
    IdT* idB = idA + n;
    float* scA = reinterpret_cast<float*>(idB + half);
    float* scB = scA + n;
    int* lenA = reinterpret_cast<int*>(scB + half);
    int* lenB = lenA + G;
    int* offs = lenB + (G + 1) / 2;  // [G + 1] candidate counts -> offsets (pruned path)
    __shared__ int sh_tau, sh_total;
    if (tid == 0) sh_tau = -1;
    {
        constexpr int U = 8;  // independent loads in flight per thread
        for (int e0 = 0; e0 < n; e0 += kMergeThreads * U) {
            float fs[U];
            IdT fi[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * kMergeThreads + tid;
                if (e < n) {
                    const int g = e / k, p = e - g * k;
                    fs[u] = scores[g * stride_s + q * k + p];
                    fi[u] = ids[g * stride_i + q * k + p];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * kMergeThreads + tid;
                if (e < n) {
                    scA[e] = fs[u];
                    idA[e] = fi[u];
                }
            }
        }
    }
    __syncthreads();
    // valid length of every list (unused slots, id < 0, sit at the tail)
    for (int g = tid; g < G; g += kMergeThreads) {
        int lo = 0, hi = k;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (idA[g * k + mid] >= 0) lo = mid + 1;
            else hi = mid;
        }
        lenA[g] = lo;
    }
    __syncthreads();
    const int m = (k + G - 1) / G;
    if (G > 1 && G * m <= half) {
        // tau = the entry of rank k-1 among the first m entries of every list (set S). S is copied
        // to a flat array first so that the rank count is one pipelined sweep.
        const int ns = G * m;
        for (int e = tid; e < ns; e += kMergeThreads) {
            const int g = e / m, p = e - g * m;
            const bool valid = p < lenA[g];
            scB[e] = valid ? scA[g * k + p] : -INFINITY;
            idB[e] = valid ? idA[g * k + p] : static_cast<IdT>(-1);
        }
        __syncthreads();
        for (int e = tid; e < ns; e += kMergeThreads) {
            const float s = scB[e];
            const IdT id = idB[e];
            if (id < 0) continue;
            int ahead = 0;
#pragma unroll 4
            for (int j = 0; j < ns; ++j) ahead += sorts_ahead<IdT>(scB[j], idB[j], s, id) ? 1 : 0;
            if (ahead == k - 1) sh_tau = (e / m) * k + (e - (e / m) * m);
        }
        __syncthreads();
        const int tau = sh_tau;
        if (tau >= 0) {  // block-uniform
            const float ts = scA[tau];
            const IdT ti = idA[tau];
            for (int g = tid; g < G; g += kMergeThreads) {  // entries of list g not behind tau
                int lo = 0, hi = lenA[g];
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (!sorts_ahead<IdT>(ts, ti, scA[g * k + mid], idA[g * k + mid])) lo = mid + 1;
                    else hi = mid;
                }
                offs[g] = lo;
            }
            __syncthreads();
            if (tid < 32) {  // exclusive scan of the counts
                int run = 0;
                for (int b0 = 0; b0 < G; b0 += 32) {
                    const int g = b0 + tid;
                    const int c = g < G ? offs[g] : 0;
                    int incl = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int up = __shfl_up_sync(0xffffffff, incl, o);
                        if (tid >= o) incl += up;
                    }
                    if (g < G) offs[g] = run + incl - c;
                    run += __shfl_sync(0xffffffff, incl, 31);
                }
                if (tid == 0) {
                    offs[G] = run;
                    sh_total = run;
                }
            }
            __syncthreads();
            const int C = sh_total;
            if (C <= kMergeCandCap && C <= half) {  // block-uniform; C >= k by construction
                for (int g = tid; g < G; g += kMergeThreads) {
                    const int o = offs[g], c = offs[g + 1] - o;
                    for (int p = 0; p < c; ++p) {
                        scB[o + p] = scA[g * k + p];
                        idB[o + p] = idA[g * k + p];
                    }
                }
                __syncthreads();
                for (int e = tid; e < C; e += kMergeThreads) {
                    const float s = scB[e];
                    const IdT id = idB[e];
                    int ahead = 0;
#pragma unroll 4
                    for (int j = 0; j < C; ++j) ahead += sorts_ahead<IdT>(scB[j], idB[j], s, id) ? 1 : 0;
                    if (ahead < k) {
                        out_scores[q * k + ahead] = s;
                        out_ids[q * k + ahead] = static_cast<int64_t>(id) + id_offset;
                    }
                }
                return;
            }
        }
    }
    IdT *idS = idA, *idD = idB;
    float *scS = scA, *scD = scB;
    int *lenS = lenA, *lenD = lenB;
    int L = G;
    while (L > 1) {
        const int Lo = (L + 1) / 2;
        for (int e = tid; e < L * k; e += kMergeThreads) {
            const int g = e / k, p = e - g * k;
            if (p >= lenS[g]) continue;
            const float s = scS[e];
            const IdT id = idS[e];
            int dst = p;
            const int pg = g ^ 1;
            if (pg < L) {  // entries of the partner list that sort ahead of this one
                const float* ps = scS + pg * k;
                const IdT* pi = idS + pg * k;
                int lo = 0, hi = lenS[pg];
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (sorts_ahead<IdT>(ps[mid], pi[mid], s, id)) lo = mid + 1;
                    else hi = mid;
                }
                dst += lo;
            }
            if (dst < k) {
                scD[(g >> 1) * k + dst] = s;
                idD[(g >> 1) * k + dst] = id;
            }
        }
        for (int g = tid; g < Lo; g += kMergeThreads) {
            const int both = lenS[2 * g] + (2 * g + 1 < L ? lenS[2 * g + 1] : 0);
            lenD[g] = both < k ? both : k;
        }
        __syncthreads();
        IdT* ti = idS; idS = idD; idD = ti;
        float* ts = scS; scS = scD; scD = ts;
        int* tl = lenS; lenS = lenD; lenD = tl;
        L = Lo;
    }
    const int len = lenS[0];
    for (int r = tid; r < k; r += kMergeThreads) {
        out_scores[q * k + r] = r < len ? scS[r] : -INFINITY;
        out_ids[q * k + r] = r < len ? static_cast<int64_t>(idS[r]) + id_offset : -1;
    }
}

template <typename IdT>
__global__ void __launch_bounds__(kMergeThreads)
topk_tree_merge_kernel(const float* __restrict__ scores, const IdT* __restrict__ ids, int G, int64_t stride_s,
                       int64_t stride_i, int64_t Q, int k, int64_t id_offset, float* __restrict__ out_scores,
                       int64_t* __restrict__ out_ids) {
    extern __shared__ __align__(16) uint8_t s_merge[];
    merge_one_query<IdT>(s_merge, scores, ids, G, stride_s, stride_i, static_cast<int64_t>(blockIdx.x), k, id_offset,
                         out_scores, out_ids);
}

// ----------------------------------------------------------------------------- peer-memory exchange + merge
// The multi-GPU exchange step as ONE kernel per rank instead of an all-gather followed by a merge:
// every rank owns an exchange buffer that its peers have mapped (CUDA IPC over NVLink/NVSwitch):
//   [header: flags[2][16], done, epoch] [parity 0: G slots] [parity 1: G slots]
// The kernel (a) stores this rank's record into slot[rank] of EVERY rank's buffer with plain
// peer stores, (b) its last block to finish publishes flags[parity][rank] = epoch in every buffer
// (release, system scope), (c) every block waits until all G flags of its own buffer show the
// epoch (acquire), (d) the blocks merge the G records where they landed. The epoch lives in the
// buffer and advances by one per launch, so a captured CUDA graph replays correctly; slots are
// double-buffered by epoch parity, and a rank cannot run two epochs ahead of a peer because it
// needs that peer's flag of the epoch in between.
constexpr int kExchMaxRanks = 16;
constexpr size_t kExchHeaderBytes = 1024;
struct ExchangeHeader {
    uint32_t flags[2][kExchMaxRanks];
    uint32_t done;
    uint32_t epoch;
    uint32_t error;  // sticky: 1 + the first peer whose record did not arrive within the timeout
};
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(kMergeThreads)
topk_exchange_merge_kernel(const uint8_t* __restrict__ local_record, uint8_t* const* __restrict__ peer_bufs, int rank,
                           int G, int64_t Q, int k, size_t rec_bytes, size_t ids_off, size_t slot_bytes,
                           float* __restrict__ out_scores, int64_t* __restrict__ out_ids, uint64_t timeout_ns) {
    extern __shared__ __align__(16) uint8_t s_merge[];
    const int tid = threadIdx.x;
    uint8_t* mine_raw = peer_bufs[rank];
    ExchangeHeader* mine = reinterpret_cast<ExchangeHeader*>(mine_raw);
    const uint32_t epoch = ld_acquire_sys(&mine->epoch) + 1u;  // advanced by the last block below
    const size_t parity = epoch & 1u;
    // (a) push the record to every rank (own buffer included)
    const size_t words = rec_bytes / 8;
    const uint2* src = reinterpret_cast<const uint2*>(local_record);
    for (int p = 0; p < G; ++p) {
        uint2* dst = reinterpret_cast<uint2*>(peer_bufs[p] + kExchHeaderBytes + (parity * G + rank) * slot_bytes);
        for (size_t i = static_cast<size_t>(blockIdx.x) * kMergeThreads + tid; i < words;
             i += static_cast<size_t>(gridDim.x) * kMergeThreads)
            dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    // (b) last block of this rank: advance the epoch, raise this rank's flag everywhere
    if (tid == 0) {
        const uint32_t prev = atomicAdd(&mine->done, 1u);
        if (prev == gridDim.x - 1) {
            mine->done = 0;
            mine->epoch = epoch;
            __threadfence_system();  // one release fence for all G flag stores (G releases would pay G round trips)
            for (int p = 0; p < G; ++p)
                st_relaxed_sys(&reinterpret_cast<ExchangeHeader*>(peer_bufs[p])->flags[parity][rank], epoch);
        }
    }
    // (c) all records of this epoch have landed in my buffer. The wait is bounded: a peer that died
    // (or never made the call) must not hang this rank's stream forever. On a timeout the kernel
    // records the missing peer in the header (sticky, read by arb_topk_exchange_status), returns
    // "no result" rows (-inf / -1) and later calls return at once.
    int lost = 0;
    if (tid < G) {
        if (ld_acquire_sys(&mine->error) != 0u) {
            lost = 1;
        } else {
            const uint64_t t0 = global_timer_ns();
            uint32_t polls = 0;
            while (ld_acquire_sys(&mine->flags[parity][tid]) != epoch) {
                if ((++polls & 1023u) == 0u && global_timer_ns() - t0 > timeout_ns) {
                    atomicCAS(&mine->error, 0u, 1u + static_cast<uint32_t>(tid));
                    lost = 1;
                    break;
                }
            }
        }
    }
    if (__syncthreads_or(lost)) {
        for (int64_t q = blockIdx.x; q < Q; q += gridDim.x)
            for (int j = tid; j < k; j += kMergeThreads) {
                out_scores[q * k + j] = -INFINITY;
                out_ids[q * k + j] = -1;
            }
        return;
    }
    // (d) merge them
    const uint8_t* base = mine_raw + kExchHeaderBytes + parity * G * slot_bytes;
    for (int64_t q = blockIdx.x; q < Q; q += gridDim.x) {
        merge_one_query<int64_t>(s_merge, reinterpret_cast<const float*>(base),
                                 reinterpret_cast<const int64_t*>(base + ids_off), G,
                                 static_cast<int64_t>(slot_bytes / 4), static_cast<int64_t>(slot_bytes / 8), q, k, 0,
                                 out_scores, out_ids);
        __syncthreads();
    }
}

// Fallback when G*k entries do not fit in shared memory: one warp per query, lane l owns lists
// l, l+32, ... and offers the best of their heads each round; a warp arg-max picks the winner.
template <typename IdT>
__global__ void __launch_bounds__(256)
topk_merge_kernel(const float* __restrict__ scores, const IdT* __restrict__ ids, int G, int64_t stride_s,
                  int64_t stride_i, int64_t Q, int k, int64_t id_offset, float* __restrict__ out_scores,
                  int64_t* __restrict__ out_ids) {
    extern __shared__ int s_pos_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* pos = s_pos_all + warp * G;
    const int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp;
    if (q >= Q) return;
    for (int g = lane; g < G; g += 32) pos[g] = 0;
    __syncwarp();
    for (int r = 0; r < k; ++r) {
        float bs = -INFINITY;
        int64_t bi = INT64_MAX;
        int bg = -1;
        for (int g = lane; g < G; g += 32) {
            const int p = pos[g];
            if (p < k) {
                const float s = scores[g * stride_s + q * k + p];
                const int64_t id = static_cast<int64_t>(ids[g * stride_i + q * k + p]);
                if (id >= 0 && (bg < 0 || s > bs || (s == bs && id < bi))) {
                    bs = s;
                    bi = id;
                    bg = g;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffff, bs, o);
            const int64_t oi = __shfl_xor_sync(0xffffffff, bi, o);
            const int og = __shfl_xor_sync(0xffffffff, bg, o);
            if (og >= 0 && (bg < 0 || os > bs || (os == bs && oi < bi))) {
                bs = os;
                bi = oi;
                bg = og;
            }
        }
        if (lane == 0) {
            out_scores[q * k + r] = bg >= 0 ? bs : -INFINITY;
            out_ids[q * k + r] = bg >= 0 ? bi + id_offset : -1;
        }
        if (bg >= 0 && (bg & 31) == lane) pos[bg] += 1;
        __syncwarp();
    }
}

template <typename IdT>
static int launch_merge_impl(const float* scores, const IdT* ids, int G, int64_t stride_s, int64_t stride_i,
                             int64_t Q, int k, int64_t id_offset, float* out_scores, int64_t* out_ids,
                             cudaStream_t stream) {
    constexpr size_t kMergeSmemMax = 200 * 1024;
    const size_t n = static_cast<size_t>(G) * k, half = static_cast<size_t>((G + 1) / 2) * k;
    const size_t tree = (n + half) * (sizeof(IdT) + 4) + (static_cast<size_t>(G) * 2 + (G + 1) / 2 + 1) * 4;
    if (tree <= kMergeSmemMax) {
        ARB_REQUIRE(Q < (1ll << 31), "topk_merge: Q too large");
        auto kern = topk_tree_merge_kernel<IdT>;
        if (tree > 48 * 1024)
            ARB_CHECK_CUDA(set_max_smem_once(kern, static_cast<int>(tree)));
        kern<<<static_cast<int>(Q), kMergeThreads, tree, stream>>>(scores, ids, G, stride_s, stride_i, Q, k, id_offset,
                                                                   out_scores, out_ids);
    } else {
        int warps = 8;
        while (warps > 1 && static_cast<size_t>(warps) * G * 4 > 48 * 1024) warps >>= 1;
        ARB_REQUIRE(static_cast<size_t>(warps) * G * 4 <= 48 * 1024, "topk_merge: too many lists (G=%d)", G);
        const int64_t blocks = (Q + warps - 1) / warps;
        ARB_REQUIRE(blocks < (1ll << 31), "topk_merge: Q too large");
        topk_merge_kernel<IdT><<<static_cast<int>(blocks), warps * 32, static_cast<size_t>(warps) * G * 4, stream>>>(
            scores, ids, G, stride_s, stride_i, Q, k, id_offset, out_scores, out_ids);
    }
    ARB_CHECK_CUDA(cudaGetLastError());
    return ARB_OK;
}

int launch_topk_merge(const float* scores, const int64_t* ids, int G, int64_t Q, int k,
                      float* out_scores, int64_t* out_ids, cudaStream_t stream) {
    ARB_REQUIRE(scores && ids && out_scores && out_ids, "topk_merge: null pointer");
    ARB_REQUIRE(G > 0 && Q > 0 && k > 0, "topk_merge: bad shape G=%d Q=%lld k=%d", G, (long long)Q, k);
    return launch_merge_impl<int64_t>(scores, ids, G, Q * k, Q * k, Q, k, 0, out_scores, out_ids, stream);
}

// A "record" is one rank's [Q,k] result in one buffer — float32 scores, then (8-byte aligned)
// int64 ids — so that a single all-gather moves both; records are merged where they land.
size_t topk_record_ids_offset(int64_t Q, int k) { return (static_cast<size_t>(Q) * k * 4 + 7) / 8 * 8; }
size_t topk_record_bytes(int64_t Q, int k) { return topk_record_ids_offset(Q, k) + static_cast<size_t>(Q) * k * 8; }

int launch_topk_merge_records(const void* records, int G, int64_t Q, int k, float* out_scores,
                              int64_t* out_ids, cudaStream_t stream) {
    ARB_REQUIRE(records && out_scores && out_ids, "topk_merge_records: null pointer");
    ARB_REQUIRE(G > 0 && Q > 0 && k > 0, "topk_merge_records: bad shape G=%d Q=%lld k=%d", G, (long long)Q, k);
    ARB_REQUIRE((reinterpret_cast<uintptr_t>(records) & 7) == 0, "topk_merge_records: records must be 8-byte aligned");
    const size_t rec = topk_record_bytes(Q, k);
    const uint8_t* base = static_cast<const uint8_t*>(records);
    return launch_merge_impl<int64_t>(reinterpret_cast<const float*>(base),
                                      reinterpret_cast<const int64_t*>(base + topk_record_ids_offset(Q, k)), G,
                                      static_cast<int64_t>(rec / 4), static_cast<int64_t>(rec / 8), Q, k, 0, out_scores,
                                      out_ids, stream);
}

size_t topk_exchange_bytes(int G, size_t slot_bytes) { return kExchHeaderBytes + 2 * static_cast<size_t>(G) * slot_bytes; }

// Host read of the sticky error word of this rank's exchange buffer (synchronises the device).
int topk_exchange_status(const void* own_buf_dev) {
    ARB_REQUIRE(own_buf_dev != nullptr, "topk_exchange_status: null buffer");
    ExchangeHeader h;
    ARB_CHECK_CUDA(cudaMemcpy(&h, own_buf_dev, sizeof(h), cudaMemcpyDeviceToHost));
    if (h.error != 0u) {
        set_error("topk_exchange_merge: the record of rank %u did not arrive within the timeout (ARB_EXCHANGE_TIMEOUT_MS); "
                  "results since then are empty — rebuild the exchange", h.error - 1u);
        return ARB_ERR_CUDA;
    }
    return ARB_OK;
}

int launch_topk_exchange_merge(const void* local_record, void* const* peer_bufs_dev, int rank, int G, int64_t Q,
                               int k, size_t slot_bytes, float* out_scores, int64_t* out_ids, cudaStream_t stream) {
    ARB_REQUIRE(local_record && peer_bufs_dev && out_scores && out_ids, "topk_exchange_merge: null pointer");
    ARB_REQUIRE(G > 1 && G <= kExchMaxRanks && rank >= 0 && rank < G, "topk_exchange_merge: bad rank %d of %d", rank, G);
    ARB_REQUIRE(Q > 0 && k > 0, "topk_exchange_merge: bad shape Q=%lld k=%d", (long long)Q, k);
    const size_t rec = topk_record_bytes(Q, k);
    ARB_REQUIRE(slot_bytes % 8 == 0 && rec <= slot_bytes, "topk_exchange_merge: record (%zu B) exceeds the slot (%zu B)", rec,
                slot_bytes);
    ARB_REQUIRE((reinterpret_cast<uintptr_t>(local_record) & 7) == 0, "topk_exchange_merge: record must be 8-byte aligned");
    const size_t n = static_cast<size_t>(G) * k, half = static_cast<size_t>((G + 1) / 2) * k;
    const size_t smem = (n + half) * (8 + 4) + (static_cast<size_t>(G) * 2 + (G + 1) / 2 + 1) * 4;
    if (smem > 200 * 1024) {
        set_error("topk_exchange_merge: G*k = %zu entries do not fit in shared memory", n);
        return ARB_ERR_UNSUPPORTED;
    }
    auto kern = topk_exchange_merge_kernel;
    if (smem > 48 * 1024)
        ARB_CHECK_CUDA(set_max_smem_once(kern, static_cast<int>(smem)));
    // every block spins on the flags, so all of them must be resident at once: at most one per SM
    const int grid = static_cast<int>(Q < num_sms() ? Q : num_sms());
    static const uint64_t timeout_ns = []() {
        const char* e = getenv("ARB_EXCHANGE_TIMEOUT_MS");
        const long ms = e ? atol(e) : 10000;
        return static_cast<uint64_t>(ms > 0 ? ms : 10000) * 1000000ull;
    }();
    kern<<<grid, kMergeThreads, smem, stream>>>(static_cast<const uint8_t*>(local_record),
                                                reinterpret_cast<uint8_t* const*>(peer_bufs_dev), rank, G, Q, k, rec,
                                                topk_record_ids_offset(Q, k), slot_bytes, out_scores, out_ids, timeout_ns);
    ARB_CHECK_CUDA(cudaGetLastError());
    return ARB_OK;
}

// ----------------------------------------------------------------------------- bf16 search
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// A query tile that hangs over the end of the query matrix is filled in by TMA's out-of-bounds
// path, and that path is slow: measured on a 5 M-row corpus, Q=1 takes 1.65 ms per pass with the
// hardware zero fill and 1.20 ms (6.4 TB/s, 98 % of the HBM peak) when the same zeros are real rows.
// So queries are copied into a buffer padded to whole tiles whenever Q is not a multiple of the
// tile (and small enough for the copy not to matter).
static int64_t padded_queries(int64_t Q, bool pair) {
    const int64_t tile = pair ? 2 * kBM : kBM;
    return (Q % tile != 0 && Q <= 16384) ? (Q + tile - 1) / tile * tile : 0;  // 0: use the caller's matrix
}

__global__ void __launch_bounds__(256)
pad_queries_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t valid, int64_t total) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        dst[i] = i < valid ? __ldg(src + i) : make_uint4(0, 0, 0, 0);
}

size_t search_workspace_bytes(int64_t Q, int64_t N, int D, int k) {
    if (Q <= 0 || N <= 0 || k <= 0 || D <= 0) return 0;
    const SearchPlan p = make_plan(Q, N, k);
    const size_t per = static_cast<size_t>(p.nsplit) * Q * k;
    const size_t pad = static_cast<size_t>(padded_queries(Q, p.pair)) * D * 2;
    const size_t prog = static_cast<size_t>(p.nq) * p.nsplit * sizeof(int);
    return align_up(per * 4, 256) + align_up(per * 4, 256) + align_up(pad, 256) + align_up(prog, 256);
}

// q / corpus: rows of D 16-bit words (bf16 values, or — tf32 — the two halves of D/2 fp32 values).
static int launch_search_16(const uint16_t* q, const uint16_t* corpus, int64_t Q, int64_t N,
                            int D, int k, float* out_scores, int64_t* out_ids, int64_t id_offset,
                            void* workspace, size_t ws_bytes, bool tf32, cudaStream_t stream) {
    ARB_REQUIRE(q && corpus && out_scores && out_ids, "search: null pointer");
    ARB_REQUIRE(Q > 0 && N > 0, "search: empty problem Q=%lld N=%lld", (long long)Q, (long long)N);
    ARB_REQUIRE(D > 0 && D % 8 == 0, "search: D=%d must be a positive multiple of 8", D);
    ARB_REQUIRE(k > 0 && k <= kMaxK, "search: k=%d out of range [1,%d]", k, kMaxK);
    ARB_REQUIRE(N < (1ll << 31) - kSBN, "search: shard too large (N=%lld); shard the corpus", (long long)N);
    ARB_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(corpus) & 15) == 0,
                "search: operands must be 16-byte aligned");
    const size_t need = search_workspace_bytes(Q, N, D, k);
    if (workspace == nullptr || ws_bytes < need) {
        set_error("search: workspace too small (%zu < %zu bytes)", ws_bytes, need);
        return ARB_ERR_WORKSPACE;
    }
    const SearchPlan p = make_plan(Q, N, k);
    const size_t per = static_cast<size_t>(p.nsplit) * Q * k;
    float* part_scores = reinterpret_cast<float*>(workspace);
    int32_t* part_ids = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(workspace) + align_up(per * 4, 256));

    // pacing slots, one per work item (SearchTileIter); ARB_SEARCH_PACE=0 switches the pacing off (A/B)
    int* progress = search_pace_ref() ? reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(workspace) + 2 * align_up(per * 4, 256) +
                                                  align_up(static_cast<size_t>(padded_queries(Q, p.pair)) * D * 2, 256))
                         : nullptr;
    const int64_t Qp = padded_queries(Q, p.pair);
    if (Qp > 0) {  // whole query tiles only: see padded_queries
        uint16_t* qpad = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(workspace) + 2 * align_up(per * 4, 256));
        const int64_t total = Qp * D / 8, valid = Q * D / 8;
        const int blocks = static_cast<int>(total < 148ll * 8 * 256 ? (total + 255) / 256 : 148 * 8);
        pad_queries_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(q), reinterpret_cast<uint4*>(qpad), valid, total);
        ARB_CHECK_CUDA(cudaGetLastError());
        q = qpad;
    }
    CUtensorMap tq, tc;
    if (!make_tmap_bf16_k64(&tq, q, static_cast<uint64_t>(Qp > 0 ? Qp : Q), static_cast<uint64_t>(D), static_cast<uint64_t>(D), kBM) ||
        !make_tmap_bf16_k64(&tc, corpus, static_cast<uint64_t>(N), static_cast<uint64_t>(D), static_cast<uint64_t>(D),
                            p.pair ? kSBN / 2 : kSBN)) {
        set_error("search: cuTensorMapEncodeTiled failed");
        return ARB_ERR_CUDA;
    }
    constexpr int kSmemMax = 227 * 1024;
    // Shared memory after the ring, in (score, id) slots of 8 B per query row: `lk` sorted-list
    // slots (0 when the list lives in registers) + the candidate buffer. The buffer holds `cbuf`
    // candidates (even, <= 32) at a row stride of cbuf + 1 slots; when not even 2 fit, 1 at stride 1.
    auto buffer_for = [&](int ring_bytes, int lk) {
        const int fit = (kSmemMax - 1024 - ring_bytes) / (kBM * 8) - lk;  // slots left for the buffer
        const int c = ((fit - 1) & ~1) < 32 ? ((fit - 1) & ~1) : 32;
        return c >= 2 ? c : 1;
    };
    auto launch = [&](auto kern, int ring_bytes, int lk, int cbuf) -> int {
        const int slots = lk + (cbuf > 1 ? cbuf + 1 : cbuf);
        const int smem = ring_bytes + slots * kBM * 8 + 1024;
        ARB_CHECK_CUDA(set_max_smem_once(kern, smem));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(static_cast<unsigned>(p.grid));
        cfg.blockDim = dim3(kSearchThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = p.pair ? 2 : 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        ARB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tq, tc, part_scores, part_ids, Q, N, D, k, p.nq, p.cps, p.nchunks,
                                          p.nsplit, cbuf, progress));
        return ARB_OK;
    };
    int lrc;
    // 32 KB stages for the pair schedule (query tile + half a corpus chunk per CTA): the deepest ring
    // that leaves the lists and a >= 16-candidate buffer in place
    constexpr int kP6 = PipeSmem<kSBN, 6, 0, 2>::kExtraOffset, kP5 = PipeSmem<kSBN, 5, 0, 2>::kExtraOffset,
                  kP4 = PipeSmem<kSBN, 4, 0, 2>::kExtraOffset, kP3 = PipeSmem<kSBN, 3, 0, 2>::kExtraOffset,
                  kP2 = PipeSmem<kSBN, 2, 0, 2>::kExtraOffset;
    constexpr int kRing4 = PipeSmem<kSBN, 4>::kExtraOffset, kRing3 = PipeSmem<kSBN, 3>::kExtraOffset,
                  kRing2 = PipeSmem<kSBN, 2>::kExtraOffset;
    static_assert((kSmemMax - 1024 - kP6) / (kBM * 8) >= 33, "register-list kernels need a 32-candidate buffer");
    static_assert((kSmemMax - 1024 - kRing4) / (kBM * 8) >= 33, "register-list kernels need a 32-candidate buffer");
    if (tf32) {  // fp32 operands (the callers over-select, so k is never tiny: no KR = 10 variant)
        if (p.pair) {
            if (k <= 16) lrc = launch(search_topk_kernel<6, 16, true, true>, kP6, 0, 32);
            else if (buffer_for(kP5, k) >= 16) lrc = launch(search_topk_kernel<5, 0, true, true>, kP5, k, buffer_for(kP5, k));
            else if (buffer_for(kP4, k) >= 16) lrc = launch(search_topk_kernel<4, 0, true, true>, kP4, k, buffer_for(kP4, k));
            else if (buffer_for(kP3, k) >= 16) lrc = launch(search_topk_kernel<3, 0, true, true>, kP3, k, buffer_for(kP3, k));
            else lrc = launch(search_topk_kernel<2, 0, true, true>, kP2, k, buffer_for(kP2, k));
        } else {
            if (k <= 16) lrc = launch(search_topk_kernel<4, 16, false, true>, kRing4, 0, 32);
            else if (buffer_for(kRing3, k) >= 16) lrc = launch(search_topk_kernel<3, 0, false, true>, kRing3, k, buffer_for(kRing3, k));
            else lrc = launch(search_topk_kernel<2, 0, false, true>, kRing2, k, buffer_for(kRing2, k));
        }
    } else if (p.pair) {
        if (k <= 10) lrc = launch(search_topk_kernel<6, 10, true>, kP6, 0, 32);
        else if (k <= 16) lrc = launch(search_topk_kernel<6, 16, true>, kP6, 0, 32);
        else if (buffer_for(kP5, k) >= 16) lrc = launch(search_topk_kernel<5, 0, true>, kP5, k, buffer_for(kP5, k));
        else if (buffer_for(kP4, k) >= 16) lrc = launch(search_topk_kernel<4, 0, true>, kP4, k, buffer_for(kP4, k));
        else if (buffer_for(kP3, k) >= 16) lrc = launch(search_topk_kernel<3, 0, true>, kP3, k, buffer_for(kP3, k));
        else lrc = launch(search_topk_kernel<2, 0, true>, kP2, k, buffer_for(kP2, k));
    } else {
        if (k <= 10) lrc = launch(search_topk_kernel<4, 10, false>, kRing4, 0, 32);
        else if (k <= 16) lrc = launch(search_topk_kernel<4, 16, false>, kRing4, 0, 32);
        else if (buffer_for(kRing3, k) >= 16) lrc = launch(search_topk_kernel<3, 0, false>, kRing3, k, buffer_for(kRing3, k));
        else lrc = launch(search_topk_kernel<2, 0, false>, kRing2, k, buffer_for(kRing2, k));
    }
    if (lrc) return lrc;
    ARB_CHECK_CUDA(cudaGetLastError());
    return launch_merge_impl<int32_t>(part_scores, part_ids, p.nsplit, Q * k, Q * k, Q, k, id_offset, out_scores,
                                      out_ids, stream);
}

int launch_search_bf16(const __nv_bfloat16* q, const __nv_bfloat16* corpus, int64_t Q, int64_t N,
                       int D, int k, float* out_scores, int64_t* out_ids, int64_t id_offset,
                       void* workspace, size_t ws_bytes, cudaStream_t stream) {
    return launch_search_16(reinterpret_cast<const uint16_t*>(q), reinterpret_cast<const uint16_t*>(corpus), Q, N, D, k,
                            out_scores, out_ids, id_offset, workspace, ws_bytes, false, stream);
}

// ----------------------------------------------------------------------------- fp32 search
// Two ways to score fp32 operands on the tensor cores; both over-select, re-score the candidates
// with exact fp32 FMAs and re-rank, so the returned scores are true fp32 dot products.
//
// mode 0 (default): ONE kind::tf32 pass straight over the stored fp32 rows — no converted copy of the
//   corpus, no extra memory. tf32 keeps 10 mantissa bits: |approx - exact| <= 2^-9 |q| |c| for every
//   row (each operand off by < 2^-10 relative, Cauchy-Schwarz). The k + kTf32Margin best approximate
//   scores are re-scored; the result is then PROVABLY the exact top-k iff the exact k-th best score
//   t_k and the worst approximate candidate score a_min satisfy t_k >= a_min + eps, eps = 2^-9 |q|
//   max|c| (no row outside the candidate list can reach t_k). The re-score kernel writes that verdict
//   per query (`unverified`); the host re-runs the few unverified queries through mode 1.
// mode 1 (exact fallback): hi/lo split folded into K,
//   x = hi + lo (both bf16, |x - hi - lo| <= 2^-17 |x|),
//   q.c ~= q_hi.c_hi + q_lo.c_hi + q_hi.c_lo = [q_hi | q_lo | q_hi] . [c_hi | c_hi | c_lo]
//   i.e. one bf16 search with D' = 3D (error ~4e-7 on unit vectors of dim 768, well inside the 1e-5
//   tie tolerance) at 3x the MMA work and a 1.5x copy of the corpus in the workspace.
constexpr int kF32Margin = 8;     // mode 1
constexpr int kTf32Margin = 22;   // mode 0: k = 10 -> 32 candidates

template <bool kIsQuery>
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t rows, int D) {
    const int64_t total = rows * (D / 4);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / (D / 4);
        const int c = static_cast<int>(i % (D / 4)) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * D + c));
        const float f[4] = {v.x, v.y, v.z, v.w};
        __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hi[j] = __float2bfloat16_rn(f[j]);
            lo[j] = __float2bfloat16_rn(f[j] - __bfloat162float(hi[j]));
        }
        __nv_bfloat16* o = out + r * 3 * D + c;
        uint2 uh, ul;
        uh.x = (uint32_t)__bfloat16_as_ushort(hi[0]) | ((uint32_t)__bfloat16_as_ushort(hi[1]) << 16);
        uh.y = (uint32_t)__bfloat16_as_ushort(hi[2]) | ((uint32_t)__bfloat16_as_ushort(hi[3]) << 16);
        ul.x = (uint32_t)__bfloat16_as_ushort(lo[0]) | ((uint32_t)__bfloat16_as_ushort(lo[1]) << 16);
        ul.y = (uint32_t)__bfloat16_as_ushort(lo[2]) | ((uint32_t)__bfloat16_as_ushort(lo[3]) << 16);
        *reinterpret_cast<uint2*>(o) = uh;
        *reinterpret_cast<uint2*>(o + D) = kIsQuery ? ul : uh;
        *reinterpret_cast<uint2*>(o + 2 * D) = kIsQuery ? uh : ul;
    }
}

// One warp per query: exact fp32 dot for each candidate, then rank by (score desc, id asc).
// verify_scale > 0 (mode 0): also decide whether the candidate list provably contains the exact
// top-k: unverified[q] = 1 iff the list is full and t_k < a_min + verify_scale * |q|.
__global__ void __launch_bounds__(128)
rescore_f32_kernel(const float* __restrict__ q, const float* __restrict__ corpus, int64_t Q, int D,
                   int kc, int k, const float* __restrict__ cand_scores, const int64_t* __restrict__ cand_ids,
                   int64_t id_offset, float* __restrict__ out_scores, int64_t* __restrict__ out_ids,
                   float verify_scale, int32_t* __restrict__ unverified) {
    extern __shared__ uint8_t smem_rs[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float* sc = reinterpret_cast<float*>(smem_rs) + warp * kc;
    int64_t* si = reinterpret_cast<int64_t*>(smem_rs + static_cast<size_t>(nw) * kc * 4 +
                                             (static_cast<size_t>(nw) * kc * 4) % 8) + warp * kc;
    const int64_t qi = static_cast<int64_t>(blockIdx.x) * nw + warp;
    if (qi >= Q) return;
    const float* qr = q + qi * D;
    float qn2 = 0.f;  // |q|^2 (mode 0)
    if (verify_scale > 0.f) {
        for (int d = lane * 4; d < D; d += 128) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(qr + d));
            qn2 = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, qn2))));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) qn2 += __shfl_xor_sync(0xffffffff, qn2, o);
    }
    for (int c = 0; c < kc; ++c) {
        const int64_t id = cand_ids[qi * kc + c];  // already offset by id_offset
        float acc = 0.f;
        if (id >= 0) {
            const float* cr = corpus + (id - id_offset) * D;
            for (int d = lane * 4; d < D; d += 128) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(qr + d));
                const float4 b = __ldg(reinterpret_cast<const float4*>(cr + d));
                acc = fmaf(a.x, b.x, acc);
                acc = fmaf(a.y, b.y, acc);
                acc = fmaf(a.z, b.z, acc);
                acc = fmaf(a.w, b.w, acc);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffff, acc, o);
        }
        if (lane == 0) {
            sc[c] = id >= 0 ? acc : -INFINITY;
            si[c] = id;
        }
    }
    __syncwarp();
    float tk = INFINITY;  // exact score that lands at rank k - 1
    for (int c = lane; c < kc; c += 32) {
        const float s = sc[c];
        const int64_t id = si[c];
        int rank = 0;
        for (int j = 0; j < kc; ++j) {
            const float sj = sc[j];
            const int64_t ij = si[j];
            // candidates that sort strictly ahead of c; empty slots (id < 0) sort last
            const bool ahead = (ij >= 0 && id < 0) || ((ij >= 0) == (id >= 0) && (sj > s || (sj == s && (ij < id || (ij == id && j < c)))));
            rank += ahead ? 1 : 0;
        }
        if (rank < k) {
            out_scores[qi * k + rank] = s;
            out_ids[qi * k + rank] = id;
        }
        if (rank == k - 1) tk = s;
    }
    if (unverified != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tk = fminf(tk, __shfl_xor_sync(0xffffffff, tk, o));
        if (lane == 0) {
            // the approximate lists are sorted: the last slot is the worst candidate; an unused last
            // slot means every corpus row is a candidate
            const bool full = cand_ids[qi * kc + kc - 1] >= 0;
            const float a_min = cand_scores[qi * kc + kc - 1];
            unverified[qi] = (full && !(tk >= a_min + verify_scale * sqrtf(qn2))) ? 1 : 0;
        }
    }
}

static int f32_kc(int k, int mode) {
    const int kc = k + (mode == 0 ? kTf32Margin : kF32Margin);
    return kc < kMaxK ? kc : kMaxK;
}

static void f32_layout(int64_t Q, int64_t N, int D, int k, int mode, size_t* off_q, size_t* off_c, size_t* off_cs,
                       size_t* off_ci, size_t* off_inner, size_t* total) {
    const int kc = f32_kc(k, mode);
    size_t o = 0;
    *off_q = o;  o += mode == 0 ? 0 : align_up(static_cast<size_t>(Q) * 3 * D * 2, 256);
    *off_c = o;  o += mode == 0 ? 0 : align_up(static_cast<size_t>(N) * 3 * D * 2, 256);
    *off_cs = o; o += align_up(static_cast<size_t>(Q) * kc * 4, 256);
    *off_ci = o; o += align_up(static_cast<size_t>(Q) * kc * 8, 256);
    *off_inner = o; o += search_workspace_bytes(Q, N, mode == 0 ? 2 * D : 3 * D, kc);
    *total = o;
}

size_t search_f32_workspace_bytes(int64_t Q, int64_t N, int D, int k, int mode) {
    if (Q <= 0 || N <= 0 || k <= 0 || D <= 0 || mode < 0 || mode > 1) return 0;
    size_t a, b, c, d, e, t;
    f32_layout(Q, N, D, k, mode, &a, &b, &c, &d, &e, &t);
    return t;
}

int launch_search_f32(const float* q, const float* corpus, int64_t Q, int64_t N, int D, int k,
                      float corpus_max_norm, float* out_scores, int64_t* out_ids, int64_t id_offset,
                      int32_t* unverified, int mode, void* workspace, size_t ws_bytes, cudaStream_t stream) {
    ARB_REQUIRE(q && corpus && out_scores && out_ids, "search_f32: null pointer");
    ARB_REQUIRE(Q > 0 && N > 0, "search_f32: empty problem Q=%lld N=%lld", (long long)Q, (long long)N);
    ARB_REQUIRE(D > 0 && D % 8 == 0, "search_f32: D=%d must be a positive multiple of 8", D);
    ARB_REQUIRE(k > 0 && k <= kMaxK, "search_f32: k=%d out of range [1,%d]", k, kMaxK);
    ARB_REQUIRE(mode == 0 || mode == 1, "search_f32: mode %d must be 0 (tf32 + verified re-score) or 1 (split-bf16)", mode);
    ARB_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(corpus) & 15) == 0,
                "search_f32: operands must be 16-byte aligned");
    size_t off_q, off_c, off_cs, off_ci, off_inner, total;
    f32_layout(Q, N, D, k, mode, &off_q, &off_c, &off_cs, &off_ci, &off_inner, &total);
    if (workspace == nullptr || ws_bytes < total) {
        set_error("search_f32: workspace too small (%zu < %zu bytes)", ws_bytes, total);
        return ARB_ERR_WORKSPACE;
    }
    const int kc = f32_kc(k, mode);
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    float* cs = reinterpret_cast<float*>(ws + off_cs);
    int64_t* ci = reinterpret_cast<int64_t*>(ws + off_ci);
    int rc;
    if (mode == 0) {
        rc = launch_search_16(reinterpret_cast<const uint16_t*>(q), reinterpret_cast<const uint16_t*>(corpus), Q, N, 2 * D, kc,
                              cs, ci, id_offset, ws + off_inner, ws_bytes - off_inner, true, stream);
    } else {
        __nv_bfloat16* q3 = reinterpret_cast<__nv_bfloat16*>(ws + off_q);
        __nv_bfloat16* c3 = reinterpret_cast<__nv_bfloat16*>(ws + off_c);
        const int blocks = num_sms() * 8;
        split_bf16_kernel<true><<<blocks, 256, 0, stream>>>(q, q3, Q, D);
        split_bf16_kernel<false><<<blocks, 256, 0, stream>>>(corpus, c3, N, D);
        ARB_CHECK_CUDA(cudaGetLastError());
        rc = launch_search_bf16(q3, c3, Q, N, 3 * D, kc, cs, ci, id_offset, ws + off_inner, ws_bytes - off_inner, stream);
    }
    if (rc) return rc;
    const int nw = 4;
    const size_t smem = static_cast<size_t>(nw) * kc * 4 + 8 + static_cast<size_t>(nw) * kc * 8;
    // 2^-9 |q| max|c| bounds the tf32 scoring error of any row (see the header comment); a little
    // slack covers the fp32 accumulation order of the MMA and of the re-score
    const float verify_scale = mode == 0 ? (1.0f / 512.0f) * corpus_max_norm * 1.001f + 1e-6f : 0.f;
    rescore_f32_kernel<<<static_cast<int>((Q + nw - 1) / nw), nw * 32, smem, stream>>>(
        q, corpus, Q, D, kc, k, cs, ci, id_offset, out_scores, out_ids, verify_scale, mode == 0 ? unverified : nullptr);
    ARB_CHECK_CUDA(cudaGetLastError());
    if (mode == 1 && unverified != nullptr) ARB_CHECK_CUDA(cudaMemsetAsync(unverified, 0, static_cast<size_t>(Q) * 4, stream));
    return ARB_OK;
}

}  // namespace arb
