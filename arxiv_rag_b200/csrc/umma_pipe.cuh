// Warp-specialised TMA -> smem ring -> tcgen05.mma -> TMEM pipeline shared by the encoder
// GEMMs and the search score kernel. One CTA per SM:
//   warp 0      TMA producer (one elected lane)
//   warp 1      TMEM owner + MMA issuer (one elected lane)
//   warps 2..   epilogue: TMEM -> registers -> (bias/GELU/residual -> smem -> TMA store | top-k)
// Tiles are 128 (M) x BN (N) x 64 (K-block); A and B are both K-major 16-bit, loaded with the
// 128-byte swizzle. The fp32 accumulator is double buffered in TMEM (2 x BN columns) so the
// epilogue of tile i overlaps the main loop of tile i+1.
#pragma once
#include "ptx.cuh"

namespace arb {

constexpr int kBM = 128;
constexpr int kBK = 64;

// Shared-memory carve-up: [STAGES x (A tile | B tile)] [PRE bytes, 1024-aligned: epilogue staging]
// [barriers] [extra...]
// CTAS = 2: CTA-pair layout, each CTA stages only its half (BN / 2 rows) of the B tile.
template <int BN, int STAGES, int PRE = 0, int CTAS = 1>
struct PipeSmem {
    static constexpr int kBN = BN;
    static constexpr int kStages = STAGES;
    static constexpr int kABytes = kBM * kBK * 2;
    static constexpr int kBBytes = (BN / CTAS) * kBK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kPreOffset = STAGES * kStageBytes;
    static constexpr int kBarOffset = kPreOffset + PRE;
    // full[STAGES] empty[STAGES] tmem_full[2] tmem_empty[2] aux[4] + tmem base ptr
    static constexpr int kNumBars = 2 * STAGES + 8;
    static constexpr int kBarBytes = kNumBars * 8 + 16;
    static constexpr int kExtraOffset = kBarOffset + ((kBarBytes + 127) / 128) * 128;

    uint8_t* base;
    __device__ __forceinline__ uint8_t* a(int s) const { return base + s * kStageBytes; }
    __device__ __forceinline__ uint8_t* b(int s) const { return base + s * kStageBytes + kABytes; }
    __device__ __forceinline__ uint8_t* pre() const { return base + kPreOffset; }
    __device__ __forceinline__ uint64_t* bars() const { return reinterpret_cast<uint64_t*>(base + kBarOffset); }
    __device__ __forceinline__ uint64_t* full(int s) const { return bars() + s; }
    __device__ __forceinline__ uint64_t* empty(int s) const { return bars() + STAGES + s; }
    __device__ __forceinline__ uint64_t* tmem_full(int s) const { return bars() + 2 * STAGES + s; }
    __device__ __forceinline__ uint64_t* tmem_empty(int s) const { return bars() + 2 * STAGES + 2 + s; }
    __device__ __forceinline__ uint64_t* aux(int s) const { return bars() + 2 * STAGES + 4 + s; }
    __device__ __forceinline__ uint32_t* tmem_ptr() const {
        return reinterpret_cast<uint32_t*>(base + kBarOffset + kNumBars * 8);
    }
    __device__ __forceinline__ uint8_t* extra() const { return base + kExtraOffset; }
};

template <int BN>
__host__ __device__ constexpr uint32_t tmem_cols_for() {
    return (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
}

// Barrier init + TMEM allocation; returns the TMEM base address to every thread.
// epi_threads = number of epilogue threads that arrive on tmem_empty per tile.
template <class SM>
__device__ __forceinline__ uint32_t pipe_setup(const SM& sm, int warp, const void* tmap_a,
                                               const void* tmap_b, uint32_t epi_threads) {
    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(tmap_a);
        tma_prefetch_desc(tmap_b);
        for (int s = 0; s < SM::kStages; ++s) {
            mbar_init(sm.full(s), 1);
            mbar_init(sm.empty(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(sm.tmem_full(s), 1);
            mbar_init(sm.tmem_empty(s), epi_threads);
        }
        for (int s = 0; s < 4; ++s) mbar_init(sm.aux(s), 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(sm.tmem_ptr(), tmem_cols_for<SM::kBN>());
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *sm.tmem_ptr();
}

template <class SM>
__device__ __forceinline__ void pipe_teardown(const SM& sm, int warp, uint32_t tmem_base) {
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, tmem_cols_for<SM::kBN>());
    }
}

// Producer: for every tile the iterator yields, stream its K-blocks through the ring.
// TileIter: bool next(int& row_a, int& row_b) — first rows of the A and B tiles.
template <class SM, class TileIter>
// stage_tx: bytes one stage's two loads deliver, when the A box is shorter than the 128-row tile
// (0 = full tiles). The MMA still reads 128 rows; the rows the box leaves untouched are never stored.
__device__ __forceinline__ void pipe_produce(const SM& sm, const void* tmap_a, const void* tmap_b,
                                             TileIter it, int kblocks, uint64_t hint_a,
                                             uint64_t hint_b, uint32_t stage_tx = 0) {
    int stage = 0;
    uint32_t phase = 0;
    int row_a, row_b;
    const uint32_t tx = stage_tx ? stage_tx : static_cast<uint32_t>(SM::kStageBytes);
    while (it.next(row_a, row_b)) {
        for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(sm.empty(stage), phase ^ 1);
            mbar_arrive_expect_tx(sm.full(stage), tx);
            tma_load_2d(tmap_a, sm.full(stage), sm.a(stage), kb * kBK, row_a, hint_a);
            tma_load_2d(tmap_b, sm.full(stage), sm.b(stage), kb * kBK, row_b, hint_b);
            if (++stage == SM::kStages) {
                stage = 0;
                phase ^= 1;
            }
        }
    }
}

// MMA issuer: 4 x (128 x BN x 16) tcgen05.mma per K-block, accumulating in the TMEM buffer
// that the epilogue has released; commits release the smem stage and publish the accumulator.
// idesc: umma_idesc_16bit(kBM, SM::kBN, A format, B format) — a runtime value, so one kernel serves
// every operand-format pair.
// kTf32: the 128-byte rows hold 32 fp32 values and the MMAs are kind::tf32 (4 K-steps of 8) —
// byte layout, swizzle and descriptor stepping are the same as for 64 16-bit values.
template <class SM, bool kTf32 = false, class TileIter>
__device__ __forceinline__ void pipe_mma(const SM& sm, uint32_t tmem_base, TileIter it, int kblocks, uint32_t idesc) {
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int row_a, row_b;
    while (it.next(row_a, row_b)) {
        mbar_wait(sm.tmem_empty(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * SM::kBN;
        for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(sm.full(stage), phase);
            tc_fence_after();
            const uint64_t da = umma_desc_sw128(smem_u32(sm.a(stage)));
            const uint64_t db = umma_desc_sw128(smem_u32(sm.b(stage)));
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
                // +32 bytes per K=16 step inside the 128-byte swizzle atom -> +2 in addr>>4 units
                if constexpr (kTf32) umma_tf32_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                else umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(sm.empty(stage));
            if (++stage == SM::kStages) {
                stage = 0;
                phase ^= 1;
            }
        }
        umma_commit(sm.tmem_full(acc));
        if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cluster of 2, tcgen05 cta_group::2): one 256 x BN tile per pair. Each CTA
// loads its own 128 rows of A and its half of the B rows; every load credits the LEADER's (rank 0)
// full barrier, the leader issues the M = 256 MMAs, and its commits arrive on the empty /
// tmem_full barriers of BOTH CTAs. Each CTA's epilogue drains its own 128 accumulator rows and
// reports to the leader's tmem_empty barrier (one arrival per warp). Compared with two
// independent CTAs this halves the B bytes staged per SM and per MMA, which buys a deeper ring.
// epi_arrivals = arrivals per accumulator hand-back: 2 CTAs x epilogue warps per CTA
template <class SM>
__device__ __forceinline__ uint32_t pipe2_setup(const SM& sm, int warp, const void* tmap_a, const void* tmap_b,
                                                uint32_t epi_arrivals) {
    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(tmap_a);
        tma_prefetch_desc(tmap_b);
        for (int s = 0; s < SM::kStages; ++s) {
            mbar_init(sm.full(s), 1);
            mbar_init(sm.empty(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(sm.tmem_full(s), 1);
            mbar_init(sm.tmem_empty(s), epi_arrivals);
        }
        for (int s = 0; s < 4; ++s) mbar_init(sm.aux(s), 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc_2cta(sm.tmem_ptr(), tmem_cols_for<SM::kBN>());
        tmem_relinquish_2cta();
    }
    tc_fence_before();
    cluster_sync_all();  // both CTAs' barriers and TMEM exist before any cross-CTA arrive / MMA
    tc_fence_after();
    return *sm.tmem_ptr();
}

template <class SM>
__device__ __forceinline__ void pipe2_teardown(const SM& sm, int warp, uint32_t tmem_base) {
    tc_fence_before();
    cluster_sync_all();  // the peer may still be reading operands / receiving commits until here
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc_2cta(tmem_base, tmem_cols_for<SM::kBN>());
    }
}

// Producer (one elected lane in EACH CTA). TileIter: bool next(int& row_a, int& row_b) with
// row_a = this CTA's first A row, row_b = first B row of the pair's tile.
template <class SM, class TileIter>
__device__ __forceinline__ void pipe2_produce(const SM& sm, const void* tmap_a, const void* tmap_b, TileIter it,
                                              int kblocks, int rank, uint64_t hint_a, uint64_t hint_b) {
    int stage = 0;
    uint32_t phase = 0;
    int row_a, row_b;
    while (it.next(row_a, row_b)) {
        for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(sm.empty(stage), phase ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(sm.full(stage), 2 * SM::kStageBytes);
            const uint32_t full_leader = map_to_cta(smem_u32(sm.full(stage)), 0);
            tma_load_2d_2cta(tmap_a, full_leader, sm.a(stage), kb * kBK, row_a, hint_a);
            tma_load_2d_2cta(tmap_b, full_leader, sm.b(stage), kb * kBK, row_b + rank * (SM::kBN / 2), hint_b);
            if (++stage == SM::kStages) {
                stage = 0;
                phase ^= 1;
            }
        }
    }
}

// MMA issuer (leader CTA only): 4 x (256 x BN x 16) per K-block.
template <class SM, bool kTf32 = false, class TileIter>
__device__ __forceinline__ void pipe2_mma(const SM& sm, uint32_t tmem_base, TileIter it, int kblocks, uint32_t idesc) {
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int row_a, row_b;
    while (it.next(row_a, row_b)) {
        // CTA-scope acquire on purpose: the arrivals only order tcgen05.ld completions (made visible
        // by tcgen05.fence), and a cluster-scope acquire would make ptxas invalidate the whole L1
        // (CCTL.IVALL) once per tile — the cache the epilogue warps keep bias / gamma / beta in.
        mbar_wait(sm.tmem_empty(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * SM::kBN;
        for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(sm.full(stage), phase);
            tc_fence_after();
            const uint64_t da = umma_desc_sw128(smem_u32(sm.a(stage)));
            const uint64_t db = umma_desc_sw128(smem_u32(sm.b(stage)));
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
                if constexpr (kTf32) umma_tf32_ss_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                else umma_bf16_ss_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_2cta(sm.empty(stage), 0b11);
            if (++stage == SM::kStages) {
                stage = 0;
                phase ^= 1;
            }
        }
        umma_commit_2cta(sm.tmem_full(acc), 0b11);
        if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
        }
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace arb
