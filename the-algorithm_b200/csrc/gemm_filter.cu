// gemm_filter.cu -- K3: batched candidate filter on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces, for a whole batch of queries at once, the per-row `metric.distance(row, query)` + heap push of
// BruteForceIndex.queryWithDistance (ann/src/main/scala/com/twitter/ann/brute_force/BruteForceIndex.scala:76-86).
// The (queries x rows) score matrix is a dense contraction; it is computed tile by tile in bf16 with fp32
// accumulation and NEVER written to memory: the epilogue compares every score against the query's running
// threshold (held in a register, one TMEM lane = one query) and only the rare survivors are appended to the
// query's candidate pool.  finalize_kernel later rescales the survivors exactly.
//
//   D[q, r] = sum_k Qs[q, k] * Rs[r, k]         Qs = bf16 query operand, Rs = bf16 shadow rows (append_kernels.cu)
//   score s = D, badness g = -s, survive iff g <= tau_q  (tau_q already carries the 2*eps margin)
//
// Tiling: queries are the MMA M dimension (128 per CTA = 128 TMEM lanes), corpus rows the N dimension
// (128 per CTA).  With CG = 2 two CTAs of a cluster pair up (tcgen05 cta_group::2, M = N = 256): each CTA
// stages only its half of both operands and the pair's tensor cores read both halves.  A row tile stays in
// shared memory while every query tile streams past it (queries come from L2), so the shadow matrix is read
// from HBM exactly once per batch.  Both operands are double buffered (TMA -> smem, mbarrier full/empty), the
// accumulator is double buffered in TMEM (2 x N columns), two threads of the leader CTA issue the MMAs of alternate
// query tiles (one issuer alone is instruction-issue-bound, see the MMA section), and 8 or 16 epilogue warps drain one
// accumulator while the next tile's MMAs run.
//
// K is not padded to the 128-byte swizzle width in HBM: a K of e.g. 208 is staged as 3 blocks of 64 (128B
// swizzle) + 1 block of 16 (32B swizzle), each with its own tensor map / UMMA descriptor, TMA zero-filling
// the tail columns.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace b200ann {

namespace {

// warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 second MMA issuer (or spare), 4.. epilogue.  The epilogue runs with EW = 8 or 16 warps (template parameter):
// the filter is a string of dependent compares, and in hit-dense chunks two warps per scheduler leave its fixed latencies
// exposed (IPC 0.38 in ncu), so those launches use four per scheduler (+25 % there); in hit-sparse chunks the epilogue
// keeps up anyway and the extra warps only take issue slots from the MMA warp (-1..5 %), so they use eight.
constexpr int kEpiWarp0 = 4;
constexpr int gemm_threads(int epi_warps) { return 32 * (kEpiWarp0 + epi_warps); }
constexpr int kTileRows = 128;        // rows (and queries) staged per CTA per tile
constexpr int kTauTiles = 16;         // query tiles whose thresholds fit the shared-memory table (b <= 4096 with CG = 2)
constexpr size_t kTauBytes = (size_t)kTauTiles * kTileRows * sizeof(uint32_t);

struct alignas(64) GemmTmaps {
    CUtensorMap q64, q32, q16, r64, r32, r16;
};

struct GemmArgs {
    long long row_begin, row_end;     // corpus rows of this launch
    int b, n_qt;                      // real queries, query tiles of 128*CG
    int kp_mma;                       // K rounded up to 16
    QueryState* qstate;
    entry_t* pool;
    int pool_cap;
    int debug_nohit;
    int nseg;                         // K segments per query tile (1 = the whole K extent is one stage)
    int na_stages, nb_stages;         // query-segment ring slots / row-tile buffers
    int a_resident;                   // the batch is one query tile: its segments are loaded once and never released
    uint32_t a_slot_stride;           // bytes between query-segment slots (1024-aligned)
    int tau_smem;                     // thresholds of all query tiles are staged in shared memory (n_qt <= kTauTiles)
    int mma_warps;                    // 1, or 2: warps 1 and 3 of the leader CTA issue the MMAs of alternate query tiles
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address: "CTA 0 of my pair"

template <int CG>
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    if constexpr (CG == 1) {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                smem_u32(dst)),
            "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
            : "memory");
    } else {
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
                "r"(smem_u32(dst)),
            "l"(map), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
            : "memory");
    }
}

template <int CG>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    if constexpr (CG == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    } else {
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         smem_u32(bar)),
                     "h"((uint16_t)3)
                     : "memory");
    }
}

template <int CG>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 1) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout type [61,64).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFFu);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, float* v) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    *v = __uint_as_float(r);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct KBlocks {
    int nb64, has32, has16;
    __device__ __forceinline__ explicit KBlocks(int kp_mma) {
        nb64 = kp_mma >> 6;
        int rem = kp_mma & 63;
        has32 = rem >= 32;
        has16 = (rem & 31) >= 16;
    }
};

}  // namespace

// SEG = false is the common case (the whole K extent is one query stage, ring of two): segment count, ring size and the
// slot / phase arithmetic are compile-time there, which keeps the single MMA-issuing thread off the critical path (the
// generic runtime version cost the headline shape 4 %).
template <int CG, bool SEED, bool SEG, int EW>
__global__ void __launch_bounds__(gemm_threads(EW), 1)
gemm_filter_kernel(const __grid_constant__ GemmTmaps tm, const GemmArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr int N_TILE = kTileRows * CG;          // MMA N (and M)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x / CG;
    const int n_clusters = gridDim.x / CG;

    const uint32_t tile_bytes = (uint32_t)kTileRows * a.kp_mma * 2;
    const uint32_t tile_stride = (tile_bytes + 1023u) & ~1023u;
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int na = a.na_stages, nb = a.nb_stages, nseg = SEG ? a.nseg : 1;   // na = slots carved out of shared memory
    auto a_slot = [&](uint32_t g) -> uint32_t { return SEG ? g % (uint32_t)na : (g & 1u); };       // !SEG: ring of two
    auto a_phase = [&](uint32_t g) -> uint32_t { return SEG ? (g / (uint32_t)na) & 1u : (g >> 1) & 1u; };
    const bool a_resident = a.a_resident != 0;
    unsigned char* smA0 = base;                                            // na slots of a.a_slot_stride bytes
    unsigned char* smB0 = base + (size_t)na * a.a_slot_stride;             // nb full-K row tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(smB0 + (size_t)nb * tile_stride);
    uint64_t* a_full = bars + 0;     // 8
    uint64_t* a_empty = bars + 8;    // 8
    uint64_t* b_full = bars + 16;    // 3
    uint64_t* b_empty = bars + 19;   // 3
    uint64_t* t_full = bars + 22;    // 2
    uint64_t* t_empty = bars + 24;   // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
    uint32_t* tau_tab = reinterpret_cast<uint32_t*>(bars + 28);   // [kTauTiles][kTileRows], present when a.tau_smem

    const KBlocks kb(a.kp_mma);
    // K segments: the 64-wide blocks are dealt out evenly, the narrow tail blocks (32 / 16 wide) go with the last segment
    const int seg_base = kb.nb64 / nseg, seg_rem = kb.nb64 % nseg;
    auto seg_first64 = [&](int sg) { return sg * seg_base + min(sg, seg_rem); };
    auto seg_n64 = [&](int sg) { return seg_base + (sg < seg_rem ? 1 : 0); };
    const uint32_t tail_bytes = (uint32_t)kTileRows * 2u * (32u * kb.has32 + 16u * kb.has16);
    const long long chunk_rows = a.row_end - a.row_begin;
    const int n_rt = (int)((chunk_rows + N_TILE - 1) / N_TILE);
    constexpr uint32_t kTmemCols = 2 * N_TILE;       // 256 (CG=1) or 512 (CG=2)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.q64) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.r64) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 8; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&t_full[i], 1);
            mbar_init(&t_empty[i], EW * CG);
        }
        for (int i = 0; i < 3; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], a.mma_warps == 2 ? 2 : 1);   // every MMA issuer hands a row tile back
        }
        mbar_fence_init();
    }
    if (warp == 2) {
        if constexpr (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // one operand segment: `n64` 128B-swizzled blocks starting at 64-block `first64`, plus the narrow tail blocks if asked
    auto load_blocks = [&](const CUtensorMap* m64, const CUtensorMap* m32, const CUtensorMap* m16, uint64_t* bar, unsigned char* dst,
                           int first64, int n64, bool with_tail, int coord_row) {
        for (int i = 0; i < n64; ++i) tma_load_2d<CG>(m64, bar, dst + i * 16384, (first64 + i) * 64, coord_row);
        if (with_tail) {
            int off = n64 * 16384, col = kb.nb64 * 64;
            if (kb.has32) {
                tma_load_2d<CG>(m32, bar, dst + off, col, coord_row);
                off += 8192;
                col += 32;
            }
            if (kb.has16) tma_load_2d<CG>(m16, bar, dst + off, col, coord_row);
        }
    };

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            uint32_t ga = 0, it = 0;
            for (int rt = cluster_id; rt < n_rt; rt += n_clusters, ++it) {
                const uint32_t bs = it % nb;
                mbar_wait(&b_empty[bs], ((it / nb) & 1) ^ 1);
                if (leader) mbar_expect_tx(&b_full[bs], tile_bytes * CG);
                load_blocks(&tm.r64, &tm.r32, &tm.r16, &b_full[bs], smB0 + (size_t)bs * tile_stride, 0, kb.nb64, true,
                            (int)(a.row_begin + (long long)rt * N_TILE + rank * kTileRows));
                for (int qt = 0; qt < a.n_qt; ++qt) {
                    if (a_resident && it > 0) continue;           // the only query tile is already resident
                    const int q0 = qt * N_TILE + rank * kTileRows;
                    for (int sg = 0; sg < nseg; ++sg, ++ga) {
                        const uint32_t slot = a_resident ? (uint32_t)sg : a_slot(ga);
                        if (!a_resident) mbar_wait(&a_empty[slot], a_phase(ga) ^ 1);
                        const bool last = sg == nseg - 1;
                        const uint32_t bytes = (uint32_t)seg_n64(sg) * 16384u + (last ? tail_bytes : 0u);
                        if (leader) mbar_expect_tx(&a_full[slot], bytes * CG);
                        load_blocks(&tm.q64, &tm.q32, &tm.q16, &a_full[slot], smA0 + (size_t)slot * a.a_slot_stride, seg_first64(sg),
                                    seg_n64(sg), last, q0);
                    }
                }
            }
        }
    } else if (warp == 1 || (warp == 3 && a.mma_warps == 2)) {
        // ================================ MMA issuer (leader CTA, one thread per issuing warp) ================================
        // ONE thread issuing every tcgen05.mma is itself the limit of this kernel: ncu's instruction sampling of the headline
        // launch showed warp 1 busy 91 % of the time, spread evenly over the ~34 SASS instructions nvcc then needed per
        // MMA (descriptor arithmetic + an ELECT / R2UR.BROADCAST loop per operand; 19 since the descriptors are formed by
        // addition below), 3.7 clocks each -- 447 instructions per 256 x 256 x 208 tile against the 1664 clocks its 13
        // MMAs occupy the tensor pipe, which therefore idled ~15 % even with no hit to handle.  With mma_warps == 2 warps 1 and 3 alternate query tiles: issuer W owns accumulator
        // stage W and query slot W (both are the tile counter's parity), waits and commits on the same barriers as before, and
        // each has two tile times to issue one tile.  tcgen05.commit tracks the MMAs of the EXECUTING thread, so a tile's
        // t_full / a_empty commits stay exact; the row tile is released when both issuers have committed (b_empty counts 2).
        if (leader && lane == 0) {
            const bool dual = !SEG && a.mma_warps == 2;
            const uint32_t my_parity = warp == 3 ? 1u : 0u;
            // instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7, 10), K-major both, N>>3 at 17, M>>4 at 24
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N_TILE >> 3) << 17) | ((uint32_t)(N_TILE >> 4) << 24);
            uint32_t ga = 0, gt = 0, it = 0;
            for (int rt = cluster_id; rt < n_rt; rt += n_clusters, ++it) {
                const uint32_t bs = it % nb;
                mbar_wait(&b_full[bs], (it / nb) & 1);
                const uint32_t b_addr = smem_u32(smB0 + (size_t)bs * tile_stride);
                for (int qt = 0; qt < a.n_qt; ++qt, ++gt) {
                    const uint32_t s = gt & 1;                    // accumulator stage
                    if (dual && s != my_parity) {                 // the other issuer's tile (nseg == 1: one slot per tile)
                        ++ga;
                        continue;
                    }
                    const uint32_t d_tmem = tmem_base + s * N_TILE;
                    uint32_t acc = 0;
                    for (int sg = 0; sg < nseg; ++sg, ++ga) {
                        const uint32_t slot = a_resident ? (uint32_t)sg : a_slot(ga);
                        if (!a_resident) mbar_wait(&a_full[slot], a_phase(ga));
                        else if (it == 0) mbar_wait(&a_full[slot], 0);
                        // the accumulator stage last: its hand-off from the epilogue is the round trip that paces the large
                        // launches, the query slot is ready long before it
                        if (sg == 0) mbar_wait(&t_empty[s], ((gt >> 1) & 1) ^ 1);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(smA0 + (size_t)slot * a.a_slot_stride);
                        const int n64 = seg_n64(sg);
                        const uint32_t b_seg = b_addr + (uint32_t)seg_first64(sg) * 16384u;
                        // descriptors by addition: the start-address field is the low 14 bits (address >> 4; shared memory
                        // is < 256 KB, so the field never carries out): +2 per 16-column step, +1024 per 64-column block
                        uint64_t da = smem_desc(a_addr, 1024, 2), db = smem_desc(b_seg, 1024, 2);
#pragma unroll 1
                        for (int i = 0; i < n64; ++i) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                umma_bf16<CG>(d_tmem, da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), idesc, acc);
                                acc = 1;
                            }
                            da += 1024;
                            db += 1024;
                        }
                        if (sg == nseg - 1) {
                            uint32_t a_off = (uint32_t)n64 * 16384u, b_off = (uint32_t)kb.nb64 * 16384u;
                            if (kb.has32) {
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    umma_bf16<CG>(d_tmem, smem_desc(a_addr + a_off + j * 32, 512, 4),
                                                  smem_desc(b_addr + b_off + j * 32, 512, 4), idesc, acc);
                                    acc = 1;
                                }
                                a_off += 8192;
                                b_off += 8192;
                            }
                            if (kb.has16) {
                                umma_bf16<CG>(d_tmem, smem_desc(a_addr + a_off, 256, 6), smem_desc(b_addr + b_off, 256, 6), idesc, acc);
                                acc = 1;
                            }
                        }
                        if (!a_resident) umma_commit<CG>(&a_empty[slot]);   // this query segment is consumed
                    }
                    umma_commit<CG>(&t_full[s]);    // accumulator ready for the epilogue
                }
                umma_commit<CG>(&b_empty[bs]);      // row tile consumed
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ================================ epilogue: TMEM -> registers -> threshold filter ================================
        // One TMEM lane = one query, so the query's threshold lives in a register and a score costs one compare
        // (the compiler folds 32 compares into a max tree).  Thresholds are constant for the launch; the one for the
        // next query tile is prefetched while the current tile is scanned.  Survivors are rare: a thread parks up to
        // kHitRegs of them in registers, releases the accumulator, reserves pool slots with ONE atomic, and writes
        // the entries one tile later, when the atomic's round trip has long completed.
        constexpr int kHitRegs = 4;
        const int e = warp - kEpiWarp0;
        const int quarter = warp & 3;                 // TMEM lanes 32*quarter .. +31 are the ones this warp may touch
        const int half = e >> 2;                      // which slice of the N columns
        constexpr int COLS_PER_WARP = N_TILE / (EW / 4);
        const int q_lane = (int)rank * kTileRows + quarter * 32 + lane;
        // Thresholds are constant for the launch and a thread meets the same n_qt of them for every row tile: they are staged
        // in shared memory once.  (A global load per tile sat on the scoreboard of the slot-reserving atomic: ncu showed
        // the epilogue's first stall of every tile there, 7 % of its samples in a hit-dense chunk.)
        const bool tau_smem = a.tau_smem != 0;
        const uint32_t tau_addr = smem_u32(tau_tab) + (uint32_t)(quarter * 32 + lane) * 4u;   // + qt * kTileRows * 4
        if (tau_smem) {
            if (half == 0) {
#pragma unroll 1
                for (int qt = 0; qt < a.n_qt; ++qt) {
                    const int q = qt * N_TILE + q_lane;
                    const uint32_t t = q < a.b ? __ldcg(&a.qstate[q].tau_key) : 0u;
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(tau_addr + (uint32_t)qt * (kTileRows * 4u)), "r"(t) : "memory");
                }
            }
            asm volatile("bar.sync 1, %0;" ::"r"(EW * 32) : "memory");
        }
        auto load_tau = [&](int qt) -> uint32_t {
            if (tau_smem) {
                uint32_t t;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(tau_addr + (uint32_t)qt * (kTileRows * 4u)));
                return t;
            }
            const int q = qt * N_TILE + q_lane;
            return q < a.b ? __ldcg(&a.qstate[q].tau_key) : 0u;
        };
        auto emit_now = [&](int q, float v, uint32_t row) {
            uint32_t slot = atomicAdd(&a.qstate[q].pool_count, 1u);
            if (slot < (uint32_t)a.pool_cap) a.pool[(size_t)q * a.pool_cap + slot] = make_entry(-v, row);
        };
        // deferred hits of the previous tile
        float pv[kHitRegs];
        uint32_t pr[kHitRegs];
        int pcnt = 0, pq = 0;
        uint32_t pslot = 0;
        auto flush_prev = [&]() {
            if (pcnt) {
#pragma unroll
                for (int i = 0; i < kHitRegs; ++i)
                    if (i < pcnt && pslot + i < (uint32_t)a.pool_cap) a.pool[(size_t)pq * a.pool_cap + pslot + i] = make_entry(-pv[i], pr[i]);
            }
            pcnt = 0;
        };
        uint32_t ge = 0;
        uint32_t tk_next = (cluster_id < n_rt) ? load_tau(0) : 0u;
        for (int rt = cluster_id; rt < n_rt; rt += n_clusters) {
            const long long tile_row0 = a.row_begin + (long long)rt * N_TILE;
            const int valid_cols = (int)min((long long)N_TILE, a.row_end - tile_row0);
            for (int qt = 0; qt < a.n_qt; ++qt, ++ge) {
                const uint32_t s = ge & 1;
                const int q = qt * N_TILE + q_lane;
                const bool q_ok = q < a.b;
                const uint32_t tk = tk_next;
                tk_next = load_tau(qt + 1 < a.n_qt ? qt + 1 : 0);   // in flight while this tile is scanned
                // scores >= thr survive; +inf rejects everything (padding query), -inf accepts everything (tau = +inf)
                const float thr = (!q_ok || a.debug_nohit) ? INFINITY : (tk >= 0xFF800000u ? -INFINITY : -float_from_order_key(tk));
                mbar_wait(&t_full[s], (ge >> 1) & 1);
                tc_fence_after();
                const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + s * N_TILE + half * COLS_PER_WARP;
                float cv[kHitRegs];
                uint32_t cr[kHitRegs];
                int ccnt = 0;
                float seed_m = -INFINITY;
#pragma unroll 1
                for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 32) {
                    float v[32];
                    tmem_ld32(t_lane + c0, v);
                    tmem_ld_wait();
                    const int col0 = half * COLS_PER_WARP + c0;
                    if constexpr (SEED) {
                        // threshold seeding: the best score of each group of COLS_PER_WARP rows (this warp's columns of the
                        // tile) goes to a fixed pool slot.  The k-th best of these group maxima bounds the k-th best row from
                        // below (k distinct groups => k rows).  One 8-byte store per thread and tile: the stores of a warp go
                        // to 32 different queries' pools (32 sectors), and with 32-row groups -- four stores per tile --
                        // they slowed the seed launch to 1050 TFLOP/s.
                        const int left = valid_cols - col0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) seed_m = fmaxf(seed_m, j < left ? v[j] : -INFINITY);
                        continue;
                    }
                    // four independent max chains (one per 8 columns) instead of one 32-long dependent chain: with two
                    // epilogue warps per scheduler the chain latency is exposed.  fmaxf drops NaN, like `>=` does.
                    float gm[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        gm[g] = fmaxf(v[8 * g], v[8 * g + 1]);
#pragma unroll
                        for (int j = 2; j < 8; ++j) gm[g] = fmaxf(gm[g], v[8 * g + j]);
                    }
                    const bool any = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3])) >= thr;
                    // Hit path.  It has to stay SMALL: the kernel must fit the 32 KB instruction cache (a 32x unrolled handler --
                    // tried twice, 42-45 KB of code -- costs 5-40 % because it stalls the MMA warp).
                    // Fast case, no TMEM re-read and nothing warp-wide: an 8-column group with exactly ONE hit -- its value is
                    // the group maximum already in a register, its column the set bit of the group's compare mask.  Groups with
                    // two or more hits, partial tiles and threads whose hit registers are full take the slow case: OR the
                    // masks across the warp and re-read every such column from TMEM (warp-uniform address).  (A 32-way select on
                    // v[] -- switch, or a 5-level select tree -- measured 1-2 % slower on the whole batch.)
                    if (__any_sync(0xFFFFFFFFu, any) && col0 < valid_cols) {
                        const int left = valid_cols - col0;
                        uint32_t mask = 0;            // columns of this lane left to the slow case
#pragma unroll
                        for (int g = 0; g < 4; ++g)
                            if (gm[g] >= thr) {   // only the 8-column groups that hold a hit are expanded
                                uint32_t m8 = 0;
#pragma unroll
                                for (int j = 0; j < 8; ++j) m8 |= (v[8 * g + j] >= thr) ? (1u << j) : 0u;
                                if ((m8 & (m8 - 1u)) == 0u && ccnt < kHitRegs && left >= 32) {
                                    const uint32_t row = (uint32_t)(tile_row0 + col0 + 8 * g + (__ffs(m8) - 1));
#pragma unroll
                                    for (int i = 0; i < kHitRegs; ++i)
                                        if (i == ccnt) {
                                            cv[i] = gm[g];
                                            cr[i] = row;
                                        }
                                    ++ccnt;
                                } else {
                                    mask |= m8 << (8 * g);
                                }
                            }
                        if (left < 32) mask &= (1u << left) - 1u;
                        if (__any_sync(0xFFFFFFFFu, mask != 0u)) {
                            uint32_t um = __reduce_or_sync(0xFFFFFFFFu, mask);
                            while (um) {
                                const int j = __ffs(um) - 1;
                                um &= um - 1;
                                float x;
                                tmem_ld1(t_lane + c0 + j, &x);
                                tmem_ld_wait();
                                if ((mask >> j) & 1u) {
                                    const uint32_t row = (uint32_t)(tile_row0 + col0 + j);
                                    if (ccnt < kHitRegs) {
#pragma unroll
                                        for (int i = 0; i < kHitRegs; ++i)
                                            if (i == ccnt) {
                                                cv[i] = x;
                                                cr[i] = row;
                                            }
                                        ++ccnt;
                                    } else {
                                        // registers full: straight to the pool.  Rare on purpose -- the store waits for the
                                        // atomic's round trip with the accumulator still held (measured: parking a hit that
                                        // could have been parked costs 1.5 % of a 1.25M-row batch)
                                        emit_now(q, x, row);
                                    }
                                }
                            }
                        }
                    }
                }
                if constexpr (SEED) {
                    if (q_ok && half * COLS_PER_WARP < valid_cols) {
                        const uint32_t slot = (uint32_t)((tile_row0 - a.row_begin) / COLS_PER_WARP) + (uint32_t)half;
                        if (slot < (uint32_t)a.pool_cap) a.pool[(size_t)q * a.pool_cap + slot] = make_entry(-seed_m, 0xFFFFFFFFu);
                    }
                }
                // The next tile's threshold was requested at the top of this tile; pin its arrival HERE, before the slot
                // reservation below: ptxas puts the load and the atomic on one scoreboard, and a wait placed after the
                // atomic (13 % of the epilogue's stall samples in a hit-heavy chunk) pays the atomic's whole round trip.
                asm volatile("" : "+r"(tk_next));
                // accumulator fully read: hand it back to the MMA warp before touching global memory
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 1) mbar_arrive(&t_empty[s]);
                    else asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(&t_empty[s]) & kPeerMask) : "memory");
                }
                flush_prev();                          // previous tile's slots were reserved one tile ago
                if (ccnt) {
                    pslot = atomicAdd(&a.qstate[q].pool_count, (uint32_t)ccnt);
                    pq = q;
                    pcnt = ccnt;
#pragma unroll
                    for (int i = 0; i < kHitRegs; ++i) {
                        pv[i] = cv[i];
                        pr[i] = cr[i];
                    }
                }
            }
        }
        flush_prev();
    }

    // ---- teardown ----
    __syncwarp();   // single-lane roles: reconverge the warp before the .aligned barriers below
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all();
    else __syncthreads();
    if (warp == 2) {
        if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side: tensor maps through the driver entry point (no link-time dependency on libcuda)

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

bool make_map(CUtensorMap* m, const void* gptr, long long n_rows, int kp, int pitch_elems, int box_cols, CUtensorMapSwizzle sw) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)n_rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)kTileRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace

namespace {

struct GemmPlan {
    int nseg, na, nb, a_resident;
    uint32_t a_slot_stride;
    size_t smem;
};

// Shared-memory plan.  Preference: whole-K query stages with double-buffered row tiles (the headline shape); when the
// operand tiles get too large for that (dim > ~208), the query tile is streamed in K segments through a small ring while
// one full-K row tile stays resident.  A batch that is a single query tile keeps all its segments resident and spends the
// rest on a deeper row pipeline.
bool plan_gemm(int kp, int n_qt, size_t smem_optin, GemmPlan* out) {
    const int kp_mma = (kp + 15) / 16 * 16;
    const int nb64 = kp_mma / 64, rem = kp_mma % 64;
    const size_t tail = (size_t)kTileRows * 2 * ((rem >= 32 ? 32 : 0) + ((rem % 32) >= 16 ? 16 : 0));
    const size_t tile = (((size_t)kTileRows * kp_mma * 2) + 1023) & ~(size_t)1023;
    const size_t fixed = 26 * 8 + 16 + 1024;
    auto slot_bytes = [&](int nseg) {
        const int base = nb64 / nseg, r = nb64 % nseg;
        size_t biggest = (size_t)(base + (r ? 1 : 0)) * 16384;
        size_t last = (size_t)base * 16384 + tail;           // the last segment carries the tail and never the remainder
        return ((biggest > last ? biggest : last) + 1023) & ~(size_t)1023;
    };
    const int segs[] = {1, 2, 4, 8};
    for (int nseg : segs) {
        if (nseg > 1 && nb64 < nseg) break;
        const size_t slot = slot_bytes(nseg);
        if (n_qt == 1) {   // resident query tile
            for (int nb = 3; nb >= 1; --nb) {
                const size_t need = (size_t)nseg * slot + (size_t)nb * tile + fixed;
                if (need <= smem_optin) {
                    *out = GemmPlan{nseg, nseg, nb, 1, (uint32_t)slot, need};
                    return true;
                }
            }
        }
        // (three query stages + ONE row buffer instead of 2 + 2 was measured: 5-8 % slower at 1.25M and 10M rows, round 2)
        {   // streamed query segments (also the fallback for one query tile that is too wide to stay resident)
            const int na_try[] = {nseg == 1 ? 2 : 3, 2};
            for (int nb = (nseg == 1 ? 2 : 1); nb >= 1; --nb)
                for (int na : na_try) {
                    const size_t need = (size_t)na * slot + (size_t)nb * tile + fixed;
                    if (need <= smem_optin) {
                        *out = GemmPlan{nseg, na, nb, 0, (uint32_t)slot, need};
                        return true;
                    }
                }
        }
    }
    return false;
}

}  // namespace

// rows per seed group (one pool entry each) of a seed-mode launch: the columns one epilogue warp owns (seed launches run
// with 8 epilogue warps)
int gemm_seed_group_rows(int cta_group) { return kTileRows * (cta_group == 2 ? 2 : 1) / (8 / 4); }

// 1 when some shared-memory plan exists for this operand width (any batch size), else 0: the caller must use the scan.
int gemm_row_stages(int kp, size_t smem_optin) {
    GemmPlan p;
    return (plan_gemm(kp, 2, smem_optin, &p) && plan_gemm(kp, 1, smem_optin, &p)) ? 1 : 0;
}

cudaError_t launch_gemm_filter(const GemmLaunch& g, cudaStream_t stream) {
    GemmTmaps tm;
    memset(&tm, 0, sizeof(tm));
    const int kp = g.kp;
    bool ok = make_map(&tm.q64, g.q_shadow, g.b_pad, kp, g.qkp, 64, CU_TENSOR_MAP_SWIZZLE_128B) &&
              make_map(&tm.q32, g.q_shadow, g.b_pad, kp, g.qkp, 32, CU_TENSOR_MAP_SWIZZLE_64B) &&
              make_map(&tm.q16, g.q_shadow, g.b_pad, kp, g.qkp, 16, CU_TENSOR_MAP_SWIZZLE_32B) &&
              make_map(&tm.r64, g.shadow, g.n_rows_total, kp, kp, 64, CU_TENSOR_MAP_SWIZZLE_128B) &&
              make_map(&tm.r32, g.shadow, g.n_rows_total, kp, kp, 32, CU_TENSOR_MAP_SWIZZLE_64B) &&
              make_map(&tm.r16, g.shadow, g.n_rows_total, kp, kp, 16, CU_TENSOR_MAP_SWIZZLE_32B);
    if (!ok) return cudaErrorInvalidValue;
    const int cg = g.cta_group;
    const int n_tile = kTileRows * cg;
    GemmArgs a{};
    a.row_begin = g.row_begin;
    a.row_end = g.row_end;
    a.b = g.b;
    a.n_qt = (g.b + n_tile - 1) / n_tile;
    a.kp_mma = (kp + 15) / 16 * 16;
    a.qstate = g.qstate;
    a.pool = g.pool;
    a.pool_cap = g.pool_cap;
    a.debug_nohit = getenv("B200ANN_NOHIT") ? 1 : 0;
    const long long rows = g.row_end - g.row_begin;
    const int n_rt = (int)((rows + n_tile - 1) / n_tile);
    if (n_rt <= 0 || a.n_qt <= 0) return cudaSuccess;
    int clusters = g.sm_count / cg;
    if (clusters > n_rt) clusters = n_rt;
    GemmPlan plan;
    size_t optin = 232448;
    {
        cudaDeviceProp prop;
        int dev = 0;
        cudaGetDevice(&dev);
        static size_t optin_cache[16] = {0};
        if (dev < 16 && !optin_cache[dev]) {
            cudaGetDeviceProperties(&prop, dev);
            optin_cache[dev] = prop.sharedMemPerBlockOptin;
        }
        if (dev < 16) optin = optin_cache[dev];
        if (!plan_gemm(kp, a.n_qt, optin, &plan)) return cudaErrorInvalidValue;
    }
    a.nseg = plan.nseg;
    a.na_stages = plan.na;
    a.nb_stages = plan.nb;
    a.a_resident = plan.a_resident;
    a.a_slot_stride = plan.a_slot_stride;
    // the threshold table rides behind the barriers when the plan leaves room for it (it does for every dim <= 208)
    a.tau_smem = (!g.seed_mode && a.n_qt <= kTauTiles && plan.smem + kTauBytes <= optin) ? 1 : 0;
    // two MMA issuers need the tile counter's parity to name both the accumulator stage and the query slot: whole-K query
    // stages in a ring of two, more than one query tile per row tile
    a.mma_warps = (g.mma_warps == 2 && plan.nseg == 1 && !plan.a_resident && plan.na == 2) ? 2 : 1;
    const size_t smem = plan.smem + (a.tau_smem ? kTauBytes : 0);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(clusters * cg);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cg;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e;
    const bool seg = !(plan.nseg == 1 && (plan.a_resident || plan.na == 2));
    const int ew = g.epi_warps == 16 ? 16 : 8;
    cfg.blockDim = dim3(gemm_threads(ew));
    void (*fn)(GemmTmaps, GemmArgs);
#define B200ANN_PICK(CGV, EWV)                                                                                     \
    (g.seed_mode ? (seg ? gemm_filter_kernel<CGV, true, true, EWV> : gemm_filter_kernel<CGV, true, false, EWV>)  \
                 : (seg ? gemm_filter_kernel<CGV, false, true, EWV> : gemm_filter_kernel<CGV, false, false, EWV>))
    if (cg == 1) fn = ew == 16 ? B200ANN_PICK(1, 16) : B200ANN_PICK(1, 8);
    else fn = ew == 16 ? B200ANN_PICK(2, 16) : B200ANN_PICK(2, 8);
#undef B200ANN_PICK
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaLaunchKernelEx(&cfg, fn, tm, a);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace b200ann
