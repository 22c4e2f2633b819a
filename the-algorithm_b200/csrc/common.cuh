// common.cuh -- shared device helpers for the exact-kNN kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace b200ann {

constexpr int kMetricL2 = 0, kMetricCosine = 1, kMetricIP = 2;

// Candidate entry: (orderable(g) << 32) | local_row, g = "badness" (smaller is nearer).
//   InnerProduct: g = -dot      Cosine: g = -dot/|a|      L2: g = sum (a-b)^2  (scan)  or  |a|^2/2 - a.b (gemm)
// Sorting entries as u64 ascending orders by g then row.
typedef unsigned long long entry_t;

// java.lang.Float.compare as an unsigned key (ann/.../common/Metric.scala:17-36):
// -0.0 < +0.0, +inf < NaN, all NaN equal.
__host__ __device__ __forceinline__ uint32_t float_order_key(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t b = c.u;
#endif
    if (f != f) return 0xFFFFFFFFu;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float float_from_order_key(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}

__device__ __forceinline__ entry_t make_entry(float g, uint32_t row) {
    return ((entry_t)float_order_key(g) << 32) | (entry_t)row;
}
__device__ __forceinline__ float entry_g(entry_t e) { return float_from_order_key((uint32_t)(e >> 32)); }
__device__ __forceinline__ uint32_t entry_row(entry_t e) { return (uint32_t)(e & 0xFFFFFFFFull); }

constexpr entry_t kEntryPad = 0xFFFFFFFFFFFFFFFFull;

// ---- per-query state shared by the scan, the GEMM filter, compaction and finalize -----------------
// tau_key: orderable key of the current acceptance threshold on g (accept g <= tau).  Starts at +inf.
struct QueryState {
    float eps_abs;   // absolute slack on g covering approximate-score error + fp32 tie granularity
    float eps_rel;   // relative slack on |g|
    float qnorm;     // |b|
    float qnorm2;    // |b|^2
    uint32_t tau_key;
    uint32_t pool_count;     // entries written to this query's pool (may exceed capacity => overflow)
    uint32_t special_count;  // rows whose approximate score was not finite (always rescored exactly)
    uint32_t flags;          // bit0: pool overflow, bit1: special overflow, bit2: survivor overflow
};

constexpr uint32_t kFlagPoolOverflow = 1u, kFlagSpecialOverflow = 2u, kFlagSurvivorOverflow = 4u;
constexpr int kSpecialCap = 256;  // per query
// Special rows (non-finite entries, zero norm under Cosine) get an all-zero shadow row whose augmented column is -kShadowBig:
// their tensor-core score is ~ -1e38, so they are never a threshold witness nor a candidate; they are rescored exactly from
// the index-wide special list instead.  Pool entries with g above kSpecialG are such rows seen before any threshold existed.
constexpr float kShadowBig = 1.0e38f;
constexpr float kSpecialG = 1.0e37f;

// threshold with margin: everything with g <= kth + margin(kth) may still belong to the exact top-k
__device__ __forceinline__ float widen(float kth, float eps_abs, float eps_rel) {
    return kth + (2.0f * eps_abs + 2.0f * eps_rel * fabsf(kth));
}

// ---- exact distance, operation for operation the oracle's distance_f64 (oracle/oracle.c): fp64 accumulation in index
// order, explicit round-to-nearest mul/add (no contraction), one rounding to fp32, `1 - x` in fp32.
// The query is pre-converted to fp64 once per CTA (`b64`, shared memory) together with its squared norm `nb` (the same
// sequential sum for every row).  Per block of 16 elements the independent work (convert, subtract / multiply) is done
// first and only the additions form the dependent chain; the row itself is read with 128-bit loads, double buffered.
struct ExactQuery {
    const double* b64;   // [d] query as doubles
    double nb;           // sum b_i^2 in index order (Cosine)
};

// sequential sum of squares of the query, exactly as the oracle accumulates it
__device__ __forceinline__ double exact_query_norm2(const double* b64, int d) {
    double nb = 0.0;
    for (int i = 0; i < d; ++i) nb = __fma_rn(b64[i], b64[i], nb);   // squares of floats are exact in fp64: same rounding as mul + add
    return nb;
}

// METRIC >= 0 fixes the metric at compile time: with a run-time metric the compiler keeps BOTH accumulators alive and selects
// per element (for InnerProduct: the Cosine norm's DFMA + two FSEL + two MOV on top of the one DFMA that is needed -- 11
// instructions per element instead of 4 in the finalize kernel's profile).
template <bool kGlobal = true, int METRIC = -1>
__device__ __forceinline__ float exact_distance_rows(int metric_rt, const float* __restrict__ a, const ExactQuery& q, int d, int l2_squared) {
    const int metric = METRIC >= 0 ? METRIC : metric_rt;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    auto ld = [&](int i) -> float4 { return kGlobal ? __ldg(a4 + i) : a4[i]; };
    const int n4 = (d + 3) >> 2;       // float4 holding at least one valid element
    const int full = d >> 4;           // blocks of 16 elements with no tail
    double s0 = 0.0, s1 = 0.0;         // L2: sum (a-b)^2 ; otherwise a.b and (Cosine) a.a
    float4 cur[4], nxt[4];
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    // Dot-product terms use ONE fused multiply-add: the product of two floats is exact in fp64, so fma(x, y, s) rounds
    // exactly like the oracle's mul-then-add.  The L2 term (a-b)^2 is NOT exact in fp64, so it keeps the oracle's
    // separate, individually rounded sub / mul / add.
    auto step = [&](float xf, double y) {
        const double x = (double)xf;
        if (metric == kMetricL2) {
            const double df = __dsub_rn(x, y);
            s0 = __dadd_rn(s0, __dmul_rn(df, df));
        } else {
            s0 = __fma_rn(x, y, s0);
            if (metric == kMetricCosine) s1 = __fma_rn(x, x, s1);
        }
    };
#pragma unroll
    for (int j = 0; j < 4; ++j) cur[j] = (j < n4) ? ld(j) : zero;
    for (int blk = 0; blk < full; ++blk) {
        const int base4 = blk << 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) nxt[j] = (base4 + 4 + j < n4) ? ld(base4 + 4 + j) : zero;
        const double* bq = q.b64 + (blk << 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            step(cur[j].x, bq[4 * j]);
            step(cur[j].y, bq[4 * j + 1]);
            step(cur[j].z, bq[4 * j + 2]);
            step(cur[j].w, bq[4 * j + 3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
    }
    // tail: fewer than 16 elements, all already in `cur`
    for (int i = full << 4; i < d; ++i) {
        const int j = (i >> 2) & 3, c = i & 3;
        const float xf = c == 0 ? cur[j].x : c == 1 ? cur[j].y : c == 2 ? cur[j].z : cur[j].w;
        step(xf, q.b64[i]);
    }
    if (metric == kMetricL2) return __double2float_rn(l2_squared ? s0 : __dsqrt_rn(s0));
    if (metric == kMetricIP) return __fsub_rn(1.0f, __double2float_rn(s0));
    const double cs = __ddiv_rn(s0, __dmul_rn(__dsqrt_rn(s1), __dsqrt_rn(q.nb)));
    return __fsub_rn(1.0f, __double2float_rn(cs));
}

// ---- the other accumulator convention (ANN_FLAG_ACCUM_F32): MetricUtil.dot is typed Float with no cast at its call site
// (Metric.scala:264-269), so the unshipped EmbeddingMath.Float.dotProduct most plausibly accumulates in fp32.  Operation
// for operation the oracle's distance_f32 (oracle/oracle.c, accum=1): sequential fp32 accumulation in index order with
// individually rounded mul / add (no contraction), correctly rounded sqrt / div, `1 - x` in fp32.  `qf` is the query in
// fp32 (shared memory), `nbf` its sequential fp32 squared norm.
__device__ __forceinline__ float exact_query_norm2_f32(const float* qf, int d) {
    float nb = 0.f;
    for (int i = 0; i < d; ++i) nb = __fadd_rn(nb, __fmul_rn(qf[i], qf[i]));
    return nb;
}

template <bool kGlobal = true, int METRIC = -1>
__device__ __forceinline__ float exact_distance_rows_f32(int metric_rt, const float* __restrict__ a, const float* qf, float nbf, int d,
                                                         int l2_squared) {
    const int metric = METRIC >= 0 ? METRIC : metric_rt;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    auto ld = [&](int i) -> float4 { return kGlobal ? __ldg(a4 + i) : a4[i]; };
    const int n4 = (d + 3) >> 2;
    float s0 = 0.f, s1 = 0.f;
    auto step = [&](float x, float y) {
        if (metric == kMetricL2) {
            const float df = __fsub_rn(x, y);
            s0 = __fadd_rn(s0, __fmul_rn(df, df));
        } else {
            s0 = __fadd_rn(s0, __fmul_rn(x, y));
            if (metric == kMetricCosine) s1 = __fadd_rn(s1, __fmul_rn(x, x));
        }
    };
    float4 cur = n4 > 0 ? ld(0) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < n4; ++j) {
        const float4 nxt = (j + 1 < n4) ? ld(j + 1) : cur;
        const int i = j << 2;
        step(cur.x, qf[i]);
        if (i + 1 < d) step(cur.y, qf[i + 1]);
        if (i + 2 < d) step(cur.z, qf[i + 2]);
        if (i + 3 < d) step(cur.w, qf[i + 3]);
        cur = nxt;
    }
    if (metric == kMetricL2) return l2_squared ? s0 : __fsqrt_rn(s0);
    if (metric == kMetricIP) return __fsub_rn(1.0f, s0);
    return __fsub_rn(1.0f, __fdiv_rn(s0, __fmul_rn(__fsqrt_rn(s1), __fsqrt_rn(nbf))));
}

// ---- mbarrier / bulk-copy PTX (Hopper+; the only async-copy path used by the scan) -----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace b200ann
