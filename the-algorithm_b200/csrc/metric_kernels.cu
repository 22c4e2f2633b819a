// metric_kernels.cu -- the Metric trait itself, evaluated on the device for plain vector pairs.
//
//   ann_distance_pairs  : Metric.distance(embedding1, embedding2) for n independent pairs
//                         (ann/src/main/scala/com/twitter/ann/common/Metric.scala:76-86; L2 :89-94, Cosine :120-125,
//                         InnerProduct :153-158 through MetricUtil :263-283)
//   ann_normalize_rows  : MetricUtil.norm (Metric.scala:285-289), the L2 normalisation HNSW / Faiss apply before they treat
//                         Cosine as InnerProduct (DistanceFunctionGenerator.scala:11-15; Hnsw.scala:149-155)
//
// Same arithmetic as the index's exact rescoring (common.cuh, exact_distance_rows) and as the oracle, operation for
// operation: fp64 accumulation in index order with individually rounded operations, one rounding to fp32.  Both calls are
// HBM bound (read 2*n*dim*4 resp. n*dim*4 bytes once): one warp owns one pair / row, its lanes load the vectors with
// coalesced accesses into shared memory, and lane 0 runs the sequential chain the parity contract prescribes.
#include <algorithm>
#include <cstdio>

#include "../../include/b200ann.h"
#include "common.cuh"
#include "kernels.h"

namespace b200ann {

namespace {

constexpr int kWarpsPerCta = 8;

// distance(a, b) with the index's conventions; a and b in shared memory
__device__ float pair_distance(int metric, const float* a, const float* b, int d, int l2_squared, int accum_f32) {
    if (accum_f32) {   // ANN_FLAG_ACCUM_F32: the oracle's distance_f32, operation for operation (b is a plain fp32 vector here)
        const float nb = metric == kMetricCosine ? exact_query_norm2_f32(b, d) : 0.f;
        // a lives in shared memory as dim floats; the float4 walk of the index kernels wants 16-byte alignment, so go scalar
        float s0 = 0.f, s1 = 0.f;
        for (int i = 0; i < d; ++i) {
            if (metric == kMetricL2) {
                const float df = __fsub_rn(a[i], b[i]);
                s0 = __fadd_rn(s0, __fmul_rn(df, df));
            } else {
                s0 = __fadd_rn(s0, __fmul_rn(a[i], b[i]));
                if (metric == kMetricCosine) s1 = __fadd_rn(s1, __fmul_rn(a[i], a[i]));
            }
        }
        if (metric == kMetricL2) return l2_squared ? s0 : __fsqrt_rn(s0);
        if (metric == kMetricIP) return __fsub_rn(1.0f, s0);
        return __fsub_rn(1.0f, __fdiv_rn(s0, __fmul_rn(__fsqrt_rn(s1), __fsqrt_rn(nb))));
    }
    if (metric == kMetricL2) {
        double acc = 0.0;
        for (int i = 0; i < d; ++i) {
            const double df = __dsub_rn((double)a[i], (double)b[i]);
            acc = __dadd_rn(acc, __dmul_rn(df, df));
        }
        return __double2float_rn(l2_squared ? acc : __dsqrt_rn(acc));
    }
    double dot = 0.0;
    for (int i = 0; i < d; ++i) dot = __fma_rn((double)a[i], (double)b[i], dot);   // float products are exact in fp64
    if (metric == kMetricIP) return __fsub_rn(1.0f, __double2float_rn(dot));
    double na = 0.0, nb = 0.0;
    for (int i = 0; i < d; ++i) na = __fma_rn((double)a[i], (double)a[i], na);
    for (int i = 0; i < d; ++i) nb = __fma_rn((double)b[i], (double)b[i], nb);
    const double cs = __ddiv_rn(dot, __dmul_rn(__dsqrt_rn(na), __dsqrt_rn(nb)));
    return __fsub_rn(1.0f, __double2float_rn(cs));
}

__global__ void __launch_bounds__(32 * kWarpsPerCta) distance_pairs_kernel(int metric, int l2_squared, int accum_f32, int dim,
                                                                           const float* a, const float* b, long long n, float* out) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sa = sm + (size_t)warp * 2 * dim;
    float* sb = sa + dim;
    for (long long p = (long long)blockIdx.x * kWarpsPerCta + warp; p < n; p += (long long)gridDim.x * kWarpsPerCta) {
        for (int i = lane; i < dim; i += 32) {
            sa[i] = a[(size_t)p * dim + i];
            sb[i] = b[(size_t)p * dim + i];
        }
        __syncwarp();
        if (lane == 0) out[p] = pair_distance(metric, sa, sb, dim, l2_squared, accum_f32);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32 * kWarpsPerCta) normalize_rows_kernel(int dim, const float* rows, long long n, float* out) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sv = sm + (size_t)warp * dim;
    for (long long r = (long long)blockIdx.x * kWarpsPerCta + warp; r < n; r += (long long)gridDim.x * kWarpsPerCta) {
        for (int i = lane; i < dim; i += 32) sv[i] = rows[(size_t)r * dim + i];
        __syncwarp();
        double n2 = 0.0;
        if (lane == 0)
            for (int i = 0; i < dim; ++i) n2 = __fma_rn((double)sv[i], (double)sv[i], n2);
        n2 = __shfl_sync(0xFFFFFFFFu, n2, 0);
        const double nrm = __dsqrt_rn(n2);
        for (int i = lane; i < dim; i += 32) out[(size_t)r * dim + i] = __double2float_rn(__ddiv_rn((double)sv[i], nrm));
        __syncwarp();
    }
}

// In place, rows of an index matrix (pitch >= dim): what ANN_FLAG_COSINE_UNIT_ROWS applies to every appended row.
__global__ void __launch_bounds__(32 * kWarpsPerCta) normalize_inplace_kernel(float* rows, long long row0, long long n, int dim, int pitch) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sv = sm + (size_t)warp * dim;
    for (long long r = (long long)blockIdx.x * kWarpsPerCta + warp; r < n; r += (long long)gridDim.x * kWarpsPerCta) {
        float* a = rows + (size_t)(row0 + r) * pitch;
        for (int i = lane; i < dim; i += 32) sv[i] = a[i];
        __syncwarp();
        double n2 = 0.0;
        if (lane == 0)
            for (int i = 0; i < dim; ++i) n2 = __fma_rn((double)sv[i], (double)sv[i], n2);
        n2 = __shfl_sync(0xFFFFFFFFu, n2, 0);
        const double nrm = __dsqrt_rn(n2);
        for (int i = lane; i < dim; i += 32) a[i] = __double2float_rn(__ddiv_rn((double)sv[i], nrm));
        __syncwarp();
    }
}

struct Staging {
    float* p[3] = {nullptr, nullptr, nullptr};
    ~Staging() {
        for (float* x : p)
            if (x) cudaFree(x);
    }
};

#define METRIC_TRY(expr)                                                                             \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            char _b[384];                                                                            \
            snprintf(_b, sizeof(_b), "%s failed: %s", #expr, cudaGetErrorString(_e));                \
            (void)cudaGetLastError();                                                                \
            return report_error(_e == cudaErrorMemoryAllocation ? ANN_ERR_OUT_OF_MEMORY : ANN_ERR_CUDA, _b); \
        }                                                                                            \
    } while (0)

int pick_device(int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        (void)cudaGetLastError();
        return report_error(ANN_ERR_NO_DEVICE, "no CUDA device: the engine has no CPU fallback");
    }
    if (device < 0 || device >= count) return report_error(ANN_ERR_INVALID_ARGUMENT, "device index out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        (void)cudaGetLastError();
        return report_error(ANN_ERR_NO_DEVICE, "device is not sm_100: the library is built for B200 only");
    }
    METRIC_TRY(cudaSetDevice(device));
    return ANN_OK;
}

int grid_for(long long n) { return (int)std::min<long long>((n + kWarpsPerCta - 1) / kWarpsPerCta, 148LL * 8); }

}  // namespace

// MetricUtil.norm on device-resident rows (convention C8), in place (out == nullptr: rows [row0, row0+n) of a matrix with
// `pitch`) or into a dense [n][dim] copy (`out`, source pitch = dim).
cudaError_t launch_normalize_rows_device(float* rows, long long row0, long long n, int dim, int pitch, float* out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const size_t smem = (size_t)kWarpsPerCta * dim * sizeof(float);
    if (out) normalize_rows_kernel<<<grid_for(n), 32 * kWarpsPerCta, smem, stream>>>(dim, rows + (size_t)row0 * dim, n, out);
    else normalize_inplace_kernel<<<grid_for(n), 32 * kWarpsPerCta, smem, stream>>>(rows, row0, n, dim, pitch);
    return cudaGetLastError();
}

}  // namespace b200ann

using namespace b200ann;

extern "C" {

int ann_distance_pairs(int32_t metric, uint32_t flags, int32_t dim, const float* a, const float* b, int64_t n, float* out,
                       int32_t device) {
    if (metric < 0 || metric > 2) return report_error(ANN_ERR_INVALID_ARGUMENT, "ann_distance_pairs: metric must be 0 (L2), 1 (Cosine) or 2 (InnerProduct)");
    if (dim < 1 || dim > 1024) return report_error(ANN_ERR_INVALID_ARGUMENT, "ann_distance_pairs: dim must be in [1, 1024]");
    if (n < 0) return report_error(ANN_ERR_INVALID_ARGUMENT, "ann_distance_pairs: n < 0");
    if (n == 0) return ANN_OK;
    if (!a || !b || !out) return report_error(ANN_ERR_NULL_POINTER, "ann_distance_pairs: NULL buffer");
    int rc = pick_device(device);
    if (rc) return rc;
    Staging st;
    const size_t bytes = (size_t)n * dim * sizeof(float);
    METRIC_TRY(cudaMalloc(&st.p[0], bytes));
    METRIC_TRY(cudaMalloc(&st.p[1], bytes));
    METRIC_TRY(cudaMalloc(&st.p[2], (size_t)n * sizeof(float)));
    METRIC_TRY(cudaMemcpy(st.p[0], a, bytes, cudaMemcpyHostToDevice));
    METRIC_TRY(cudaMemcpy(st.p[1], b, bytes, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)kWarpsPerCta * 2 * dim * sizeof(float);
    METRIC_TRY(cudaFuncSetAttribute(distance_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    distance_pairs_kernel<<<grid_for(n), 32 * kWarpsPerCta, smem>>>(metric, (flags & ANN_FLAG_L2_SQUARED) ? 1 : 0, (flags & ANN_FLAG_ACCUM_F32) ? 1 : 0, dim, st.p[0], st.p[1],
                                                                   (long long)n, st.p[2]);
    METRIC_TRY(cudaGetLastError());
    METRIC_TRY(cudaMemcpy(out, st.p[2], (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    return ANN_OK;
}

int ann_normalize_rows(int32_t dim, const float* rows, int64_t n, float* out, int32_t device) {
    if (dim < 1 || dim > 1024) return report_error(ANN_ERR_INVALID_ARGUMENT, "ann_normalize_rows: dim must be in [1, 1024]");
    if (n < 0) return report_error(ANN_ERR_INVALID_ARGUMENT, "ann_normalize_rows: n < 0");
    if (n == 0) return ANN_OK;
    if (!rows || !out) return report_error(ANN_ERR_NULL_POINTER, "ann_normalize_rows: NULL buffer");
    int rc = pick_device(device);
    if (rc) return rc;
    Staging st;
    const size_t bytes = (size_t)n * dim * sizeof(float);
    METRIC_TRY(cudaMalloc(&st.p[0], bytes));
    METRIC_TRY(cudaMalloc(&st.p[1], bytes));
    METRIC_TRY(cudaMemcpy(st.p[0], rows, bytes, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)kWarpsPerCta * dim * sizeof(float);
    normalize_rows_kernel<<<grid_for(n), 32 * kWarpsPerCta, smem>>>(dim, st.p[0], (long long)n, st.p[1]);
    METRIC_TRY(cudaGetLastError());
    METRIC_TRY(cudaMemcpy(out, st.p[1], bytes, cudaMemcpyDeviceToHost));
    return ANN_OK;
}

}  // extern "C"
