// exact_fallback.cu -- tie-robust exact selection for the queries the bounded candidate selector cannot answer.
//
// The fast paths (scan_kernel.cu, gemm_filter.cu) keep a bounded superset of the top-k; when more rows than that
// tie inside the error margin of rank k (identical rows, distances that collapse to one fp32 value, NaN queries, a
// zero-norm Cosine query ...) they flag the query instead of guessing.  This path answers such a query exactly, for
// any input, with the same contract: ascending (Float.compare(distance), id), i.e. what
// BruteForceIndex.queryWithDistance (BruteForceIndex.scala:66-91) returns up to its heap-dependent tie order.
//
//   1. exact_all      : one thread per row, the oracle's fp64-accumulate / round-once distance -> order key per row
//   2. radix select   : 12 MSB-first passes of 8 bits over the 96-bit composite key (distance key : id), each a
//                       256-bin histogram over the rows still matching the prefix + a one-thread pick of the bin
//                       holding rank k.  After the last pass the prefix IS the k-th smallest composite key.
//   3. collect + sort : rows with key < K* (and as many == K* as needed) -> one CTA bitonic sort -> outputs.
// HBM traffic per query: n*dim*4 (distances) + 12 * n * 12 bytes (select); only degenerate inputs ever pay it.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace b200ann {

namespace {

struct SelectState {
    uint32_t prefix_d;              // decided high bytes of the distance key
    unsigned long long prefix_id;   // decided high bytes of the (sign-flipped) id
    uint32_t k_rem;                 // rank still to locate inside the current prefix class
    uint32_t n_less, n_equal;       // collect counters
    uint32_t hist[256];
};

__device__ __forceinline__ unsigned long long id_key(long long id) { return (unsigned long long)id ^ 0x8000000000000000ull; }

template <bool ACC32>
__global__ void __launch_bounds__(256) exact_all_kernel(const float* __restrict__ rows, long long n, int pitch, int dim, int metric,
                                                        int l2_squared, const float* __restrict__ query, uint32_t* __restrict__ dkey) {
    extern __shared__ double q64[];
    float* q32 = reinterpret_cast<float*>(q64);
    __shared__ double nb_s;
    __shared__ float nbf_s;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) {
        if (ACC32) q32[i] = query[i];
        else q64[i] = (double)query[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (ACC32) nbf_s = (metric == kMetricCosine) ? exact_query_norm2_f32(q32, dim) : 0.f;
        else nb_s = (metric == kMetricCosine) ? exact_query_norm2(q64, dim) : 0.0;
    }
    __syncthreads();
    const ExactQuery eq{q64, ACC32 ? 0.0 : nb_s};
    const float nbf = ACC32 ? nbf_s : 0.f;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const float* rp = rows + (size_t)r * pitch;
        dkey[r] = float_order_key(ACC32 ? exact_distance_rows_f32(metric, rp, q32, nbf, dim, l2_squared)
                                        : exact_distance_rows(metric, rp, eq, dim, l2_squared));
    }
}

__global__ void init_state_kernel(SelectState* st, uint32_t k) {
    if (threadIdx.x == 0) {
        st->prefix_d = 0;
        st->prefix_id = 0;
        st->k_rem = k;
        st->n_less = 0;
        st->n_equal = 0;
    }
    st->hist[threadIdx.x] = 0;
}

// pass 0..3 refine the distance key (byte 3-pass of dkey), pass 4..11 refine the id among rows with dkey == prefix_d
__global__ void __launch_bounds__(256) radix_hist_kernel(const uint32_t* __restrict__ dkey, const int64_t* __restrict__ ids, long long n,
                                                         SelectState* st, int pass) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t pd = st->prefix_d;
    const unsigned long long pi = st->prefix_id;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const uint32_t dk = dkey[r];
        if (pass < 4) {
            const int shift = 24 - 8 * pass;
            const bool match = pass == 0 || (dk >> (shift + 8)) == (pd >> (shift + 8));
            if (match) atomicAdd(&h[(dk >> shift) & 255u], 1u);
        } else if (dk == pd) {
            const unsigned long long ik = id_key(ids[r]);
            const int shift = 56 - 8 * (pass - 4);
            const bool match = pass == 4 || (ik >> (shift + 8)) == (pi >> (shift + 8));
            if (match) atomicAdd(&h[(ik >> shift) & 255u], 1u);
        }
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void radix_pick_kernel(SelectState* st, int pass) {
    if (threadIdx.x == 0) {
        uint32_t k = st->k_rem, cum = 0;
        int bin = 255;
        for (int i = 0; i < 256; ++i) {
            uint32_t c = st->hist[i];
            if (cum + c >= k) {
                bin = i;
                break;
            }
            cum += c;
        }
        st->k_rem = k - cum;
        if (pass < 4) st->prefix_d |= (uint32_t)bin << (24 - 8 * pass);
        else st->prefix_id |= (unsigned long long)bin << (56 - 8 * (pass - 4));
    }
    __syncthreads();
    st->hist[threadIdx.x] = 0;
}

// rows strictly below K* fill the front of the list; rows equal to K* (duplicate (distance,id) pairs) fill up to k
__global__ void __launch_bounds__(256) collect_kernel(const uint32_t* __restrict__ dkey, const int64_t* __restrict__ ids, long long n,
                                                      SelectState* st, uint32_t k, uint32_t* out_key, long long* out_id) {
    const uint32_t pd = st->prefix_d;
    const unsigned long long pi = st->prefix_id;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const uint32_t dk = dkey[r];
        if (dk > pd) continue;
        const long long id = ids[r];
        const unsigned long long ik = id_key(id);
        if (dk < pd || ik < pi) {
            uint32_t slot = atomicAdd(&st->n_less, 1u);
            if (slot < k) {
                out_key[slot] = dk;
                out_id[slot] = id;
            }
        } else if (ik == pi) {
            uint32_t e = atomicAdd(&st->n_equal, 1u);
            // equal keys go to the back: slot k-1-e; the strict ones never reach those slots (n_less <= k - #equal needed)
            if (e < k) {
                out_key[k + e] = dk;      // staged in the second half, merged by the sort kernel
                out_id[k + e] = id;
            }
        }
    }
}

__global__ void __launch_bounds__(512) fallback_sort_kernel(SelectState* st, uint32_t k, const uint32_t* in_key, const long long* in_id,
                                                            int64_t* out_ids, float* out_dist, int32_t* out_count, int k_out) {
    extern __shared__ __align__(16) unsigned char sm[];
    int n2 = 2;
    while (n2 < (int)k) n2 <<= 1;
    long long* cid = reinterpret_cast<long long*>(sm);
    uint32_t* ckey = reinterpret_cast<uint32_t*>(sm + (size_t)n2 * 8);
    const uint32_t n_less = min(st->n_less, k);
    const uint32_t n_eq = min(st->n_equal, k - n_less);
    const uint32_t total = n_less + n_eq;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        if ((uint32_t)i < n_less) {
            ckey[i] = in_key[i];
            cid[i] = in_id[i];
        } else if ((uint32_t)i < total) {
            ckey[i] = in_key[k + (i - n_less)];
            cid[i] = in_id[k + (i - n_less)];
        } else {
            ckey[i] = 0xFFFFFFFFu;
            cid[i] = 0x7FFFFFFFFFFFFFFFll;
        }
    }
    __syncthreads();
    for (int kk = 2; kk <= n2; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    uint32_t ka = ckey[i], kb = ckey[ixj];
                    long long ia = cid[i], ib = cid[ixj];
                    bool up = ((i & kk) == 0);
                    bool gt = ka > kb || (ka == kb && ia > ib);
                    if (gt == up) {
                        ckey[i] = kb;
                        ckey[ixj] = ka;
                        cid[i] = ib;
                        cid[ixj] = ia;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int j = threadIdx.x; j < k_out; j += blockDim.x) {
        if ((uint32_t)j < total) {
            out_ids[j] = cid[j];
            out_dist[j] = float_from_order_key(ckey[j]);
        } else {
            out_ids[j] = -1;
            out_dist[j] = INFINITY;
        }
    }
    if (threadIdx.x == 0 && out_count) *out_count = (int32_t)total;
}

}  // namespace

size_t fallback_scratch_bytes(long long n, int k) {
    return sizeof(SelectState) + 256 + (size_t)n * 4 + (size_t)2 * k * (4 + 8) + 64;
}

// One query, exactly.  `scratch` must hold fallback_scratch_bytes(n, k).  Launches 29 kernels on `stream`.
cudaError_t launch_exact_fallback(const FallbackParams& p, cudaStream_t stream, long long* launches) {
    unsigned char* base = reinterpret_cast<unsigned char*>(p.scratch);
    SelectState* st = reinterpret_cast<SelectState*>(base);
    size_t off = (sizeof(SelectState) + 255) & ~(size_t)255;
    uint32_t* dkey = reinterpret_cast<uint32_t*>(base + off);
    off += ((size_t)p.n_rows * 4 + 15) & ~(size_t)15;
    long long* lid = reinterpret_cast<long long*>(base + off);
    off += (size_t)2 * p.k * 8;
    uint32_t* lkey = reinterpret_cast<uint32_t*>(base + off);
    const int grid = (int)std::min<long long>((p.n_rows + 255) / 256, 148 * 8);
    if (p.accum_f32)
        exact_all_kernel<true><<<grid, 256, p.dim * sizeof(double), stream>>>(p.rows, p.n_rows, p.pitch, p.dim, p.metric, p.l2_squared, p.query, dkey);
    else
        exact_all_kernel<false><<<grid, 256, p.dim * sizeof(double), stream>>>(p.rows, p.n_rows, p.pitch, p.dim, p.metric, p.l2_squared, p.query, dkey);
    init_state_kernel<<<1, 256, 0, stream>>>(st, (uint32_t)p.k);
    for (int pass = 0; pass < 12; ++pass) {
        radix_hist_kernel<<<grid, 256, 0, stream>>>(dkey, p.ids, p.n_rows, st, pass);
        radix_pick_kernel<<<1, 256, 0, stream>>>(st, pass);
    }
    collect_kernel<<<grid, 256, 0, stream>>>(dkey, p.ids, p.n_rows, st, (uint32_t)p.k, lkey, lid);
    int n2 = 2;
    while (n2 < p.k) n2 <<= 1;
    {   // k in (4096, 16384] needs 96..192 KB of dynamic shared memory: opt in like every other large-smem kernel
        cudaError_t e = cudaFuncSetAttribute(fallback_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)n2 * 12));
        if (e != cudaSuccess) return e;
    }
    fallback_sort_kernel<<<1, 512, (size_t)n2 * 12, stream>>>(st, (uint32_t)p.k, lkey, lid, p.out_ids, p.out_dist, p.out_count, p.k_out);
    if (launches) *launches += 28;
    return cudaGetLastError();
}

}  // namespace b200ann
