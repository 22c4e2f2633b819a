// append_kernels.cu -- K1 (norms / shadow rows at append time) and query preparation.
//
// The reference appends one JVM object per row and recomputes both norms for every (row, query) pair
// (BruteForceIndex.scala:48-52; Metric.scala:278-283).  Here a batch lands in the row-major device matrix with
// one copy, and one pass over the new rows produces everything the query kernels need later:
//   row_norm[r]   |a| rounded up            -> error bounds of the approximate passes
//   inv_norm[r]   1/|a|  (Cosine)           -> turns the scan's dot product into |b|*cos
//   shadow[r][:]  bf16 operand of the tensor-core filter (K-major, zero padded to kp):
//                 InnerProduct: a          Cosine: a/|a|          L2: [a, h1, h2, h3],  h1+h2+h3 = -|a|^2/2
//                 (the query operand carries [b, 1, 1, 1], so the MMA itself yields a.b - |a|^2/2)
// HBM bound: reads n*dim*4 bytes once, writes n*(kp*2 + 8) bytes.
#include "common.cuh"
#include "kernels.h"

namespace b200ann {

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

__global__ void __launch_bounds__(256) append_rows_kernel(AppendParams p) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float local_max = 0.f, local_resid = 0.f;
    unsigned long long local_special = 0;
    for (long long r = warp; r < p.n_new; r += nwarps) {
        const long long row = p.slots ? p.slots[r] : p.row0 + r;
        float* a = const_cast<float*>(p.rows) + (size_t)row * p.pitch;
        if (p.slots) {  // in-place update: move the staged embedding into the matrix first
            const float* src = p.staged + (size_t)r * p.dim;
            for (int i = lane; i < p.dim; i += 32) a[i] = src[i];
            __syncwarp();
        }
        double n2 = 0.0;
        bool finite = true;
        for (int i = lane; i < p.dim; i += 32) {
            float v = a[i];
            finite = finite && ((v - v) == 0.0f);
            n2 += (double)v * (double)v;
        }
        n2 = warp_sum(n2);
        finite = __all_sync(0xFFFFFFFFu, finite);
        const double nrm = sqrt(n2);
        const float nrm_up = __double2float_ru(nrm);
        const bool special = !finite || !((nrm_up - nrm_up) == 0.0f) || (p.metric == kMetricCosine && n2 == 0.0);
        const float inv = (float)(1.0 / nrm);
        if (lane == 0) {
            p.row_norm[row] = nrm_up;
            if (p.inv_norm) p.inv_norm[row] = inv;
            if (special) {
                local_special += 1;
                const unsigned long long slot = atomicAdd(p.n_special, 1ull);
                if (slot < (unsigned long long)kSpecialCap) p.special_list[slot] = (uint32_t)row;
            } else {
                local_max = fmaxf(local_max, nrm_up);
            }
        }
        if (p.shadow) {
            __nv_bfloat16* s = p.shadow + (size_t)row * p.kp;
            const float scale = (p.metric == kMetricCosine) ? inv : 1.0f;
            const int aug = p.dim;   // first augmented column: 0 for ordinary rows, -kShadowBig for special ones
            double r2 = 0.0;
            for (int i = lane; i < p.kp; i += 32) {
                float v = 0.f;
                if (i < p.dim && !special) v = a[i] * scale;
                if (i == aug && special) v = -kShadowBig;
                const __nv_bfloat16 h = __float2bfloat16_rn(v);
                s[i] = h;
                if (i < p.dim && !special) {
                    const double df = (double)v - (double)__bfloat162float(h);   // exact
                    r2 += df * df;
                }
            }
            // Measured rounding error of this shadow row, |shadow - operand| in the 2-norm, rounded up.  The filter's error
            // bound uses the maximum over the index instead of the worst case 2^-8 |a| (about 2.5x smaller on dense
            // data, and never an under-statement).  Under Cosine the operand is a/|a|, and the fp32 `a[i] * inv` itself
            // is within 2^-22 (2-norm) of it.
            r2 = warp_sum(r2);
            if (lane == 0 && !special) {
                float ra = __double2float_ru(sqrt(r2)) * 1.000001f;
                if (p.metric == kMetricCosine) ra += 2.3841858e-7f;
                local_resid = fmaxf(local_resid, ra);
            }
            if (p.metric == kMetricL2 && lane == 0 && !special) {
                double x = -0.5 * n2;
                __nv_bfloat16 h1 = __double2bfloat16(x);
                double r1 = x - (double)__bfloat162float(h1);
                __nv_bfloat16 h2 = __double2bfloat16(r1);
                double r2 = r1 - (double)__bfloat162float(h2);
                __nv_bfloat16 h3 = __double2bfloat16(r2);
                s[p.dim + 0] = h1;
                s[p.dim + 1] = h2;
                s[p.dim + 2] = h3;
            }
        }
    }
    if (lane == 0) {
        if (local_max > 0.f) atomicMax(p.max_norm_bits, __float_as_uint(local_max));
        if (local_resid > 0.f) atomicMax(p.max_resid_bits, __float_as_uint(local_resid));
        (void)local_special;
    }
}

__global__ void __launch_bounds__(256) prep_queries_kernel(PrepParams p, int b_pad) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const float u24 = 5.9604645e-8f;          // 2^-24
    const float t22 = 2.3841858e-7f;          // 2^-22
    const float tiny = (float)(p.dim + 4) * 2.8e-45f;  // (dim+4) * ~2^-148, underflow slack
    for (int q = warp; q < b_pad; q += nwarps) {
        if (q >= p.b) {  // padding rows of the GEMM operand
            if (p.q_shadow)
                for (int i = lane; i < p.qkp; i += 32) p.q_shadow[(size_t)q * p.qkp + i] = __float2bfloat16_rn(0.f);
            continue;
        }
        const float* src = p.queries + (size_t)q * p.dim;
        double n2 = 0.0, rb2 = 0.0;
        bool finite = true;
        for (int i = lane; i < p.pitch; i += 32) {
            float v = i < p.dim ? src[i] : 0.f;
            finite = finite && ((v - v) == 0.0f);
            n2 += (double)v * (double)v;
            const double df = (double)v - (double)__bfloat162float(__float2bfloat16_rn(v));   // exact
            rb2 += df * df;
            if (p.q_padded) p.q_padded[(size_t)q * p.pitch + i] = v;
        }
        n2 = warp_sum(n2);
        rb2 = warp_sum(rb2);
        finite = __all_sync(0xFFFFFFFFu, finite);
        if (p.q_shadow) {
            for (int i = lane; i < p.qkp; i += 32) {
                float v = 0.f;
                if (i < p.dim) v = src[i];
                else if (i == p.dim || (p.metric == kMetricL2 && i < p.dim + 3)) v = 1.0f;
                p.q_shadow[(size_t)q * p.qkp + i] = __float2bfloat16_rn(v);
            }
        }
        if (p.pub_keys)
            for (int i = lane; i < p.pub_stride; i += 32) p.pub_keys[(size_t)q * p.pub_stride + i] = 0xFFFFFFFFu;
        if (lane == 0) {
            const float nb = __double2float_ru(sqrt(n2));
            const float amax = __uint_as_float(*p.max_norm_bits) * 1.0000002f;
            float eps_abs, eps_rel;
            if (p.path == 1) {  // fp32 FMA accumulation in arbitrary order
                const float gamma = 1.05f * (float)(p.dim + 4) * u24;
                if (p.metric == kMetricIP) {
                    eps_abs = gamma * amax * nb + t22;
                    eps_rel = t22;
                } else if (p.metric == kMetricCosine) {
                    eps_abs = (gamma + 2.f * t22) * nb;
                    eps_rel = 0.f;
                } else {
                    eps_abs = 0.f;
                    eps_rel = gamma + t22;
                }
            } else {
                // bf16 operands, fp32 accumulation in the tensor core.  With a^ = a + da, b^ = b + db (the stored operands):
                //   a^.b^ - a.b = da.b^ + a.db,   |.| <= ra (nb + rb) + |a| rb
                // where ra = max over the index of the MEASURED |da| (append kernel), rb = the measured |db| of this query.
                // The accumulation of kp products in fp32 adds at most kp 2^-22 |a^| |b^|.
                const float ra = __uint_as_float(*p.max_resid_bits);
                const float rb = finite ? __double2float_ru(sqrt(rb2)) * 1.000001f : 0.f;
                const float nbh = nb + rb;
                const float acc = (float)p.kp * t22;
                if (p.metric == kMetricIP) {
                    eps_abs = 1.01f * (ra * nbh + amax * rb + acc * (amax + ra) * nbh) + t22;
                    eps_rel = t22;
                } else if (p.metric == kMetricCosine) {   // operand rows are unit vectors
                    eps_abs = 1.01f * (ra * nbh + rb + acc * (1.f + ra) * nbh) + 2.f * t22 * nb;
                    eps_rel = 0.f;
                } else {
                    eps_abs = 1.01f * (ra * nbh + amax * rb) + (float)p.kp * 2.f * t22 * (amax * nbh + amax * amax) + t22 * nb * nb;
                    eps_rel = t22;
                }
            }
            if (p.accum_f32) {
                // ANN_FLAG_ACCUM_F32: the value the answer is ordered by is itself a sequential fp32 sum, up to
                // gamma32 * sum|terms| away from the real distance; that error joins the filter's own on the badness scale.
                const float g32 = 1.05f * (float)(p.dim + 4) * u24;
                if (p.metric == kMetricIP) eps_abs += g32 * amax * nb;
                else if (p.metric == kMetricCosine) eps_abs += (2.f * g32 + 4.f * u24) * nb;
                else if (p.path == 1) eps_rel += g32;                       // scan: g = sum (a-b)^2, a sum of non-negative terms
                else eps_abs += 0.5f * g32 * (amax + nb) * (amax + nb) * 1.0001f;   // gemm: g = (L2^2 - |b|^2) / 2
            }
            QueryState st;
            st.eps_abs = eps_abs + tiny;
            st.eps_rel = eps_rel;
            st.qnorm = nb;
            st.qnorm2 = (float)n2;
            st.tau_key = 0xFF800000u;  // +inf
            st.pool_count = 0;
            st.special_count = 0;
            // Degenerate queries cannot be served by an approximate filter: a non-finite entry (every distance NaN), a zero
            // Cosine query (0/0), or magnitudes whose products could overflow fp32 in the tensor core.  They are flagged
            // up front and answered by the exact fallback (exact_fallback.cu).
            const bool bad = !finite || !((nb - nb) == 0.0f) || (p.metric == kMetricCosine && n2 == 0.0) ||
                             (p.path == 2 && !(amax * nb < 1e30f && amax < 1e37f));
            st.flags = bad ? kFlagSpecialOverflow : 0u;
            p.qstate[q] = st;
            if (bad) atomicAdd(p.bad_queries, 1ull);
        }
    }
}

}  // namespace

cudaError_t launch_append(const AppendParams& p, cudaStream_t stream) {
    if (p.n_new <= 0) return cudaSuccess;
    long long warps = p.n_new;
    long long blocks = (warps + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    append_rows_kernel<<<(int)blocks, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_prep_queries(const PrepParams& p, cudaStream_t stream) {
    int b_pad = p.q_shadow ? ((p.b + 127) / 128) * 128 : p.b;
    int blocks = (b_pad + 7) / 8;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    prep_queries_kernel<<<blocks, 256, 0, stream>>>(p, b_pad);
    return cudaGetLastError();
}

}  // namespace b200ann
