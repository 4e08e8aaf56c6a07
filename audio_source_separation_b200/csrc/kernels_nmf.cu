// Single-channel NMF multiplicative updates (src/algorithm/nmf.py:150-595) in fp64.
// The problems are tiny (cfg1: a 257 x 128 target, 263 KB) and latency bound, so everything is
// kept in double precision -- parity with the reference is then ~1e-12 -- and the work is spread
// over many small CTAs:
//   basis T      : one warp per (problem, bin) row, reduction over frames
//   activation V : one thread per frame, deterministic two-stage reduction over bin chunks
//   loss         : one warp per row, then the shared fixed-order finish kernel
// Every variant has the form  F <- F * g( sum s1(z, tv) * other, sum s2(z, tv) * other ).
#include "handle.h"

namespace {

constexpr int NMF_KC = 8;   // basis vectors accumulated per pass

__device__ __forceinline__ double powq(double x, double e) {
    if (e == 1.0) return x;
    if (e == 0.0) return 1.0;
    if (e == 0.5) return sqrt(x);
    if (e == 2.0) return x * x;
    if (e == 3.0) return x * x * x;
    return pow(x, e);
}

__device__ __forceinline__ void nmf_stats(const NmfMath& m, double z, double tv, double& s1, double& s2) {
    const double eps = m.eps;
    switch (m.kind) {
        case 0:   // EUC  src/algorithm/nmf.py:189-205
            tv = tv < eps ? eps : tv;
            s1 = z * powq(tv, m.b);
            s2 = powq(tv, m.a);
            break;
        case 1:   // KL   :249-264
            tv = tv < eps ? eps : tv;
            s1 = z / tv;
            s2 = powq(tv, m.b);
            break;
        case 2:   // IS   :310-325, :339-354
            tv = tv < eps ? eps : tv;
            s1 = z / powq(tv, m.p);
            s2 = 1.0 / tv;
            break;
        case 3: {   // t   :406-426
            tv = tv < eps ? eps : tv;
            const double zf = z > eps ? z : eps;
            const double h = 1.0 / (2.0 / ((2.0 + m.nu) * tv) + m.nu / ((2.0 + m.nu) * zf));
            s1 = h / (tv * tv);
            s2 = 1.0 / tv;
            break;
        }
        default:   // Cauchy  :461-595
            if (m.alg == BSS_ALG_NAIVE || m.alg == BSS_ALG_MM) {
                tv = tv < eps ? eps : tv;
                double c = 2.0 * z + tv * tv;
                c = c < eps ? eps : c;
                s1 = 1.0 / tv;
                s2 = 3.0 * (tv / c);
            } else if (m.alg == BSS_ALG_ME) {
                double s = tv * tv + z;
                s = s < eps ? eps : s;
                s1 = 0.75 * (tv / s);
                s2 = 1.0 / tv;
            } else {   // mm_fast
                double c = 2.0 * z + tv * tv;
                double ctv = c * tv;
                ctv = ctv < eps ? eps : ctv;
                s1 = z / ctv;
                c = c < eps ? eps : c;
                s2 = tv / c;
            }
            break;
    }
}

__device__ __forceinline__ double nmf_combine(const NmfMath& m, double a, double b) {
    const double eps = m.eps;
    if (m.kind == 4) {
        if (m.alg == BSS_ALG_ME) {   // a = A, b = B:  B / max(A + sqrt(A^2 + 2 B A), eps)
            double d = a + sqrt(a * a + 2.0 * b * a);
            d = d < eps ? eps : d;
            return b / d;
        }
        b = b < eps ? eps : b;
        const double r = a / b;
        return m.alg == BSS_ALG_NAIVE ? r : sqrt(r);
    }
    b = b < eps ? eps : b;
    return powq(a / b, m.q);
}

// ------------------------------------------------------------------------------------------- basis
// T[f,k] *= g(sum_t s1 V[k,t], sum_t s2 V[k,t]); one warp per (b, f); old row kept in shared memory
__global__ void __launch_bounds__(128) nmf_basis_kernel(const NmfMath m, const double* Z, double* Tm, const double* V, int B, int F,
                                                        int T, int K) {
    extern __shared__ double nmf_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= (long long)B * F) return;
    const int b = (int)(row / F);
    double* trow = nmf_smem + (size_t)warp * K;
    double* tg = Tm + (size_t)row * K;
    for (int k = lane; k < K; k += 32) trow[k] = tg[k];
    __syncwarp();
    const double* z = Z + (size_t)row * T;
    const double* v = V + (size_t)b * K * T;
    for (int k0 = 0; k0 < K; k0 += NMF_KC) {
        double num[NMF_KC], den[NMF_KC];
#pragma unroll
        for (int kk = 0; kk < NMF_KC; ++kk) num[kk] = den[kk] = 0.0;
        for (int t = lane; t < T; t += 32) {
            double tv = 0.0;
            double vk[NMF_KC];
#pragma unroll
            for (int kk = 0; kk < NMF_KC; ++kk) vk[kk] = 0.0;
            for (int k = 0; k < K; ++k) {
                const double vv = v[(size_t)k * T + t];
                tv = fma(trow[k], vv, tv);
#pragma unroll
                for (int kk = 0; kk < NMF_KC; ++kk)
                    if (k == k0 + kk) vk[kk] = vv;
            }
            double s1, s2;
            nmf_stats(m, z[t], tv, s1, s2);
#pragma unroll
            for (int kk = 0; kk < NMF_KC; ++kk) {
                num[kk] = fma(s1, vk[kk], num[kk]);
                den[kk] = fma(s2, vk[kk], den[kk]);
            }
        }
#pragma unroll
        for (int kk = 0; kk < NMF_KC; ++kk) {
            const double a = warp_sum(num[kk]);
            const double c = warp_sum(den[kk]);
            if (lane == kk && k0 + kk < K) tg[k0 + kk] = trow[k0 + kk] * nmf_combine(m, a, c);
        }
    }
}

// ------------------------------------------------------------------------------------------- activation
// stage 1: block = 128 consecutive frames x one chunk of bins x one group of NMF_KC basis vectors;
// part: [B][n_chunks][K][2][T]
__global__ void __launch_bounds__(128) nmf_act_partial_kernel(const NmfMath m, const double* Z, const double* Tm, const double* V,
                                                              double* part, int F, int T, int K, int n_chunks, int bins_per_chunk) {
    extern __shared__ double nmf_smem[];   // V columns of this block: [K][128]
    const int t = blockIdx.x * 128 + threadIdx.x;
    const int chunk = blockIdx.y % n_chunks;
    const int k0 = (blockIdx.y / n_chunks) * NMF_KC;
    const int b = blockIdx.z;
    const bool live = t < T;
    const double* v = V + (size_t)b * K * T;
    for (int k = 0; k < K; ++k) nmf_smem[k * 128 + threadIdx.x] = live ? v[(size_t)k * T + t] : 0.0;
    // each thread reads back only what it wrote: no barrier needed
    double num[NMF_KC], den[NMF_KC];
#pragma unroll
    for (int kk = 0; kk < NMF_KC; ++kk) num[kk] = den[kk] = 0.0;
    const int f_begin = chunk * bins_per_chunk;
    const int f_end = min(F, f_begin + bins_per_chunk);
    if (live) {
        for (int f = f_begin; f < f_end; ++f) {
            const double* trow = Tm + ((size_t)b * F + f) * K;
            double tv = 0.0;
            double tk[NMF_KC];
#pragma unroll
            for (int kk = 0; kk < NMF_KC; ++kk) tk[kk] = 0.0;
            for (int k = 0; k < K; ++k) {
                const double tt = __ldg(trow + k);
                tv = fma(tt, nmf_smem[k * 128 + threadIdx.x], tv);
#pragma unroll
                for (int kk = 0; kk < NMF_KC; ++kk)
                    if (k == k0 + kk) tk[kk] = tt;
            }
            double s1, s2;
            nmf_stats(m, Z[((size_t)b * F + f) * T + t], tv, s1, s2);
#pragma unroll
            for (int kk = 0; kk < NMF_KC; ++kk) {
                num[kk] = fma(s1, tk[kk], num[kk]);
                den[kk] = fma(s2, tk[kk], den[kk]);
            }
        }
#pragma unroll
        for (int kk = 0; kk < NMF_KC; ++kk)
            if (k0 + kk < K) {
                double* dst = part + ((((size_t)b * n_chunks + chunk) * K + k0 + kk) * 2) * T + t;
                dst[0] = num[kk];
                dst[T] = den[kk];
            }
    }
}

__global__ void __launch_bounds__(256) nmf_act_finish_kernel(const NmfMath m, const double* part, double* V, int B, int T, int K,
                                                             int n_chunks) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * K * T) return;
    const int t = (int)(idx % T);
    const long long bk = idx / T;
    const int k = (int)(bk % K);
    const int b = (int)(bk / K);
    double num = 0.0, den = 0.0;
    for (int c = 0; c < n_chunks; ++c) {
        const double* src = part + ((((size_t)b * n_chunks + c) * K + k) * 2) * T + t;
        num += src[0];
        den += src[T];
    }
    V[idx] *= nmf_combine(m, num, den);
}

// ------------------------------------------------------------------------------------------- loss
// per row: sum_t criterion((T V)^(2/domain), target)     src/algorithm/nmf.py:172-174 etc.
__global__ void __launch_bounds__(128) nmf_loss_kernel(const NmfMath m, const double* Z, const double* Tm, const double* V,
                                                       double* terms, int B, int F, int T, int K) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= (long long)B * F) return;
    const int b = (int)(row / F);
    const double* trow = Tm + (size_t)row * K;
    const double* z = Z + (size_t)row * T;
    const double* v = V + (size_t)b * K * T;
    double s = 0.0;
    for (int t = lane; t < T; t += 32) {
        double tv = 0.0;
        for (int k = 0; k < K; ++k) tv = fma(__ldg(trow + k), v[(size_t)k * T + t], tv);
        const double A = m.kind == 4 ? tv : powq(tv, m.loss_expo);
        const double zz = z[t];
        double l;
        if (m.kind == 0) {
            l = (zz - A) * (zz - A);
        } else {
            const double a = A + m.loss_eps, zt = zz + m.loss_eps;
            if (m.kind == 1)
                l = zt * log(zt / a) + a - zt;                                     // src/criterion/divergence.py:34-45
            else if (m.kind == 2) {
                const double r = zt / a;                                            // divergence.py:21-32
                l = r - log(r) - 1.0;
            } else if (m.kind == 3)
                l = log(a) + (2.0 + m.nu) / 2.0 * log(1.0 + (2.0 / m.nu) * (zt / a));   // nmf.py:367-371
            else
                l = log(zt / a) + 1.5 * log((2.0 * zt * zt + a * a) / (3.0 * zt * zt));   // nmf.py:434-441
        }
        s += l;
    }
    s = warp_sum(s);
    if (lane == 0) terms[row] = s;
}

}  // namespace

int launch_nmf_update(bss_handle* h, const NmfMath& m, const double* Z, double* Tm, double* V, int B, int F, int T, int K) {
    {
        const long long rows = (long long)B * F;
        const int wpc = 4;
        nmf_basis_kernel<<<(unsigned)cdiv(rows, wpc), wpc * 32, (size_t)wpc * K * sizeof(double), h->stream>>>(m, Z, Tm, V, B, F, T,
                                                                                                           K);
        h->launches++;
        BSS_CUDA(h, cudaGetLastError());
    }
    const int t_blocks = (int)cdiv(T, 128);
    const int n_kc = (int)cdiv(K, NMF_KC);
    long long want = (long long)h->n_sm * 4;
    int n_chunks = (int)cdiv(want, (long long)t_blocks * n_kc * B);
    if (n_chunks < 1) n_chunks = 1;
    int bins_per_chunk = (int)cdiv(F, n_chunks);
    if (bins_per_chunk < 8) bins_per_chunk = F < 8 ? F : 8;
    n_chunks = (int)cdiv(F, bins_per_chunk);
    const size_t need = (size_t)B * n_chunks * K * 2 * T;
    if (need > h->npart_elems) {
        if (h->npart) cudaFree(h->npart);
        h->npart = nullptr;
        h->npart_elems = 0;
        BSS_CUDA(h, cudaMalloc(&h->npart, need * sizeof(double)));
        h->npart_elems = need;
    }
    const size_t smem = (size_t)K * 128 * sizeof(double);
    static bool attr_done = false;
    if (!attr_done) {
        BSS_CUDA(h, cudaFuncSetAttribute(nmf_act_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 128 * 8));
        attr_done = true;
    }
    dim3 grid(t_blocks, n_chunks * n_kc, B);
    nmf_act_partial_kernel<<<grid, 128, smem, h->stream>>>(m, Z, Tm, V, h->npart, F, T, K, n_chunks, bins_per_chunk);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    const long long total = (long long)B * K * T;
    nmf_act_finish_kernel<<<(unsigned)cdiv(total, 256), 256, 0, h->stream>>>(m, h->npart, V, B, T, K, n_chunks);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_nmf_loss(bss_handle* h, const NmfMath& m, const double* Z, const double* Tm, const double* V, double* terms, int B, int F,
                    int T, int K) {
    const long long rows = (long long)B * F;
    nmf_loss_kernel<<<(unsigned)cdiv(rows, 4), 128, 0, h->stream>>>(m, Z, Tm, V, terms, B, F, T, K);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}
