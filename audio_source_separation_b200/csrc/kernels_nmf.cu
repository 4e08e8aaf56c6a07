// Single-channel NMF multiplicative updates (src/algorithm/nmf.py:150-595) in fp64.
// The problems are tiny (cfg1: a 257 x 128 target, 263 KB) and latency bound, so everything is
// kept in double precision -- parity with the reference is then ~1e-12 -- and the work is spread
// over many small CTAs:
//   basis T      : one warp per (problem, bin) row, reduction over frames
//   activation V : one thread per frame, deterministic two-stage reduction over bin chunks
//   loss         : one warp per row, then the shared fixed-order finish kernel
// Every variant has the form  F <- F * g( sum s1(z, tv) * other, sum s2(z, tv) * other ).
#include <cooperative_groups.h>
#include <cstdlib>

#include "handle.h"

namespace cg = cooperative_groups;

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

// criterion((T V)^(2/domain), target) of one element     src/algorithm/nmf.py:172-174 etc.
__device__ __forceinline__ double nmf_loss_term(const NmfMath& m, double zz, double tv) {
    const double A = m.kind == 4 ? tv : powq(tv, m.loss_expo);
    if (m.kind == 0) return (zz - A) * (zz - A);
    const double a = A + m.loss_eps, zt = zz + m.loss_eps;
    if (m.kind == 1) return zt * log(zt / a) + a - zt;                                    // src/criterion/divergence.py:34-45
    if (m.kind == 2) {
        const double r = zt / a;                                                         // divergence.py:21-32
        return r - log(r) - 1.0;
    }
    if (m.kind == 3) return log(a) + (2.0 + m.nu) / 2.0 * log(1.0 + (2.0 / m.nu) * (zt / a));   // nmf.py:367-371
    return log(zt / a) + 1.5 * log((2.0 * zt * zt + a * a) / (3.0 * zt * zt));                 // nmf.py:434-441
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
        const double l = nmf_loss_term(m, z[t], tv);
        s += l;
    }
    s = warp_sum(s);
    if (lane == 0) terms[row] = s;
}

// ------------------------------------------------------------------------------------------- fused cluster kernel
// Small problems (cfg1: 257 x 128, K = 4) are pure launch latency when every phase is its own kernel.  Here one
// thread-block CLUSTER of 8 CTAs owns a problem for ALL iterations: the target is split by rows over the CTAs and
// stays in shared memory (fp64), every CTA keeps its basis rows and a full copy of the activation, the basis update
// is CTA-local, the activation update reduces the 8 per-CTA partial sums through distributed shared memory in a
// fixed order (every CTA computes the same bits), and the optional per-iteration loss is gathered by rank 0.
// Three cluster barriers per iteration replace five kernel launches.
constexpr int NMF_CL = 8;

struct NmfFusedParams {
    NmfMath m;
    const double* Z;
    double* Tm;
    double* V;
    double* loss_hist;   // [n_iter][B] or null
    int B, F, T, K, Fr, n_iter;
};

// NT threads per CTA: the work is fp64 latency, so every SM of the cluster runs as many warps as the register budget of
// the per-thread accumulators (4 K doubles) allows: 1024 threads up to K = 3, 512 above.  The activation partial sums
// use 128 frame lanes x NT / 128 row groups.
template <int KT, int NT>
__global__ void __launch_bounds__(NT) nmf_fused_kernel(const NmfFusedParams p) {
    constexpr int NMF_FT = NT;
    constexpr int NMF_RG = NT / 128;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ double nmf_smem[];
    const int r = (int)cluster.block_rank();
    const int b = (int)(blockIdx.x / NMF_CL);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int F = p.F, T = p.T, Fr = p.Fr;
    constexpr int K = KT;
    const NmfMath& m = p.m;
    double* Zs = nmf_smem;            // [Fr][T]
    double* Ts = Zs + (size_t)Fr * T; // [Fr][K]
    double* Vs = Ts + Fr * K;         // [K][T]
    double* Pn = Vs + K * T;          // [K][T] partial numerators of this CTA's rows
    double* Pd = Pn + K * T;
    double* Gn = Pd + K * T;          // [RG][K][T] per row group
    double* Gd = Gn + NMF_RG * K * T;
    double* red = Gd + NMF_RG * K * T;   // [40]
    const int f0 = r * Fr;
    const int nrows = max(0, min(Fr, F - f0));
    const double* Zg = p.Z + ((size_t)b * F + f0) * T;
    double* Tg = p.Tm + ((size_t)b * F + f0) * K;
    double* Vg = p.V + (size_t)b * K * T;
    for (int i = tid; i < nrows * T; i += blockDim.x) Zs[i] = Zg[i];
    for (int i = tid; i < nrows * K; i += blockDim.x) Ts[i] = Tg[i];
    for (int i = tid; i < K * T; i += blockDim.x) Vs[i] = Vg[i];
    __syncthreads();
    for (int it = 0; it < p.n_iter; ++it) {
        // basis rows of this CTA (src/algorithm/nmf.py: "update basis" half of every update_once_*)
        for (int row = warp; row < nrows; row += NMF_FT / 32) {
            double num[K], den[K], tk[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                num[k] = den[k] = 0.0;
                tk[k] = Ts[row * K + k];
            }
            for (int t = lane; t < T; t += 32) {
                double tv = 0.0, vk[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    vk[k] = Vs[k * T + t];
                    tv = fma(tk[k], vk[k], tv);
                }
                double s1, s2;
                nmf_stats(m, Zs[row * T + t], tv, s1, s2);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    num[k] = fma(s1, vk[k], num[k]);
                    den[k] = fma(s2, vk[k], den[k]);
                }
            }
            __syncwarp();   // every lane has read the old basis row (tk) before lane k overwrites element k
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double a = warp_sum(num[k]);
                const double c = warp_sum(den[k]);
                if (lane == k) Ts[row * K + k] = tk[k] * nmf_combine(m, a, c);
            }
        }
        __syncthreads();
        // partial sums of the activation update: 128 frame lanes x NMF_RG row groups
        {
            const int tl = tid & 127, g = tid >> 7;
            for (int t = tl; t < T; t += 128) {
                double num[K], den[K], vk[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    num[k] = den[k] = 0.0;
                    vk[k] = Vs[k * T + t];
                }
                for (int row = g; row < nrows; row += NMF_RG) {
                    double tv = 0.0, tk[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        tk[k] = Ts[row * K + k];
                        tv = fma(tk[k], vk[k], tv);
                    }
                    double s1, s2;
                    nmf_stats(m, Zs[row * T + t], tv, s1, s2);
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        num[k] = fma(s1, tk[k], num[k]);
                        den[k] = fma(s2, tk[k], den[k]);
                    }
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    Gn[(g * K + k) * T + t] = num[k];
                    Gd[(g * K + k) * T + t] = den[k];
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < K * T; i += blockDim.x) {
            double num = 0.0, den = 0.0;
#pragma unroll
            for (int g = 0; g < NMF_RG; ++g) {
                num += Gn[g * K * T + i];
                den += Gd[g * K * T + i];
            }
            Pn[i] = num;
            Pd[i] = den;
        }
        cluster.sync();
        // every CTA adds the 8 partial sums in rank order (distributed shared memory) and updates its copy of V
        for (int i = tid; i < K * T; i += blockDim.x) {
            double num = 0.0, den = 0.0;
#pragma unroll
            for (int q = 0; q < NMF_CL; ++q) {
                num += cluster.map_shared_rank(Pn, q)[i];   // mapped on the fly: 24 remote pointers would not fit in 64 registers
                den += cluster.map_shared_rank(Pd, q)[i];
            }
            Vs[i] *= nmf_combine(m, num, den);
        }
        cluster.sync();
        if (p.loss_hist) {
            double s = 0.0;
            for (int row = warp; row < nrows; row += NMF_FT / 32)
                for (int t = lane; t < T; t += 32) {
                    double tv = 0.0;
#pragma unroll
                    for (int k = 0; k < K; ++k) tv = fma(Ts[row * K + k], Vs[k * T + t], tv);
                    s += nmf_loss_term(m, Zs[row * T + t], tv);
                }
            s = warp_sum(s);
            if (lane == 0) red[8 + warp] = s;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int w = 0; w < NMF_FT / 32; ++w) t += red[8 + w];
                red[0] = t;
            }
            cluster.sync();
            if (r == 0 && tid == 0) {
                double t = 0.0;
#pragma unroll
                for (int q = 0; q < NMF_CL; ++q) t += cluster.map_shared_rank(red, q)[0];
                p.loss_hist[(size_t)it * p.B + b] = t;
            }
        }
    }
    cluster.sync();   // nobody leaves while its shared memory may still be read remotely
    for (int i = tid; i < nrows * K; i += blockDim.x) Tg[i] = Ts[i];
    if (r == 0)
        for (int i = tid; i < K * T; i += blockDim.x) Vg[i] = Vs[i];
}

template <int KT>
int launch_nmf_fused_t(bss_handle* h, const NmfFusedParams& p) {
    constexpr int NT = KT <= 3 ? 1024 : 512;
    constexpr int RG = NT / 128;
    const size_t smem = ((size_t)p.Fr * p.T + (size_t)p.Fr * KT + (3 + 2 * RG) * (size_t)KT * p.T + 40) * sizeof(double);
    if (smem > (size_t)h->max_smem - 1024) return BSS_EUNSUPPORTED;   // caller falls back to the per-phase kernels
    static bool attr_done = false;
    if (!attr_done) {
        BSS_CUDA(h, cudaFuncSetAttribute(nmf_fused_kernel<KT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.B * NMF_CL), 1, 1);
    cfg.blockDim = dim3(NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NMF_CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    BSS_CUDA(h, cudaLaunchKernelEx(&cfg, nmf_fused_kernel<KT, NT>, p));
    h->launches++;
    return BSS_OK;
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

// whole loop of `n_iter` updates (and losses, when loss_hist is non-null) in one cluster launch; *done = false when the
// problem does not fit (K > 8 or the row slice exceeds shared memory) and the per-phase kernels must be used
int launch_nmf_fused(bss_handle* h, const NmfMath& m, const double* Z, double* Tm, double* V, double* loss_hist, int B, int F, int T,
                     int K, int n_iter, bool* done) {
    *done = false;
    if (K > NMF_KC || n_iter < 1 || getenv("BSSGPU_NO_CLUSTER")) return BSS_OK;
    NmfFusedParams p{};
    p.m = m;
    p.Z = Z;
    p.Tm = Tm;
    p.V = V;
    p.loss_hist = loss_hist;
    p.B = B;
    p.F = F;
    p.T = T;
    p.K = K;
    p.Fr = (int)cdiv(F, NMF_CL);
    p.n_iter = n_iter;
    int rc = BSS_OK;
    switch (K) {
        case 1: rc = launch_nmf_fused_t<1>(h, p); break;
        case 2: rc = launch_nmf_fused_t<2>(h, p); break;
        case 3: rc = launch_nmf_fused_t<3>(h, p); break;
        case 4: rc = launch_nmf_fused_t<4>(h, p); break;
        case 5: rc = launch_nmf_fused_t<5>(h, p); break;
        case 6: rc = launch_nmf_fused_t<6>(h, p); break;
        case 7: rc = launch_nmf_fused_t<7>(h, p); break;
        default: rc = launch_nmf_fused_t<8>(h, p); break;
    }
    if (rc == BSS_EUNSUPPORTED) return BSS_OK;
    if (rc == BSS_OK) *done = true;
    return rc;
}
