// Gauss-ILRMA with a shared ("partitioned") basis: src/bss/ilrma.py:368-408, :313-320, :493-495.
// The model variance is sum_k Z[n,k] T[f,k] V[k,t], i.e. an ordinary per-source model with the
// effective factors T_eff[n,f,k] = Z[n,k] T[f,k] and V_eff[n,k,t] = V[k,t].  The heavy passes over the
// mixture are therefore the ordinary source-model kernels (kernels_mu.cu) run on the effective factors
// in "raw statistics" mode; the small kernels here combine those statistics across sources and bins in
// a fixed order (deterministic) and keep Z, T, V.
#include "handle.h"

namespace {

__global__ void __launch_bounds__(256) part_expand_kernel(const float* latent, const float* basis, const float* act, float* beff,
                                                         float* aeff, int B, int N, int F, int K, int Tp) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nb = (long long)B * N * F * K;
    const long long na = (long long)B * N * K * Tp;
    if (idx < nb) {
        const int k = (int)(idx % K);
        long long r = idx / K;
        const int f = (int)(r % F);
        r /= F;
        const int n = (int)(r % N);
        const int b = (int)(r / N);
        beff[idx] = latent[((size_t)b * N + n) * K + k] * basis[((size_t)b * F + f) * K + k];
    } else if (idx < nb + na) {
        const long long j = idx - nb;
        const int t = (int)(j % Tp);
        long long r = j / Tp;
        const int k = (int)(r % K);
        r /= K;
        const int b = (int)(r / N);
        aeff[j] = act[((size_t)b * K + k) * Tp + t];
    }
}

// block per (b, n, k): z = sqrt(sum_f T[f,k] num[n,f,k] / max(sum_f T[f,k] den[n,f,k], eps))     ilrma.py:378-382
__global__ void __launch_bounds__(256) part_latent_kernel(const float* raw, const float* basis, float* zraw, int N, int F, int K,
                                                         double eps) {
    __shared__ double red[2][8];
    const long long bnk = blockIdx.x;
    const int k = (int)(bnk % K);
    const long long bn = bnk / K;
    const int b = (int)(bn / N);
    double num = 0.0, den = 0.0;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        const double t = (double)basis[((size_t)b * F + f) * K + k];
        const float* r = raw + (((size_t)bn * F + f) * K + k) * 2;
        num += t * (double)r[0];
        den += t * (double)r[1];
    }
    num = warp_sum(num);
    den = warp_sum(den);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = num;
        red[1][threadIdx.x >> 5] = den;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            a += red[0][w];
            c += red[1][w];
        }
        if (c < eps) c = eps;
        zraw[bnk] = (float)sqrt(a / c);
    }
}

// Z = Z / Z.sum(axis=0)      ilrma.py:383
__global__ void __launch_bounds__(64) part_latent_norm_kernel(const float* zraw, float* latent, int B, int N, int K) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * K) return;
    const int k = idx % K, b = idx / K;
    double s = 0.0;
    for (int n = 0; n < N; ++n) s += (double)zraw[((size_t)b * N + n) * K + k];
    for (int n = 0; n < N; ++n) latent[((size_t)b * N + n) * K + k] = (float)((double)zraw[((size_t)b * N + n) * K + k] / s);
}

// T[f,k] *= sqrt(sum_n Z[n,k] num[n,f,k] / max(sum_n Z[n,k] den[n,f,k], eps))      ilrma.py:390-394
__global__ void __launch_bounds__(256) part_basis_kernel(const float* raw, const float* latent, float* basis, int B, int N, int F,
                                                        int K, double eps) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * F * K) return;
    const int k = (int)(idx % K);
    const long long bf = idx / K;
    const int f = (int)(bf % F);
    const int b = (int)(bf / F);
    double num = 0.0, den = 0.0;
    for (int n = 0; n < N; ++n) {
        const double z = (double)latent[((size_t)b * N + n) * K + k];
        const float* r = raw + ((((size_t)b * N + n) * F + f) * K + k) * 2;
        num += z * (double)r[0];
        den += z * (double)r[1];
    }
    if (den < eps) den = eps;
    basis[idx] = (float)((double)basis[idx] * sqrt(num / den));
}

// V[k,t] *= sqrt(sum_{n,chunks} num / max(sum den, eps)); part [B][n_chunks][N][K][2][Tp]      ilrma.py:401-405
__global__ void __launch_bounds__(256) part_act_finish_kernel(const float* part, float* act, int B, int N, int K, int T, int Tp,
                                                             int n_chunks, double eps) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * K * Tp) return;
    const int t = (int)(idx % Tp);
    const long long bk = idx / Tp;
    const int k = (int)(bk % K);
    const int b = (int)(bk / K);
    if (t >= T) {
        act[idx] = 0.f;
        return;
    }
    double num = 0.0, den = 0.0;
    for (int c = 0; c < n_chunks; ++c)
        for (int n = 0; n < N; ++n) {
            const float* src = part + (((((size_t)b * n_chunks + c) * N + n) * K + k) * 2) * Tp + t;
            num += (double)src[0];
            den += (double)src[Tp];
        }
    if (den < eps) den = eps;
    act[idx] = (float)((double)act[idx] * sqrt(num / den));
}

// block per b: Zaux = Z / aux^2; s[k] = sum_n Zaux; T[:,k] *= s[k]; Z = Zaux / s       ilrma.py:313-320 (domain = 2)
__global__ void __launch_bounds__(256) part_normalize_kernel(const double* aux, float* latent, float* basis, int N, int F, int K) {
    __shared__ double s[64];
    const int b = blockIdx.x;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double acc = 0.0;
        for (int n = 0; n < N; ++n) {
            const double a = aux[(size_t)b * N + n];
            acc += (double)latent[((size_t)b * N + n) * K + k] / (a * a);
        }
        s[k] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < F * K; i += blockDim.x) {
        float* p = basis + (size_t)b * F * K + i;
        *p = (float)((double)*p * s[i % K]);
    }
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
        const int n = i / K, k = i - n * K;
        const double a = aux[(size_t)b * N + n];
        float* p = latent + (size_t)b * N * K + i;
        *p = (float)((double)*p / (a * a) / s[k]);
    }
}

}  // namespace

int launch_part_expand(bss_handle* h) {
    const long long n = (long long)h->B * h->N * h->F * h->K + (long long)h->B * h->N * h->K * h->Tp;
    part_expand_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(h->latent, h->basis, h->act, h->beff, h->aeff, h->B, h->N, h->F,
                                                                     h->K, h->Tp);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_part_latent(bss_handle* h) {
    part_latent_kernel<<<(unsigned)((long long)h->B * h->N * h->K), 256, 0, h->stream>>>(h->praw, h->basis, h->latent2, h->N, h->F, h->K,
                                                                                        h->cfg.eps);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    part_latent_norm_kernel<<<(unsigned)cdiv((long long)h->B * h->K, 64), 64, 0, h->stream>>>(h->latent2, h->latent, h->B, h->N, h->K);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_part_basis(bss_handle* h) {
    const long long n = (long long)h->B * h->F * h->K;
    part_basis_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(h->praw, h->latent, h->basis, h->B, h->N, h->F, h->K, h->cfg.eps);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_part_act_finish(bss_handle* h, int n_chunks) {
    const long long n = (long long)h->B * h->K * h->Tp;
    part_act_finish_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(h->part, h->act, h->B, h->N, h->K, h->T, h->Tp, n_chunks,
                                                                         h->cfg.eps);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_part_normalize(bss_handle* h) {
    part_normalize_kernel<<<h->B, 256, 0, h->stream>>>(h->aux, h->latent, h->basis, h->N, h->F, h->K);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}
