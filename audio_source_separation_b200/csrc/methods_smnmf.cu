// Sawada's multichannel IS-NMF orchestration: src/bss/mnmf.py:116-635 (MultichannelISNMF, author='Sawada').
// State (all fp64): H = h->sH [B][F][N][C*C] packed Hermitian, Z = h->sZ [B][N][K], T = h->sT [B][F][K], V = h->sV [B][K][T].
#include <vector>

#include "methods.h"

namespace {

template <typename T>
int dalloc(bss_handle* h, T** p, size_t n) {
    if (n == 0) n = 1;
    BSS_CUDA(h, cudaMalloc((void**)p, n * sizeof(T)));
    BSS_CUDA(h, cudaMemsetAsync(*p, 0, n * sizeof(T), h->stream));
    return BSS_OK;
}

int put(bss_handle* h, double* dev, const void* src, size_t n) {
    BSS_CUDA(h, cudaMemcpyAsync(dev, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    BSS_CUDA(h, bss_wait(h));
    return BSS_OK;
}
int get(bss_handle* h, const double* dev, void* dst, size_t n) {
    BSS_CUDA(h, cudaMemcpyAsync(dst, dev, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    BSS_CUDA(h, bss_wait(h));
    return BSS_OK;
}

}  // namespace

int smnmf_allocate(bss_handle* h) {
    const size_t B = h->B, C = h->C, N = h->N, F = h->F, T = h->T, Tp = h->Tp, K = h->K;
    int bins = 0, chunks = 0;
    smnmf_act_plan(h, &bins, &chunks);
    BSS_TRY(dalloc(h, &h->X, B * F * C * Tp));
    BSS_TRY(dalloc(h, &h->lossbuf, B * F + B));
    BSS_TRY(dalloc(h, &h->sH, B * F * N * C * C));
    BSS_TRY(dalloc(h, &h->sZ, B * N * K));
    BSS_TRY(dalloc(h, &h->sT, B * F * K));
    BSS_TRY(dalloc(h, &h->sV, B * K * T));
    BSS_TRY(dalloc(h, &h->sStat, 2 * B * N * F * T));
    const size_t part_nk = 2 * B * F * N * K, part_act = 2 * B * (size_t)chunks * K * T;
    BSS_TRY(dalloc(h, &h->sPart, part_nk > part_act ? part_nk : part_act));
    BSS_TRY(dalloc(h, &h->sAcc, B * F * N * 2 * C * C));
    return BSS_OK;
}

// H[f,n] = I      src/bss/mnmf.py:225-229
int smnmf_reset(bss_handle* h) {
    const size_t C = h->C, CC = C * C, n_mat = (size_t)h->B * h->F * h->N;
    std::vector<double> eye(n_mat * CC, 0.0);
    for (size_t m = 0; m < n_mat; ++m)
        for (size_t c = 0; c < C; ++c) eye[m * CC + c] = 1.0;
    return put(h, h->sH, eye.data(), eye.size());
}

// update_once_sawada: basis, activation, latent, spatial, each from a freshly reconstructed model      mnmf.py:311-315
int smnmf_update_once(bss_handle* h) {
    for (int which = 0; which < 3; ++which) {
        BSS_TRY(launch_smnmf_stats(h));
        BSS_TRY(launch_smnmf_factor(h, which));
    }
    return launch_smnmf_spatial(h, h->cfg.normalize != BSS_NORMALIZE_NONE);
}

// compute_negative_loglikelihood_sawada      mnmf.py:575-589
int smnmf_loss(bss_handle* h) {
    const size_t BF = (size_t)h->B * h->F;
    double* result = h->lossbuf + BF;
    BSS_CUDA(h, cudaMemsetAsync(result, 0, sizeof(double) * h->B, h->stream));
    BSS_TRY(launch_smnmf_loss_terms(h, h->lossbuf));
    return launch_loss_finish(h, h->lossbuf, nullptr, 0.0, h->B, h->F, result);
}

int smnmf_separate(bss_handle* h, cf* out) { return launch_smnmf_separate(h, out); }

// host layouts: spatial (F,N,C,C) complex128, latent (N,K), basis (F,K), activation (K,T) float64
int smnmf_set_state(bss_handle* h, int which, const void* src, int dtype) {
    const size_t B = h->B, C = h->C, CC = C * C, N = h->N, F = h->F, T = h->T, K = h->K;
    switch (which) {
        case BSS_STATE_SPATIAL: {
            if (dtype != BSS_C128) return bss_fail(h, BSS_EINVAL, "spatial is exchanged as complex128");
            // Hermitian part, packed
            const size_t n_mat = B * F * N;
            std::vector<double> packed(n_mat * CC);
            const double* s = (const double*)src;
            for (size_t m = 0; m < n_mat; ++m) {
                const double* o = s + m * CC * 2;
                double* q = packed.data() + m * CC;
                for (size_t i = 0; i < C; ++i) q[i] = o[(i * C + i) * 2];
                size_t e = C;
                for (size_t i = 1; i < C; ++i)
                    for (size_t j = 0; j < i; ++j) {
                        q[e] = 0.5 * (o[(i * C + j) * 2] + o[(j * C + i) * 2]);
                        q[e + 1] = 0.5 * (o[(i * C + j) * 2 + 1] - o[(j * C + i) * 2 + 1]);
                        e += 2;
                    }
            }
            return put(h, h->sH, packed.data(), packed.size());
        }
        case BSS_STATE_LATENT:
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "latent is exchanged as float64");
            return put(h, h->sZ, src, B * N * K);
        case BSS_STATE_BASIS:
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "basis is exchanged as float64");
            return put(h, h->sT, src, B * F * K);
        case BSS_STATE_ACTIVATION:
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "activation is exchanged as float64");
            return put(h, h->sV, src, B * K * T);
    }
    return bss_fail(h, BSS_EINVAL, "state cannot be set");
}

int smnmf_get_state(bss_handle* h, int which, void* dst, int dtype) {
    const size_t B = h->B, C = h->C, CC = C * C, N = h->N, F = h->F, T = h->T, K = h->K;
    switch (which) {
        case BSS_STATE_SPATIAL: {
            if (dtype != BSS_C128) return bss_fail(h, BSS_EINVAL, "spatial is exchanged as complex128");
            const size_t n_mat = B * F * N;
            std::vector<double> packed(n_mat * CC);
            BSS_TRY(get(h, h->sH, packed.data(), packed.size()));
            double* out = (double*)dst;
            for (size_t m = 0; m < n_mat; ++m) {
                const double* q = packed.data() + m * CC;
                double* o = out + m * CC * 2;
                for (size_t i = 0; i < C; ++i) {
                    o[(i * C + i) * 2] = q[i];
                    o[(i * C + i) * 2 + 1] = 0.0;
                }
                size_t e = C;
                for (size_t i = 1; i < C; ++i)
                    for (size_t j = 0; j < i; ++j) {
                        o[(i * C + j) * 2] = q[e];
                        o[(i * C + j) * 2 + 1] = q[e + 1];
                        o[(j * C + i) * 2] = q[e];
                        o[(j * C + i) * 2 + 1] = -q[e + 1];
                        e += 2;
                    }
            }
            return BSS_OK;
        }
        case BSS_STATE_LATENT:
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "latent is exchanged as float64");
            return get(h, h->sZ, dst, B * N * K);
        case BSS_STATE_BASIS:
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "basis is exchanged as float64");
            return get(h, h->sT, dst, B * F * K);
        case BSS_STATE_ACTIVATION:
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "activation is exchanged as float64");
            return get(h, h->sV, dst, B * K * T);
        case BSS_STATE_ESTIMATION: return bss_separate(h, dst, dtype, 0);
    }
    return bss_fail(h, BSS_EINVAL, "unknown state");
}
