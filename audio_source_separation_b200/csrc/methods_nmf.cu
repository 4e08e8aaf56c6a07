// Single-channel NMF multiplicative updates: src/algorithm/nmf.py:150-595.
#include "methods.h"

namespace {

// exponents of `T <- T * (num / den)^q` per divergence and domain d
int nmf_math(bss_handle* h, NmfMath* out) {
    NmfMath m{};
    const double d = h->cfg.domain;
    m.alg = h->cfg.algorithm;
    m.nu = h->cfg.nu;
    m.eps = h->cfg.eps;
    m.loss_expo = 2.0 / d;
    m.loss_eps = h->cfg.eps;
    switch (h->cfg.method) {
        case BSS_NMF_EUC:   // src/algorithm/nmf.py:182-207
            if (m.alg != BSS_ALG_MM) return bss_fail(h, BSS_EINVAL, "Not support this update for EUC-NMF.");
            m.kind = 0;
            m.a = (4.0 - d) / d;
            m.b = (2.0 - d) / d;
            m.q = d / (4.0 - d);
            break;
        case BSS_NMF_KL:    // :241-266
            if (m.alg != BSS_ALG_MM) return bss_fail(h, BSS_EINVAL, "Not support this update for KL-NMF.");
            m.kind = 1;
            m.b = (2.0 - d) / d;
            m.q = d / 2.0;
            m.loss_eps = 1e-12;   // src/criterion/divergence.py:3 (module constant, not the model's eps)
            break;
        case BSS_NMF_IS:    // :302-356
            m.kind = 2;
            m.p = (d + 2.0) / d;
            if (m.alg == BSS_ALG_MM)
                m.q = d / (d + 2.0);
            else if (m.alg == BSS_ALG_ME) {
                if (d != 2.0) return bss_fail(h, BSS_EINVAL, "Only domain = 2 is supported.");
                m.q = 1.0;
            } else
                return bss_fail(h, BSS_EINVAL, "Not support this update for IS-NMF.");
            m.loss_eps = 1e-12;
            break;
        case BSS_NMF_T:     // :397-428
            if (m.alg != BSS_ALG_MM) return bss_fail(h, BSS_EINVAL, "Not support this update for t-NMF.");
            if (d != 2.0) return bss_fail(h, BSS_EINVAL, "`domain` is expected 2.");
            m.kind = 3;
            m.q = 0.5;
            break;
        case BSS_NMF_CAUCHY:   // :461-595
            if (d != 2.0) return bss_fail(h, BSS_EINVAL, "Only 'domain' = 2 is supported.");
            if (m.alg < BSS_ALG_MM || m.alg > BSS_ALG_MM_FAST) return bss_fail(h, BSS_EINVAL, "Not support this update for Cauchy-NMF.");
            m.kind = 4;
            break;
        default: return bss_fail(h, BSS_EINVAL, "unknown NMF method");
    }
    *out = m;
    return BSS_OK;
}

template <typename T>
int dalloc(bss_handle* h, T** p, size_t n) {
    if (n == 0) n = 1;
    BSS_CUDA(h, cudaMalloc((void**)p, n * sizeof(T)));
    BSS_CUDA(h, cudaMemsetAsync(*p, 0, n * sizeof(T), h->stream));
    return BSS_OK;
}

}  // namespace

int nmf_allocate(bss_handle* h) {
    const size_t B = h->B, F = h->F, T = h->T, K = h->K;
    NmfMath m;
    BSS_TRY(nmf_math(h, &m));   // reject unsupported combinations at creation
    BSS_TRY(dalloc(h, &h->nz, B * F * T));
    BSS_TRY(dalloc(h, &h->nt, B * F * K));
    BSS_TRY(dalloc(h, &h->nv, B * K * T));
    BSS_TRY(dalloc(h, &h->lossbuf, B * F + B));
    return BSS_OK;
}

int nmf_update_once(bss_handle* h) { return nmf_run(h, 1, nullptr); }

// n_iter updates; with loss_hist (device, [n_iter][B]) the criterion after every update is recorded as well
// (NMFbase.update, src/algorithm/nmf.py:165-174).  Small problems run as ONE cluster launch for the whole loop.
int nmf_run(bss_handle* h, int n_iter, double* loss_hist) {
    if (!h->has_input) return bss_fail(h, BSS_ESTATE, "Specify data!");
    NmfMath m;
    BSS_TRY(nmf_math(h, &m));
    bool done = false;
    BSS_TRY(launch_nmf_fused(h, m, h->nz, h->nt, h->nv, loss_hist, h->B, h->F, h->T, h->K, n_iter, &done));
    if (done) return BSS_OK;
    for (int i = 0; i < n_iter; ++i) {
        BSS_TRY(launch_nmf_update(h, m, h->nz, h->nt, h->nv, h->B, h->F, h->T, h->K));
        if (loss_hist) {
            BSS_TRY(nmf_loss(h));
            BSS_CUDA(h, cudaMemcpyAsync(loss_hist + (size_t)i * h->B, h->lossbuf + (size_t)h->B * h->F, sizeof(double) * h->B,
                                        cudaMemcpyDeviceToDevice, h->stream));
        }
    }
    return BSS_OK;
}

int nmf_loss(bss_handle* h) {
    if (!h->has_input) return bss_fail(h, BSS_ESTATE, "Specify data!");
    NmfMath m;
    BSS_TRY(nmf_math(h, &m));
    double* result = h->lossbuf + (size_t)h->B * h->F;
    BSS_CUDA(h, cudaMemsetAsync(result, 0, sizeof(double) * h->B, h->stream));
    BSS_TRY(launch_nmf_loss(h, m, h->nz, h->nt, h->nv, h->lossbuf, h->B, h->F, h->T, h->K));
    return launch_loss_finish(h, h->lossbuf, nullptr, 0.0, h->B, h->F, result);
}
