// update_once / loss / separate of the determined methods (Gauss-ILRMA, t-ILRMA, AuxIVA):
// which kernels run, in which order, on which buffers.  All work is queued on h->stream; nothing
// here synchronises with the host.
#include <cstdlib>

#include "methods.h"

namespace {

bool uses_model(const bss_handle* h) { return h->cfg.method == BSS_GAUSS_ILRMA || h->cfg.method == BSS_T_ILRMA; }
bool is_iss(const bss_handle* h) { return h->cfg.spatial == BSS_SPATIAL_ISS; }

template <typename T>
int dalloc(bss_handle* h, T** p, size_t n) {
    if (n == 0) n = 1;
    BSS_CUDA(h, cudaMalloc((void**)p, n * sizeof(T)));
    BSS_CUDA(h, cudaMemsetAsync(*p, 0, n * sizeof(T), h->stream));
    return BSS_OK;
}

MuArgs mu_args(bss_handle* h) {
    MuArgs m{};
    m.X = h->X;
    m.Y = is_iss(h) ? h->Y : nullptr;
    m.Wf = h->Wf;
    m.basis = h->basis;
    m.basis_out = h->basis2;
    m.act = h->act;
    m.B = h->B;
    m.F = h->F;
    m.C = h->C;
    m.T = h->T;
    m.Tp = h->Tp;
    m.K = h->K;
    m.eps = (float)h->cfg.eps;
    m.sel_m = m.sel_n = -1;
    if (h->cfg.method == BSS_T_ILRMA) {
        m.mode = 1;
        m.nu = (float)h->cfg.nu;
        m.p_exp = 2.f;
        m.q_exp = 0.5f;
    } else {
        const double d = h->cfg.domain;
        m.mode = 0;
        m.p_exp = (float)((d + 2.0) / d);
        m.q_exp = (float)(d / (d + 2.0));
    }
    return m;
}

CovArgs cov_args(bss_handle* h) {
    CovArgs c{};
    c.X = h->X;
    c.U = h->U;
    c.B = h->B;
    c.F = h->F;
    c.C = h->C;
    c.NW = h->N;
    c.T = h->T;
    c.Tp = h->Tp;
    c.K = h->K;
    c.eps = (float)h->cfg.eps;
    c.basis = h->basis;
    c.act = h->act;
    c.wfr = h->wfr;
    c.iw = h->iw;
    c.expo = 1.f;
    c.n_sel = h->N;
    for (int i = 0; i < 8; ++i) c.wsel[i] = i;
    return c;
}

IpArgs ip_args(bss_handle* h, bool want_power) {
    IpArgs a{};
    a.W = h->W;
    a.Wf = h->Wf;
    a.U = h->U;
    a.Cx = h->Cx;
    a.gate = h->gate;
    a.pw = want_power ? h->pw : nullptr;
    a.flags = h->flags;
    a.order = nullptr;
    a.eigval = nullptr;
    a.variant = h->opt_ip_kernel;
    a.B = h->B;
    a.F = h->F;
    a.C = h->C;
    a.threshold = h->cfg.threshold;
    a.eps = h->cfg.eps;
    a.use_gate = 1;
    a.floor_den = 0;
    a.pair_m = a.pair_n = -1;
    return a;
}

// The basis update computes |y|^2 = |W x|^2 of every frame anyway; with n_basis == 2 (the compile-time-K kernels) it hands
// them to the activation update as float tiles (half the bytes of X, no second demixing).  Same arithmetic, same values:
// results are bit-identical to the two-pass form (BSSGPU_NO_POWER_HANDOFF=1 selects it, for A/B measurements).
bool power_handoff(bss_handle* h) {
    static const bool disabled = getenv("BSSGPU_NO_POWER_HANDOFF") != nullptr;
    if (disabled || is_iss(h) || h->K != 2) return false;
    if (((size_t)h->N * h->Tp) % 4 != 0) return false;   // 16-byte bulk copies of float rows
    if (!h->P) {
        const size_t n_blocks = ((size_t)h->Tp + BSS_XSLAB - 1) / BSS_XSLAB;   // block-major: [B][block][F][N][128]
        if (cudaMalloc((void**)&h->P, (size_t)h->B * n_blocks * h->F * h->N * BSS_XSLAB * sizeof(float)) != cudaSuccess) {
            cudaGetLastError();
            h->P = nullptr;
            return false;   // not enough memory for the hand-off buffer: stream X twice
        }
    }
    return true;
}

int source_model(bss_handle* h, int sel_m, int sel_n) {
    MuArgs m = mu_args(h);
    m.sel_m = sel_m;
    m.sel_n = sel_n;
    // one pass over X for both factor updates where the fused kernel covers the configuration (kernels_mu_fused.cu);
    // BSS_OPT_SOURCE_MODEL = 1 keeps the three-pass form (basis kernel -> power tiles -> activation kernel)
    if (h->opt_source_model != 1 && !is_iss(h)) {
        bool fused = false;
        int n_chunks = 0;
        BSS_TRY(launch_mu_fused(h, m, &n_chunks, &fused));
        if (fused) {
            float* t = h->basis;
            h->basis = h->basis2;
            h->basis2 = t;
            m.basis = h->basis;
            m.basis_out = h->basis2;
            h->last_source_model = 2;
            return launch_mu_act_finish(h, m, h->act, n_chunks);
        }
    }
    h->last_source_model = 1;
    const bool handoff = power_handoff(h);
    if (handoff) m.Pout = h->P;
    BSS_TRY(launch_mu_basis(h, m));
    float* t = h->basis;
    h->basis = h->basis2;
    h->basis2 = t;
    m.basis = h->basis;
    m.basis_out = h->basis2;
    m.Pout = nullptr;
    if (handoff) m.Pin = h->P;
    return launch_mu_act(h, m, h->act);
}

// `normalize` tail of update_once for the filter-based (IP / IP2) updates
int normalize_filter(bss_handle* h, double domain) {
    if (h->cfg.normalize == BSS_NORMALIZE_POWER)
        return launch_normalize_power(h, h->W, h->Wf, uses_model(h) ? h->basis : nullptr, h->pw, h->B, h->N, h->C, h->F, h->K,
                                      domain, h->cfg.eps, h->aux);
    if (h->cfg.normalize == BSS_NORMALIZE_PROJECTION_BACK) {
        BSS_TRY(launch_pb_scale(h, h->W, h->Cx, (double2*)h->scale, h->B, h->F, h->C, h->cfg.reference_id));
        return launch_normalize_pb(h, h->W, h->Wf, uses_model(h) ? h->basis : nullptr, (const double2*)h->scale, h->B, h->N,
                                   h->C, h->F, h->K, domain);
    }
    return BSS_OK;
}

// `normalize` tail for ISS, which carries Y instead of W
int normalize_estimates(bss_handle* h, double domain) {
    if (h->cfg.normalize == BSS_NORMALIZE_POWER) {
        BSS_TRY(launch_aux_from_power(h, h->pw, h->aux, h->B, h->N, h->F, h->cfg.eps));
        return launch_scale_y(h, h->Y, h->basis, h->aux, nullptr, h->B, h->N, h->F, h->Tp, h->K, domain);
    }
    if (h->cfg.normalize == BSS_NORMALIZE_PROJECTION_BACK) {
        BSS_TRY(bss_filter_from_estimates(h));
        BSS_TRY(launch_pb_scale(h, h->W, h->Cx, (double2*)h->scale, h->B, h->F, h->C, h->cfg.reference_id));
        return launch_scale_y(h, h->Y, h->basis, nullptr, (const double2*)h->scale, h->B, h->N, h->F, h->Tp, h->K, domain);
    }
    return BSS_OK;
}

int need_pair(bss_handle* h) {
    if (h->pair_m < 0 || h->pair_n < 0) return bss_fail(h, BSS_ESTATE, "IP2: no update pair selected");
    return BSS_OK;
}

}  // namespace

int bss_allocate(bss_handle* h) {
    const size_t B = h->B, C = h->C, N = h->N, F = h->F, Tp = h->Tp, K = h->K;
    BSS_TRY(dalloc(h, &h->X, B * F * C * Tp));
    BSS_TRY(dalloc(h, &h->W, B * F * N * C));
    BSS_TRY(dalloc(h, &h->Wf, B * F * N * C));
    BSS_TRY(dalloc(h, &h->U, B * N * F * C * C));
    BSS_TRY(dalloc(h, &h->Cx, B * F * C * C));
    BSS_TRY(dalloc(h, &h->gate, B * N * F));
    BSS_TRY(dalloc(h, &h->pw, B * N * F));
    BSS_TRY(dalloc(h, &h->scale, B * N * F * 2));
    BSS_TRY(dalloc(h, &h->logdet, B * F));
    BSS_TRY(dalloc(h, &h->aux, B * N));
    BSS_TRY(dalloc(h, &h->lossbuf, B * F + B));
    BSS_TRY(dalloc(h, &h->order, B * F * 2));
    BSS_TRY(dalloc(h, &h->eigval, B * F * 2));
    if (uses_model(h)) {
        if (h->cfg.partitioning) {
            BSS_TRY(dalloc(h, &h->basis, B * F * K));
            BSS_TRY(dalloc(h, &h->basis2, B * F * K));
            BSS_TRY(dalloc(h, &h->act, B * K * Tp));
            BSS_TRY(dalloc(h, &h->latent, B * N * K));
            BSS_TRY(dalloc(h, &h->latent2, B * N * K));
            BSS_TRY(dalloc(h, &h->beff, B * N * F * K));
            BSS_TRY(dalloc(h, &h->aeff, B * N * K * Tp));
            BSS_TRY(dalloc(h, &h->praw, B * N * F * K * 2));
        } else {
            BSS_TRY(dalloc(h, &h->basis, B * N * F * K));
            BSS_TRY(dalloc(h, &h->basis2, B * N * F * K));
            BSS_TRY(dalloc(h, &h->act, B * N * K * Tp));
        }
        if (h->cfg.method == BSS_T_ILRMA && !h->iw) BSS_TRY(dalloc(h, &h->iw, B * F * N * Tp));
    } else {
        BSS_TRY(dalloc(h, &h->wfr, B * N * Tp));
        BSS_TRY(dalloc(h, &h->wraw, B * N * Tp));
        if (h->cfg.method == BSS_GAUSS_IDLMA && !h->iw) BSS_TRY(dalloc(h, &h->iw, B * F * N * Tp));
    }
    if (is_iss(h)) {
        BSS_TRY(dalloc(h, &h->Y, B * F * N * Tp));
        BSS_TRY(dalloc(h, &h->G2x, B * F * N * C));
    }
    return BSS_OK;
}

// W = I for every bin (src/bss/ilrma.py:67-69); for ISS the estimates start as Y = X
int bss_reset_filter(bss_handle* h) {
    BSS_TRY(launch_identity_filter(h, h->W, h->Wf, (long long)h->B * h->F, h->N, h->C));
    h->has_filter = true;
    h->pair_m = h->pair_n = -1;
    if (is_iss(h) && h->has_input) BSS_TRY(bss_refresh_estimates(h));
    return BSS_OK;
}

int bss_refresh_estimates(bss_handle* h) {
    if (!h->Y) BSS_TRY(dalloc(h, &h->Y, (size_t)h->B * h->F * h->N * h->Tp));
    BSS_TRY(launch_separate(h, h->X, h->Wf, nullptr, h->Y, nullptr, h->B, h->C, h->F, h->T, h->Tp));
    h->y_valid = true;
    return BSS_OK;
}

int bss_filter_from_estimates(bss_handle* h) {
    if (!h->Y || !h->y_valid) return bss_fail(h, BSS_ESTATE, "no estimates to derive a filter from");
    if (!h->G2x) BSS_TRY(dalloc(h, &h->G2x, (size_t)h->B * h->F * h->N * h->C));
    BSS_TRY(launch_cross_cov(h, h->Y, h->X, h->G2x, (long long)h->B * h->F, h->C, h->T, h->Tp));
    BSS_TRY(launch_lsq_filter(h, h->G2x, h->Cx, h->W, (long long)h->B * h->F, h->C));
    return launch_sync_wf(h, h->W, h->Wf, (long long)h->B * h->F * h->N * h->C);
}

int bss_covariance_only(bss_handle* h) {
    if (h->cfg.method == BSS_FAST_MNMF) return mnmf_covariance_only(h);
    CovArgs c = cov_args(h);
    if (h->cfg.method == BSS_GAUSS_ILRMA && h->cfg.partitioning) {
        BSS_TRY(launch_part_expand(h));
        c.basis = h->beff;
        c.act = h->aeff;
        c.wmode = WM_ILRMA;
    } else if (h->cfg.method == BSS_GAUSS_ILRMA) {
        c.wmode = WM_ILRMA;
        c.expo = (float)(2.0 / h->cfg.domain);
    } else if (uses_model(h)) {
        c.wmode = WM_EXPLICIT;
    } else {
        c.wmode = WM_FRAME;
    }
    return launch_covariance(h, c);
}

int ilrma_update_once(bss_handle* h) {
    if (h->cfg.partitioning) return ilrma_partitioned_update_once(h);
    const double d = h->cfg.domain;
    const int sp = h->cfg.spatial;
    if (sp == BSS_SPATIAL_IP2) BSS_TRY(need_pair(h));
    BSS_TRY(source_model(h, sp == BSS_SPATIAL_IP2 ? h->pair_m : -1, sp == BSS_SPATIAL_IP2 ? h->pair_n : -1));
    if (sp == BSS_SPATIAL_ISS) {
        BSS_TRY(launch_iss(h, h->Y, 0, h->basis, h->act, nullptr, h->cfg.normalize == BSS_NORMALIZE_POWER ? h->pw : nullptr, h->B,
                           h->N, h->F, h->T, h->Tp, h->K, (float)(2.0 / d), (float)h->cfg.eps));
        h->has_filter = false;
        return normalize_estimates(h, d);
    }
    CovArgs c = cov_args(h);
    c.wmode = WM_ILRMA;
    c.expo = (float)(2.0 / d);
    IpArgs ip = ip_args(h, h->cfg.normalize == BSS_NORMALIZE_POWER);
    if (sp == BSS_SPATIAL_IP2) {
        c.n_sel = 2;
        c.wsel[0] = h->pair_m;
        c.wsel[1] = h->pair_n;
        ip.pair_m = h->pair_m;
        ip.pair_n = h->pair_n;
        ip.order = h->order;
        ip.eigval = h->eigval;
    }
    BSS_TRY(launch_covariance(h, c));
    BSS_TRY(launch_ip(h, ip));
    h->y_valid = false;
    return normalize_filter(h, d);
}

int tilrma_update_once(bss_handle* h) {
    if (h->cfg.partitioning) return bss_fail(h, BSS_EUNSUPPORTED, "tILRMA with partitioning is not supported");
    if (h->cfg.normalize == BSS_NORMALIZE_PROJECTION_BACK)
        return bss_fail(h, BSS_EINVAL, "Not support normalization based on projection-back. Choose 'power' or 'projection-back'");
    BSS_TRY(source_model(h, -1, -1));
    BSS_TRY(launch_t_weights(h, h->X, h->Wf, h->basis, h->act, h->iw, h->B, h->C, h->F, h->K, h->Tp, (float)h->cfg.nu,
                             (float)h->cfg.eps));
    CovArgs c = cov_args(h);
    c.wmode = WM_EXPLICIT;
    BSS_TRY(launch_covariance(h, c));
    IpArgs ip = ip_args(h, h->cfg.normalize == BSS_NORMALIZE_POWER);
    ip.use_gate = 0;    // plain inverse, src/bss/ilrma.py:975
    ip.floor_den = 1;   // src/bss/ilrma.py:981
    BSS_TRY(launch_ip(h, ip));
    h->y_valid = false;
    return normalize_filter(h, 2.0);   // src/bss/ilrma.py:849: exponent 2
}

int auxiva_update_once(bss_handle* h) {
    const int kind = h->cfg.method == BSS_AUX_LAPLACE_IVA ? 0 : 1;
    const int sp = h->cfg.spatial;
    if (sp == BSS_SPATIAL_ISS) {
        BSS_TRY(launch_frame_weights(h, h->Y, nullptr, 1, h->wfr, nullptr, h->B, h->C, h->F, h->T, h->Tp, kind, (float)h->cfg.eps));
        h->has_filter = false;
        return launch_iss(h, h->Y, 1, nullptr, nullptr, h->wfr, nullptr, h->B, h->N, h->F, h->T, h->Tp, 0, 1.f, (float)h->cfg.eps);
    }
    if (sp == BSS_SPATIAL_IP2) {
        if (kind == 1) return bss_fail(h, BSS_EUNSUPPORTED, "In progress...");   // src/bss/iva.py:777-778
        BSS_TRY(need_pair(h));
    }
    BSS_TRY(launch_frame_weights(h, h->X, h->Wf, 0, h->wfr, nullptr, h->B, h->C, h->F, h->T, h->Tp, kind, (float)h->cfg.eps));
    CovArgs c = cov_args(h);
    c.wmode = WM_FRAME;
    IpArgs ip = ip_args(h, false);
    if (sp == BSS_SPATIAL_IP2) {
        c.n_sel = 2;
        c.wsel[0] = h->pair_m;
        c.wsel[1] = h->pair_n;
        ip.pair_m = h->pair_m;
        ip.pair_n = h->pair_n;
        ip.order = h->order;
        ip.eigval = h->eigval;
    }
    BSS_TRY(launch_covariance(h, c));
    BSS_TRY(launch_ip(h, ip));
    h->y_valid = false;
    return BSS_OK;
}

// GaussIDLMA (src/sss/idlma.py:142-210): the source variances are whatever the caller's DNN produced
// (BSS_STATE_VARIANCE); the handle owns the spatial half: U[n,f] = mean_t x x^H / R[n,f,t], the gated IP sweep,
// and (bss_normalize) the projection-back normalisation (W <- diag(scale) W, what :155-158 computes by least squares)
int idlma_set_variance(bss_handle* h, const double* r) {
    const size_t n = (size_t)h->B * h->N * h->F * h->T;
    BSS_TRY(ensure_staging(h, n * sizeof(double)));
    BSS_CUDA(h, cudaMemcpyAsync(h->staging, r, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    BSS_TRY(launch_import_variance(h, (const double*)h->staging, h->iw, h->B, h->N, h->F, h->T, h->Tp, h->cfg.eps));
    BSS_CUDA(h, bss_wait(h));   // the host buffer is borrowed for this call only
    h->has_variance = true;
    return BSS_OK;
}

int idlma_update_once(bss_handle* h) {
    if (!h->has_variance) return bss_fail(h, BSS_ESTATE, "GaussIDLMA: set the source variances (dnn_output) first");
    if (!h->has_filter) return bss_fail(h, BSS_ESTATE, "GaussIDLMA: no demixing filter");
    CovArgs c = cov_args(h);
    c.wmode = WM_EXPLICIT;
    BSS_TRY(launch_covariance(h, c));
    BSS_TRY(launch_ip(h, ip_args(h, false)));   // gate on, no denominator floor: src/sss/idlma.py:200-207
    h->y_valid = false;
    return BSS_OK;
}

// src/sss/idlma.py:150-165; 'power' / False raise in the reference: the wrapper raises before calling
int idlma_normalize(bss_handle* h) {
    if (h->cfg.normalize != BSS_NORMALIZE_PROJECTION_BACK)
        return bss_fail(h, BSS_EINVAL, "Not support normalization based on power. Choose 'power' or 'projection-back'");
    return normalize_filter(h, 2.0);
}

int bss_loss_device(bss_handle* h) {
    const size_t BF = (size_t)h->B * h->F;
    double* result = h->lossbuf + BF;
    BSS_CUDA(h, cudaMemsetAsync(result, 0, sizeof(double) * h->B, h->stream));
    if (is_iss(h)) BSS_TRY(bss_filter_from_estimates(h));
    BSS_TRY(launch_logdet(h, h->W, h->logdet, (long long)BF, h->C, 0));
    const double coef = 2.0 * (double)h->T;
    if (uses_model(h)) {
        if (h->cfg.partitioning) return ilrma_partitioned_loss(h);
        MuArgs m = mu_args(h);
        const float expo = h->cfg.method == BSS_T_ILRMA ? 1.f : (float)(2.0 / h->cfg.domain);
        BSS_TRY(launch_ilrma_loss(h, m, expo, h->lossbuf));
        return launch_loss_finish(h, h->lossbuf, h->logdet, coef, h->B, h->F, result);
    }
    if (h->cfg.method == BSS_GAUSS_IDLMA) {
        if (!h->has_variance) return bss_fail(h, BSS_ESTATE, "GaussIDLMA: set the source variances (dnn_output) first");
        BSS_TRY(launch_idlma_loss(h, h->X, h->Wf, h->iw, h->lossbuf, h->B, h->C, h->F, h->T, h->Tp));
        return launch_loss_finish(h, h->lossbuf, h->logdet, coef, h->B, h->F, result);
    }
    const int kind = h->cfg.method == BSS_AUX_LAPLACE_IVA ? 0 : 1;
    // Laplace: the reference takes |Y| of the stored estimates for ISS and of W X otherwise; Gauss always
    // recomputes W X (src/bss/iva.py:796)
    const bool from_y = is_iss(h) && kind == 0;
    BSS_TRY(launch_frame_weights(h, from_y ? h->Y : h->X, h->Wf, from_y ? 1 : 0, nullptr, h->wraw, h->B, h->C, h->F, h->T, h->Tp,
                                 kind, (float)h->cfg.eps));
    BSS_TRY(launch_sum_frames(h, h->wraw, h->B, h->N, h->T, h->Tp, kind, kind == 0 ? 2.0 : (double)h->F, h->cfg.eps, result));
    return launch_loss_finish(h, nullptr, h->logdet, coef, h->B, h->F, result);
}

int bss_separate_to(bss_handle* h, cf* out, int apply_pb) {
    if (is_iss(h)) {
        if (!h->y_valid) return bss_fail(h, BSS_ESTATE, "no estimates");
        const double2* scale = nullptr;
        if (apply_pb) {
            BSS_TRY(bss_filter_from_estimates(h));
            BSS_TRY(launch_pb_scale(h, h->W, h->Cx, (double2*)h->scale, h->B, h->F, h->C, h->cfg.reference_id));
            scale = (const double2*)h->scale;
        }
        return launch_export_y(h, h->Y, scale, out, h->B, h->N, h->F, h->T, h->Tp);
    }
    const double2* scale = nullptr;
    if (apply_pb) {
        BSS_TRY(launch_pb_scale(h, h->W, h->Cx, (double2*)h->scale, h->B, h->F, h->C, h->cfg.reference_id));
        scale = (const double2*)h->scale;
    }
    return launch_separate(h, h->X, h->Wf, scale, nullptr, out, h->B, h->C, h->F, h->T, h->Tp);
}
