// GaussILRMA with a shared (partitioned) basis: src/bss/ilrma.py:368-408 (source model), :313-320
// (normalisation), :493-495 / :543-548 (variance used by the spatial update), :664-668 (loss).
#include "methods.h"

namespace {

bool iss(const bss_handle* h) { return h->cfg.spatial == BSS_SPATIAL_ISS; }

// source-model arguments on the effective per-source factors
MuArgs eff_args(bss_handle* h) {
    MuArgs m{};
    m.X = h->X;
    m.Y = iss(h) ? h->Y : nullptr;
    m.Wf = h->Wf;
    m.basis = h->beff;
    m.basis_out = nullptr;
    m.act = h->aeff;
    m.B = h->B;
    m.F = h->F;
    m.C = h->C;
    m.T = h->T;
    m.Tp = h->Tp;
    m.K = h->K;
    m.mode = 0;
    m.p_exp = 2.f;
    m.q_exp = 0.5f;
    m.eps = (float)h->cfg.eps;
    m.sel_m = m.sel_n = -1;
    m.raw = h->praw;
    return m;
}

}  // namespace

int ilrma_partitioned_update_once(bss_handle* h) {
    if (h->cfg.normalize == BSS_NORMALIZE_PROJECTION_BACK)
        return bss_fail(h, BSS_EUNSUPPORTED,
                        "Not support 'projection-back' based normalization for partitioninig function. Choose 'power' based normalization.");
    if (h->cfg.domain != 2.0) return bss_fail(h, BSS_EINVAL, "Not support domain = " + std::to_string(h->cfg.domain));
    if (h->cfg.spatial == BSS_SPATIAL_IP2) return bss_fail(h, BSS_EUNSUPPORTED, "Not support partitioning function.");
    const MuArgs m = eff_args(h);
    // latent Z (assigned, not multiplied into the old Z: ilrma.py:382), then Z /= sum_n Z
    BSS_TRY(launch_part_expand(h));
    BSS_TRY(launch_mu_basis(h, m));
    BSS_TRY(launch_part_latent(h));
    // shared basis T
    BSS_TRY(launch_part_expand(h));
    BSS_TRY(launch_mu_basis(h, m));
    BSS_TRY(launch_part_basis(h));
    // shared activation V
    BSS_TRY(launch_part_expand(h));
    int n_chunks = 0;
    BSS_TRY(launch_mu_act(h, m, nullptr, &n_chunks));
    BSS_TRY(launch_part_act_finish(h, n_chunks));
    // spatial model on R = sum_k Z T V
    BSS_TRY(launch_part_expand(h));
    const bool power = h->cfg.normalize == BSS_NORMALIZE_POWER;
    if (iss(h)) {
        BSS_TRY(launch_iss(h, h->Y, 0, h->beff, h->aeff, nullptr, power ? h->pw : nullptr, h->B, h->N, h->F, h->T, h->Tp, h->K, 1.f,
                           (float)h->cfg.eps));
        h->has_filter = false;
        if (power) {
            BSS_TRY(launch_aux_from_power(h, h->pw, h->aux, h->B, h->N, h->F, h->cfg.eps));
            BSS_TRY(launch_scale_y(h, h->Y, nullptr, h->aux, nullptr, h->B, h->N, h->F, h->Tp, h->K, 2.0));
            BSS_TRY(launch_part_normalize(h));
        }
        return BSS_OK;
    }
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
    c.basis = h->beff;
    c.act = h->aeff;
    c.wmode = WM_ILRMA;
    c.expo = 1.f;
    c.n_sel = h->N;
    for (int i = 0; i < 8; ++i) c.wsel[i] = i;
    BSS_TRY(launch_covariance(h, c));
    IpArgs ip{};
    ip.W = h->W;
    ip.Wf = h->Wf;
    ip.U = h->U;
    ip.Cx = h->Cx;
    ip.gate = h->gate;
    ip.pw = power ? h->pw : nullptr;
    ip.flags = h->flags;
    ip.B = h->B;
    ip.F = h->F;
    ip.C = h->C;
    ip.threshold = h->cfg.threshold;
    ip.eps = h->cfg.eps;
    ip.use_gate = 1;
    ip.floor_den = 0;
    ip.pair_m = ip.pair_n = -1;
    BSS_TRY(launch_ip(h, ip));
    h->y_valid = false;
    if (power) {
        BSS_TRY(launch_normalize_power(h, h->W, h->Wf, nullptr, h->pw, h->B, h->N, h->C, h->F, h->K, 2.0, h->cfg.eps, h->aux));
        BSS_TRY(launch_part_normalize(h));
    }
    return BSS_OK;
}

// called by bss_loss_device after log|det W| is in h->logdet and the result slot is zeroed
int ilrma_partitioned_loss(bss_handle* h) {
    BSS_TRY(launch_part_expand(h));
    MuArgs m = eff_args(h);
    m.raw = nullptr;
    BSS_TRY(launch_ilrma_loss(h, m, (float)(2.0 / h->cfg.domain), h->lossbuf));
    return launch_loss_finish(h, h->lossbuf, h->logdet, 2.0 * (double)h->T, h->B, h->F, h->lossbuf + (size_t)h->B * h->F);
}
