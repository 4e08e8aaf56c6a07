// FastMNMF orchestration: src/bss/mnmf.py:637-946 (FastMultichannelISNMF, non-partitioned).
// State: Q = h->W [B][F][M][M] (fp64) with its fp32 mirror h->Wf, g = h->G [B][N][F][M],
// W = h->basis [B][N][F][K], H = h->act [B][N][K][Tp].
#include "methods.h"

namespace {

template <typename T>
int dalloc(bss_handle* h, T** p, size_t n) {
    if (n == 0) n = 1;
    BSS_CUDA(h, cudaMalloc((void**)p, n * sizeof(T)));
    BSS_CUDA(h, cudaMemsetAsync(*p, 0, n * sizeof(T), h->stream));
    return BSS_OK;
}

}  // namespace

int mnmf_allocate(bss_handle* h) {
    if (h->cfg.partitioning) return bss_fail(h, BSS_EINVAL, "Not support partitioning function.");   // src/bss/mnmf.py:785,829
    const size_t B = h->B, M = h->C, N = h->N, F = h->F, Tp = h->Tp, K = h->K;
    BSS_TRY(dalloc(h, &h->X, B * F * M * Tp));
    BSS_TRY(dalloc(h, &h->W, B * F * M * M));
    BSS_TRY(dalloc(h, &h->Wf, B * F * M * M));
    BSS_TRY(dalloc(h, &h->U, B * M * F * M * M));
    BSS_TRY(dalloc(h, &h->Cx, B * F * M * M));
    BSS_TRY(dalloc(h, &h->gate, B * M * F));
    BSS_TRY(dalloc(h, &h->scale, B * (N > M ? N : M) * F * 2));
    BSS_TRY(dalloc(h, &h->logdet, B * F));
    BSS_TRY(dalloc(h, &h->lossbuf, B * F + B));
    BSS_TRY(dalloc(h, &h->G, B * N * F * M));
    BSS_TRY(dalloc(h, &h->G2, B * N * F * M));
    BSS_TRY(dalloc(h, &h->basis, B * N * F * K));
    BSS_TRY(dalloc(h, &h->basis2, B * N * F * K));
    BSS_TRY(dalloc(h, &h->act, B * N * K * Tp));
    BSS_TRY(dalloc(h, &h->iw, B * F * M * Tp));
    BSS_TRY(dalloc(h, &h->xt, B * F * ((M + 1) / 2 * 2) * Tp));   // x~ tiles have an even number of rows
    return BSS_OK;
}

// Q = I, g = 1e-2 with g[m % N, :, m] = 1      src/bss/mnmf.py:660-663
int mnmf_reset(bss_handle* h) {
    const size_t B = h->B, M = h->C, N = h->N, F = h->F;
    const size_t nq = B * F * M * M, ng = B * N * F * M;
    BSS_TRY(ensure_pinned(h, nq * sizeof(double2) + ng * sizeof(float)));
    double2* q = (double2*)h->pinned;
    float* g = (float*)(q + nq);
    for (size_t i = 0; i < nq; ++i) {
        const size_t rc = i % (M * M);
        q[i] = make_double2((rc / M) == (rc % M) ? 1.0 : 0.0, 0.0);
    }
    for (size_t i = 0; i < ng; ++i) {
        const size_t m = i % M;
        const size_t n = (i / (M * F)) % N;
        g[i] = (m % N) == n ? 1.f : 1e-2f;
    }
    BSS_CUDA(h, cudaMemcpyAsync(h->W, q, nq * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
    BSS_CUDA(h, cudaMemcpyAsync(h->G, g, ng * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    BSS_TRY(launch_sync_wf(h, h->W, h->Wf, (long long)nq));
    BSS_CUDA(h, bss_wait(h));
    h->has_filter = true;
    return BSS_OK;
}

int mnmf_update_once(bss_handle* h) {
    if (h->cfg.normalize != BSS_NORMALIZE_NONE && h->cfg.normalize != BSS_NORMALIZE_POWER)
        return bss_fail(h, BSS_EINVAL, "Not support normalization based on projection-back. Choose 'power'");   // mnmf.py:772-773
    // x~ = |Q x|^2 once per iteration: Q only changes in update_diagonalizer below
    BSS_TRY(launch_mnmf_xt(h));
    // update_NMF (mnmf.py:775-815)
    BSS_TRY(launch_mnmf_basis(h));
    BSS_TRY(launch_mnmf_act(h));
    // update_SCM (:817-846)
    BSS_TRY(launch_mnmf_scm(h));
    // update_diagonalizer (:848-888): R is fixed during the sweep over channels, so all M weighted
    // covariances come from one pass and the Gauss-Seidel sweep runs per bin in registers
    // eight channels: one kernel, inverse variances computed on the fly, a lane per Hermitian entry (kernels_cov8.cu) ...
    bool lane_per_entry = false;
    BSS_TRY(launch_covariance8(h, &lane_per_entry));
    // ... otherwise all M weighted covariances of a bin as one tensor-core contraction (kernels_cov_mma.cu) ...
    bool on_tensor_cores = lane_per_entry;
    if (!lane_per_entry) BSS_TRY(launch_mnmf_weights(h, 1));
    if (!lane_per_entry) BSS_TRY(launch_covariance_mma(h, h->X, h->iw, h->U, h->B, h->F, h->C, h->C, h->T, h->Tp, &on_tensor_cores));
    // ... or, for shapes it does not cover, the CUDA-core kernel with explicit weights
    if (!on_tensor_cores) BSS_TRY(launch_mnmf_weights(h, 0));
    CovArgs c{};
    c.X = h->X;
    c.U = h->U;
    c.B = h->B;
    c.F = h->F;
    c.C = h->C;
    c.NW = h->C;
    c.T = h->T;
    c.Tp = h->Tp;
    c.wmode = WM_EXPLICIT;
    c.iw = h->iw;
    c.n_sel = h->C;
    for (int i = 0; i < 8; ++i) c.wsel[i] = i;
    if (!on_tensor_cores) BSS_TRY(launch_covariance(h, c));
    IpArgs ip{};
    ip.W = h->W;
    ip.Wf = h->Wf;
    ip.U = h->U;
    ip.gate = h->gate;
    ip.flags = h->flags;
    ip.B = h->B;
    ip.F = h->F;
    ip.C = h->C;
    ip.threshold = h->cfg.threshold;
    ip.eps = h->cfg.eps;
    ip.use_gate = 1;
    ip.floor_den = 1;   // mnmf.py:882-883
    ip.pair_m = ip.pair_n = -1;
    BSS_TRY(launch_ip(h, ip));
    if (h->cfg.normalize == BSS_NORMALIZE_POWER) BSS_TRY(launch_mnmf_normalize(h));
    return BSS_OK;
}

// loss = sum((x~+eps)/(y~+eps) + log(y~+eps)) - T sum_f log|det(Q Q^T)|      mnmf.py:890-917
int mnmf_loss(bss_handle* h) {
    const size_t BF = (size_t)h->B * h->F;
    double* result = h->lossbuf + BF;
    BSS_CUDA(h, cudaMemsetAsync(result, 0, sizeof(double) * h->B, h->stream));
    BSS_TRY(launch_logdet(h, h->W, h->logdet, (long long)BF, h->C, 1));
    BSS_TRY(launch_mnmf_xt(h));
    BSS_TRY(launch_mnmf_loss_terms(h));
    return launch_loss_finish(h, h->lossbuf, h->logdet, (double)h->T, h->B, h->F, result);
}

// the covariance-accumulate step of update_diagonalizer alone (bss_time_covariance): inverse weights from (W, H, g), then
// all M weighted covariances of every bin
int mnmf_covariance_only(bss_handle* h) {
    bool lane_per_entry = false;
    BSS_TRY(launch_covariance8(h, &lane_per_entry));
    if (lane_per_entry) return BSS_OK;
    BSS_TRY(launch_mnmf_weights(h, 1));
    bool on_tensor_cores = false;
    BSS_TRY(launch_covariance_mma(h, h->X, h->iw, h->U, h->B, h->F, h->C, h->C, h->T, h->Tp, &on_tensor_cores));
    if (on_tensor_cores) return BSS_OK;
    BSS_TRY(launch_mnmf_weights(h, 0));
    CovArgs c{};
    c.X = h->X;
    c.U = h->U;
    c.B = h->B;
    c.F = h->F;
    c.C = h->C;
    c.NW = h->C;
    c.T = h->T;
    c.Tp = h->Tp;
    c.wmode = WM_EXPLICIT;
    c.iw = h->iw;
    c.n_sel = h->C;
    for (int i = 0; i < 8; ++i) c.wsel[i] = i;
    return launch_covariance(h, c);
}

int mnmf_separate(bss_handle* h, cf* out) { return launch_mnmf_separate(h, out); }
