// Host <-> device movement of the model state (reference layouts on the host side) and the
// stateless primitive entry points used by the parity tests.
#include <cstring>
#include <vector>

#include "handle.h"
#include "methods.h"

namespace {

bool is_nmf(int m) { return m >= BSS_NMF_EUC && m <= BSS_NMF_CAUCHY; }

// rows x T float64 on the host  <->  rows x Tp float32 on the device (pad zeroed)
int put_rows(bss_handle* h, float* dev, const double* src, size_t rows, int T, int Tp) {
    if (!dev) return bss_fail(h, BSS_EINVAL, "this model has no such state");
    BSS_TRY(ensure_pinned(h, rows * Tp * sizeof(float)));
    float* p = (float*)h->pinned;
    for (size_t r = 0; r < rows; ++r) {
        for (int t = 0; t < T; ++t) p[r * Tp + t] = (float)src[r * T + t];
        for (int t = T; t < Tp; ++t) p[r * Tp + t] = 0.f;
    }
    BSS_CUDA(h, cudaMemcpyAsync(dev, p, rows * Tp * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    BSS_CUDA(h, bss_wait(h));
    return BSS_OK;
}

int get_rows(bss_handle* h, const float* dev, double* dst, size_t rows, int T, int Tp) {
    if (!dev) return bss_fail(h, BSS_EINVAL, "this model has no such state");
    BSS_TRY(ensure_pinned(h, rows * Tp * sizeof(float)));
    float* p = (float*)h->pinned;
    BSS_CUDA(h, cudaMemcpyAsync(p, dev, rows * Tp * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    BSS_CUDA(h, bss_wait(h));
    for (size_t r = 0; r < rows; ++r)
        for (int t = 0; t < T; ++t) dst[r * T + t] = (double)p[r * Tp + t];
    return BSS_OK;
}

// NMF state is fp64 on the device: plain copies
int put_f64(bss_handle* h, double* dev, const void* src, size_t n) {
    if (!dev) return bss_fail(h, BSS_EINVAL, "this model has no such state");
    BSS_CUDA(h, cudaMemcpyAsync(dev, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    BSS_CUDA(h, bss_wait(h));
    return BSS_OK;
}
int get_f64(bss_handle* h, const double* dev, void* dst, size_t n) {
    if (!dev) return bss_fail(h, BSS_EINVAL, "this model has no such state");
    BSS_CUDA(h, cudaMemcpyAsync(dst, dev, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    BSS_CUDA(h, bss_wait(h));
    return BSS_OK;
}

size_t basis_rows(const bss_handle* h) {
    if (is_nmf(h->cfg.method) || h->cfg.partitioning) return (size_t)h->B * h->F;
    return (size_t)h->B * h->N * h->F;
}
size_t act_rows(const bss_handle* h) {
    if (is_nmf(h->cfg.method) || h->cfg.partitioning) return (size_t)h->B * h->K;
    return (size_t)h->B * h->N * h->K;
}

}  // namespace

extern "C" {

int bss_set_state(bss_handle* h, int which, const void* src, int dtype) {
    if (!h || !src) return BSS_EINVAL;
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    if (is_nmf(h->cfg.method)) {
        if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "NMF state is exchanged as float64");
        switch (which) {
            case BSS_STATE_TARGET:
                BSS_TRY(put_f64(h, h->nz, src, (size_t)h->B * h->F * h->T));
                h->has_input = true;
                return BSS_OK;
            case BSS_STATE_BASIS: return put_f64(h, h->nt, src, (size_t)h->B * h->F * h->K);
            case BSS_STATE_ACTIVATION: return put_f64(h, h->nv, src, (size_t)h->B * h->K * h->T);
        }
        return bss_fail(h, BSS_EINVAL, "state cannot be set");
    }
    if (h->cfg.method == BSS_IS_MNMF) return smnmf_set_state(h, which, src, dtype);
    switch (which) {
        case BSS_STATE_DEMIX_FILTER:
        case BSS_STATE_DIAGONALIZER: {
            if (dtype != BSS_C128) return bss_fail(h, BSS_EINVAL, "filters are exchanged as complex128");
            if (!h->W) return bss_fail(h, BSS_EINVAL, "this model has no demixing filter");
            const size_t n = (size_t)h->B * h->F * h->C * h->C;
            BSS_CUDA(h, cudaMemcpyAsync(h->W, src, n * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
            BSS_TRY(launch_sync_wf(h, h->W, h->Wf, (long long)n));
            BSS_CUDA(h, bss_wait(h));
            h->has_filter = true;
            h->y_valid = false;
            if (h->cfg.spatial == BSS_SPATIAL_ISS && h->cfg.method != BSS_FAST_MNMF && h->has_input)
                BSS_TRY(bss_refresh_estimates(h));
            return BSS_OK;
        }
        case BSS_STATE_BASIS:
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "basis is exchanged as float64");
            return put_rows(h, h->basis, (const double*)src, basis_rows(h), h->K, h->K);
        case BSS_STATE_ACTIVATION:
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "activation is exchanged as float64");
            return put_rows(h, h->act, (const double*)src, act_rows(h), h->T, h->Tp);
        case BSS_STATE_LATENT:
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "latent is exchanged as float64");
            return put_rows(h, h->latent, (const double*)src, (size_t)h->B * h->N, h->K, h->K);
        case BSS_STATE_SPATIAL:
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "spatial_covariance is exchanged as float64");
            return put_rows(h, h->G, (const double*)src, (size_t)h->B * h->N * h->F, h->C, h->C);
        case BSS_STATE_TARGET:
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "target is exchanged as float64");
            BSS_TRY(put_rows(h, h->target, (const double*)src, (size_t)h->B * h->F, h->T, h->Tp));
            h->has_input = true;
            return BSS_OK;
        case BSS_STATE_VARIANCE:
            if (h->cfg.method != BSS_GAUSS_IDLMA) return bss_fail(h, BSS_EINVAL, "only GaussIDLMA takes external variances");
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "variances are exchanged as float64");
            return idlma_set_variance(h, (const double*)src);
    }
    return bss_fail(h, BSS_EINVAL, "state cannot be set");
}

int bss_get_state(bss_handle* h, int which, void* dst, int dtype) {
    if (!h || !dst) return BSS_EINVAL;
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    if (is_nmf(h->cfg.method)) {
        if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "NMF state is exchanged as float64");
        switch (which) {
            case BSS_STATE_TARGET: return get_f64(h, h->nz, dst, (size_t)h->B * h->F * h->T);
            case BSS_STATE_BASIS: return get_f64(h, h->nt, dst, (size_t)h->B * h->F * h->K);
            case BSS_STATE_ACTIVATION: return get_f64(h, h->nv, dst, (size_t)h->B * h->K * h->T);
        }
        return bss_fail(h, BSS_EINVAL, "unknown state");
    }
    if (h->cfg.method == BSS_IS_MNMF) return smnmf_get_state(h, which, dst, dtype);
    switch (which) {
        case BSS_STATE_DEMIX_FILTER:
        case BSS_STATE_DIAGONALIZER: {
            if (dtype != BSS_C128) return bss_fail(h, BSS_EINVAL, "filters are exchanged as complex128");
            if (!h->W) return bss_fail(h, BSS_EINVAL, "this model has no demixing filter");
            const size_t n = (size_t)h->B * h->F * h->C * h->C;
            BSS_CUDA(h, cudaMemcpyAsync(dst, h->W, n * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
            BSS_CUDA(h, bss_wait(h));
            return BSS_OK;
        }
        case BSS_STATE_ESTIMATION: return bss_separate(h, dst, dtype, 0);
        case BSS_STATE_BASIS:
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "basis is exchanged as float64");
            return get_rows(h, h->basis, (double*)dst, basis_rows(h), h->K, h->K);
        case BSS_STATE_ACTIVATION:
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "activation is exchanged as float64");
            return get_rows(h, h->act, (double*)dst, act_rows(h), h->T, h->Tp);
        case BSS_STATE_LATENT:
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "latent is exchanged as float64");
            return get_rows(h, h->latent, (double*)dst, (size_t)h->B * h->N, h->K, h->K);
        case BSS_STATE_SPATIAL:
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "spatial_covariance is exchanged as float64");
            return get_rows(h, h->G, (double*)dst, (size_t)h->B * h->N * h->F, h->C, h->C);
        case BSS_STATE_TARGET:
            if (dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "target is exchanged as float64");
            return get_rows(h, h->target, (double*)dst, (size_t)h->B * h->F, h->T, h->Tp);
        case BSS_STATE_COVARIANCE: {
            if (dtype != BSS_C128) return bss_fail(h, BSS_EINVAL, "covariances are exchanged as complex128");
            if (!h->U) return bss_fail(h, BSS_EINVAL, "this model has no covariances");
            const int C = h->C, CC = C * C;
            const size_t n_mat = (size_t)h->B * h->C * h->F;
            BSS_TRY(ensure_pinned(h, n_mat * CC * sizeof(double)));
            double* p = (double*)h->pinned;
            BSS_CUDA(h, cudaMemcpyAsync(p, h->U, n_mat * CC * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            BSS_CUDA(h, bss_wait(h));
            double* out = (double*)dst;
            for (size_t m = 0; m < n_mat; ++m) {
                const double* q = p + m * CC;
                double* o = out + m * CC * 2;
                for (int i = 0; i < C; ++i) {
                    o[(i * C + i) * 2] = q[i];
                    o[(i * C + i) * 2 + 1] = 0.0;
                }
                int e = C;
                for (int i = 1; i < C; ++i)
                    for (int j = 0; j < i; ++j) {
                        o[(i * C + j) * 2] = q[e];
                        o[(i * C + j) * 2 + 1] = q[e + 1];
                        o[(j * C + i) * 2] = q[e];
                        o[(j * C + i) * 2 + 1] = -q[e + 1];
                        e += 2;
                    }
            }
            return BSS_OK;
        }
        case BSS_STATE_GATE: {
            if (dtype != BSS_I32) return bss_fail(h, BSS_EINVAL, "gate is exchanged as int32");
            if (!h->gate) return bss_fail(h, BSS_EINVAL, "this model has no gate");
            const size_t n = (size_t)h->B * h->C * h->F;
            BSS_CUDA(h, cudaMemcpyAsync(dst, h->gate, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            BSS_CUDA(h, bss_wait(h));
            return BSS_OK;
        }
        case BSS_STATE_ORDER: {
            if (dtype != BSS_I32) return bss_fail(h, BSS_EINVAL, "order is exchanged as int32");
            if (!h->order || h->cfg.spatial != BSS_SPATIAL_IP2) return bss_fail(h, BSS_EINVAL, "only the pairwise (IP2) update has an eigenvalue order");
            BSS_CUDA(h, cudaMemcpyAsync(dst, h->order, (size_t)h->B * h->F * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            BSS_CUDA(h, bss_wait(h));
            return BSS_OK;
        }
        case BSS_STATE_EIGVAL: {
            if (dtype != BSS_C128) return bss_fail(h, BSS_EINVAL, "eigenvalues are exchanged as complex128");
            if (!h->eigval || h->cfg.spatial != BSS_SPATIAL_IP2) return bss_fail(h, BSS_EINVAL, "only the pairwise (IP2) update has eigenvalues");
            BSS_CUDA(h, cudaMemcpyAsync(dst, h->eigval, (size_t)h->B * h->F * 2 * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
            BSS_CUDA(h, bss_wait(h));
            return BSS_OK;
        }
    }
    return bss_fail(h, BSS_EINVAL, "unknown state");
}

// ------------------------------------------------------------------------------------------- primitives
static int primitive_handle(int device, int C, int F, int T, bss_handle** out) {
    bss_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.method = BSS_AUX_LAPLACE_IVA;
    cfg.spatial = BSS_SPATIAL_IP;
    cfg.n_batch = 1;
    cfg.n_channels = cfg.n_sources = C;
    cfg.n_bins = F;
    cfg.n_frames = T;
    cfg.n_basis = 1;
    cfg.device = device;
    cfg.domain = 2.0;
    cfg.eps = 1e-12;
    cfg.threshold = 1e12;
    return bss_create(&cfg, out);
}

int bss_weighted_covariance(int device, int n_channels, int n_weights, int n_bins, int n_frames, const void* x, const double* r,
                            void* u) {
    if (!x || !r || !u) return BSS_EINVAL;
    if (n_weights != n_channels) return BSS_EINVAL;   // every caller in the reference has one weight set per source
    bss_handle* h = nullptr;
    int rc = primitive_handle(device, n_channels, n_bins, n_frames, &h);
    if (rc != BSS_OK) return rc;
    auto done = [&](int code) {
        bss_destroy(h);
        return code;
    };
    rc = bss_set_input(h, x, BSS_C128);
    if (rc != BSS_OK) return done(rc);
    const size_t nr = (size_t)n_weights * n_bins * n_frames;
    double* r_dev = nullptr;
    if (cudaMalloc(&r_dev, nr * sizeof(double)) != cudaSuccess) return done(BSS_ENOMEM);
    if (cudaMalloc(&h->iw, (size_t)n_bins * n_weights * h->Tp * sizeof(float)) != cudaSuccess) {
        cudaFree(r_dev);
        return done(BSS_ENOMEM);
    }
    cudaMemcpyAsync(r_dev, r, nr * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    rc = launch_import_weights(h, r_dev, h->iw, n_weights, n_bins, n_frames, h->Tp);
    if (rc == BSS_OK) {
        CovArgs c{};
        c.X = h->X;
        c.U = h->U;
        c.B = 1;
        c.F = n_bins;
        c.C = n_channels;
        c.NW = n_weights;
        c.T = n_frames;
        c.Tp = h->Tp;
        c.wmode = WM_EXPLICIT;
        c.iw = h->iw;
        c.n_sel = n_weights;
        for (int i = 0; i < 8; ++i) c.wsel[i] = i;
        rc = launch_covariance(h, c);
    }
    if (rc == BSS_OK) rc = bss_get_state(h, BSS_STATE_COVARIANCE, u, BSS_C128);
    cudaFree(r_dev);
    return done(rc);
}

int bss_ip_update(int device, int n_channels, int n_bins, void* w, const void* u, int32_t* gate, double threshold, int floor_den,
                  double eps) {
    if (!w || !u) return BSS_EINVAL;
    bss_handle* h = nullptr;
    int rc = primitive_handle(device, n_channels, n_bins, 2, &h);
    if (rc != BSS_OK) return rc;
    auto done = [&](int code) {
        bss_destroy(h);
        return code;
    };
    const int C = n_channels, CC = C * C;
    rc = bss_set_state(h, BSS_STATE_DEMIX_FILTER, w, BSS_C128);
    if (rc != BSS_OK) return done(rc);
    // pack the Hermitian covariances
    const size_t n_mat = (size_t)C * n_bins;
    std::vector<double> packed(n_mat * CC);
    const double* uu = (const double*)u;
    for (size_t m = 0; m < n_mat; ++m) {
        const double* o = uu + m * CC * 2;
        double* q = packed.data() + m * CC;
        for (int i = 0; i < C; ++i) q[i] = o[(i * C + i) * 2];
        int e = C;
        for (int i = 1; i < C; ++i)
            for (int j = 0; j < i; ++j) {
                q[e] = o[(i * C + j) * 2];
                q[e + 1] = o[(i * C + j) * 2 + 1];
                e += 2;
            }
    }
    cudaMemcpyAsync(h->U, packed.data(), packed.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    IpArgs a{};
    a.W = h->W;
    a.Wf = h->Wf;
    a.U = h->U;
    a.gate = h->gate;
    a.flags = h->flags;
    a.B = 1;
    a.F = n_bins;
    a.C = C;
    a.threshold = threshold;
    a.eps = eps;
    a.use_gate = 1;
    a.floor_den = floor_den;
    a.pair_m = a.pair_n = -1;
    rc = launch_ip(h, a);
    if (rc == BSS_OK) rc = bss_synchronize(h);
    if (rc == BSS_OK) rc = bss_get_state(h, BSS_STATE_DEMIX_FILTER, w, BSS_C128);
    if (rc == BSS_OK && gate) rc = bss_get_state(h, BSS_STATE_GATE, gate, BSS_I32);
    return done(rc);
}

int bss_least_squares_map(int device, int n_rows_a, int n_rows_b, int n_bins, int n_frames, const void* a, const void* b,
                          void* out) {
    if (!a || !b || !out) return BSS_EINVAL;
    if (n_rows_a < 1 || n_rows_a > 8 || n_rows_b < 1 || n_rows_b > 8 || n_bins < 1 || n_frames < 1) return BSS_EINVAL;
    // a throw-away handle only for its stream, error slot and singular-bin flag (two channels, two frames: no big buffers)
    bss_handle* h = nullptr;
    int rc = primitive_handle(device, 2, 1, 2, &h);
    if (rc != BSS_OK) return rc;
    const size_t na = (size_t)n_rows_a * n_bins * n_frames, nb = (size_t)n_rows_b * n_bins * n_frames;
    const size_t no = (size_t)n_rows_a * n_rows_b * n_bins;
    double2 *da = nullptr, *db = nullptr, *dout = nullptr;
    auto done = [&](int code) {
        if (da) cudaFree(da);
        if (db) cudaFree(db);
        if (dout) cudaFree(dout);
        bss_destroy(h);
        return code;
    };
    if (cudaMalloc(&da, na * sizeof(double2)) != cudaSuccess || cudaMalloc(&db, nb * sizeof(double2)) != cudaSuccess ||
        cudaMalloc(&dout, no * sizeof(double2)) != cudaSuccess)
        return done(BSS_ENOMEM);
    cudaMemcpyAsync(da, a, na * sizeof(double2), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(db, b, nb * sizeof(double2), cudaMemcpyHostToDevice, h->stream);
    rc = launch_lsq_map(h, da, db, dout, n_rows_a, n_rows_b, n_bins, n_frames);
    if (rc == BSS_OK) rc = bss_synchronize(h);
    if (rc == BSS_OK) {
        cudaMemcpyAsync(out, dout, no * sizeof(double2), cudaMemcpyDeviceToHost, h->stream);
        if (bss_wait(h) != cudaSuccess) rc = BSS_ECUDA;
    }
    return done(rc);
}

int bss_demix(int device, int n_channels, int n_bins, int n_frames, int flags, const void* x, const void* w, void* y) {
    (void)flags;
    if (!x || !w || !y) return BSS_EINVAL;
    bss_handle* h = nullptr;
    int rc = primitive_handle(device, n_channels, n_bins, n_frames, &h);
    if (rc != BSS_OK) return rc;
    rc = bss_set_input(h, x, BSS_C128);
    if (rc == BSS_OK) rc = bss_set_state(h, BSS_STATE_DEMIX_FILTER, w, BSS_C128);
    if (rc == BSS_OK) rc = bss_separate(h, y, BSS_C128, 0);
    bss_destroy(h);
    return rc;
}

int bss_projection_back_scale(int device, int n_channels, int n_bins, int n_frames, const void* x, const void* w, int reference_id,
                              void* scale) {
    if (!x || !w || !scale) return BSS_EINVAL;
    if (reference_id < 0 || reference_id >= n_channels) return BSS_EINVAL;
    bss_handle* h = nullptr;
    int rc = primitive_handle(device, n_channels, n_bins, n_frames, &h);
    if (rc != BSS_OK) return rc;
    auto done = [&](int code) {
        bss_destroy(h);
        return code;
    };
    rc = bss_set_input(h, x, BSS_C128);
    if (rc == BSS_OK) rc = bss_set_state(h, BSS_STATE_DEMIX_FILTER, w, BSS_C128);
    if (rc == BSS_OK) rc = launch_pb_scale(h, h->W, h->Cx, (double2*)h->scale, 1, n_bins, n_channels, reference_id);
    if (rc == BSS_OK) rc = bss_synchronize(h);
    if (rc == BSS_OK) {
        cudaMemcpyAsync(scale, h->scale, (size_t)n_channels * n_bins * sizeof(double2), cudaMemcpyDeviceToHost, h->stream);
        bss_wait(h);
    }
    return done(rc);
}

}  // extern "C"
