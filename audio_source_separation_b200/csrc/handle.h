// Internal definition of bss_handle: device-resident state of one batch of mixtures.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/bssgpu.h"
#include "common.cuh"

// Device layouts (B = batch, Tp = n_frames rounded up to an even number so every frame row is a
// multiple of 16 bytes -- the granule of cp.async.bulk; pad frames hold zeros):
//   X     cf      [B][F][C][Tp]     mixture, bin-major: one bin tile (C rows) is contiguous
//   Y     cf      [B][F][N][Tp]     estimates (ISS state / scratch), same tiling
//   W     double2 [B][F][N][C]      demixing filters (FastMNMF: diagonaliser Q)
//   basis float   [B][N][F][K]      ([B][F][K] when partitioned)
//   act   float   [B][N][K][Tp]     ([B][K][Tp] when partitioned)
//   U     double  [B][NW][F][C*C]   packed Hermitian weighted covariances (diag, then lower (re,im))
//   Cx    double  [B][F][C*C]       packed plain covariance mean_t x x^H
struct bss_handle {
    bss_config cfg{};
    int B = 1, C = 0, N = 0, F = 0, T = 0, Tp = 0, K = 0;
    int n_sm = 148;
    int max_smem = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    void* graph_rec_exec = nullptr;   // same for bss_run_record: two x (update_once, loss, append to the device-side history)
    uint64_t graph_rec_sig = 0;
    int64_t graph_rec_launches = 0;
    int* loss_counter = nullptr;   // device-side write position of the loss history
    void* graph_exec = nullptr;    // cudaGraphExec_t of the last captured pair of iterations (bss_run)
    uint64_t graph_sig = 0;        // signature of every device pointer the captured kernels were given (0: do not reuse)
    int64_t graph_launches = 0;    // kernel launches per replay of graph_exec
    std::string err;

    bool has_input = false;
    bool has_filter = false;      // W holds a valid demixing filter (false for ISS between updates)
    bool has_variance = false;    // GaussIDLMA: BSS_STATE_VARIANCE has been set
    bool y_valid = false;         // Y buffer holds separate(X, W) / the ISS state
    int pair_m = -1, pair_n = -1; // IP2 pair
    bool pair_started = false;

    cf* X = nullptr;
    cf* Y = nullptr;
    double2* W = nullptr;
    float* basis = nullptr;
    float* act = nullptr;
    float* basis2 = nullptr;       // double buffer of basis (the basis update reads all K of the old one)
    cf* Wf = nullptr;              // [B][F][N][C] fp32 mirror of W for the streaming kernels
    float* latent = nullptr;       // [B][N][K]
    double* U = nullptr;
    double* Cx = nullptr;
    int32_t* gate = nullptr;       // [B][N][F]
    int32_t* flags = nullptr;      // [0] singular-bin counter
    double* pw = nullptr;          // [B][N][F] per-bin source power (normalisation)
    double* scale = nullptr;       // [B][N][F] complex (projection back)  -> 2 doubles each
    float* wfr = nullptr;          // [B][N][Tp] AuxIVA frame weights (inverse, floored)
    float* wraw = nullptr;         // [B][N][Tp] AuxIVA frame weights (raw)
    int32_t* order = nullptr;      // [B][F][2] IP2 eigenvalue order
    double2* eigval = nullptr;     // [B][F][2] IP2 eigenvalues that `order` indexes
    int opt_source_model = 0;      // BSS_OPT_SOURCE_MODEL
    int last_source_model = 0;     // BSS_INFO_SOURCE_MODEL
    int opt_blocking_sync = 0;     // BSS_OPT_BLOCKING_SYNC
    int opt_async_input = 0;       // BSS_OPT_ASYNC_INPUT
    cudaEvent_t ev_block = nullptr;
    int opt_act_chunks = 0;        // BSS_OPT_ACT_CHUNKS
    int last_act_chunks = 0;       // BSS_INFO_ACT_CHUNKS
    int opt_ip_kernel = 0;         // BSS_OPT_IP_KERNEL
    int last_ip_kernel = 0;        // BSS_INFO_IP_KERNEL
    int64_t graph_replays = 0;     // BSS_INFO_GRAPH_REPLAYS
    double* logdet = nullptr;      // [B][F]
    double* aux = nullptr;         // [B][N] power-normalisation factors of the last update
    double2* G2x = nullptr;        // [B][F][N][C] cross covariance Y X^H / T (ISS filter recovery)
    float* part = nullptr;         // partial sums of the cross-bin reductions
    size_t part_elems = 0;
    float* P = nullptr;            // [B][128-frame block][F][N][128] floats: source power handed from the basis to the activation update
    float* iw = nullptr;           // [B][F][NW][Tp] explicit inverse weights (generic covariance path)
    double* lossbuf = nullptr;     // [B][F] per-bin loss terms + [B] results
    void* staging = nullptr;       // device staging for host <-> device layout conversion
    size_t staging_bytes = 0;
    void* scratch2 = nullptr;      // second device scratch (ISTFT frames + time-domain output of the waveform path)
    size_t scratch2_bytes = 0;
    float* fft_win = nullptr;      // cached STFT tables of the waveform path: window, twiddles (kernels_stft.cu)
    float2* fft_tw = nullptr;
    int fft_N = 0;
    double fft_win_sum = 0.0;
    std::vector<double> fft_window_host;
    void* pinned = nullptr;        // pinned host bounce buffer
    size_t pinned_bytes = 0;

    float* latent2 = nullptr;      // scratch for the partitioned latent update
    float* beff = nullptr;         // [B][N][F][K]  partitioned model: Z[n,k] T[f,k]
    float* aeff = nullptr;         // [B][N][K][Tp] partitioned model: V[k,t] replicated per source
    float* praw = nullptr;         // [B][N][F][K][2] raw statistics of the basis kernel
    // FastMNMF
    float* G = nullptr;            // [B][N][F][M]
    float* G2 = nullptr;           // double buffer of G (the spatial update reads all sources of the old one)
    float* xt = nullptr;           // [B][F][M][Tp] |Q x|^2 (tile layout, rows = M); valid while xt_valid
    bool xt_valid = false;
    float* mn_acc = nullptr;       // numerator / denominator accumulators
    float* mn_acc2 = nullptr;
    // NMF (fp64 throughout): target [B][F][T], factors [B][F][K] and [B][K][T]
    float* target = nullptr;       // unused legacy slot
    double* nz = nullptr;
    double* nt = nullptr;
    double* nv = nullptr;
    double* npart = nullptr;       // partial sums of the activation update
    size_t npart_elems = 0;
    // Sawada IS-MNMF (fp64 throughout)
    double* sH = nullptr;          // [B][F][N][C*C] packed Hermitian spatial covariances
    double* sZ = nullptr;          // [B][N][K] latent
    double* sT = nullptr;          // [B][F][K] basis
    double* sV = nullptr;          // [B][K][T] activation
    double* sStat = nullptr;       // [2][B][N][F][T] tr(Xh^-1 X Xh^-1 H_n), tr(Xh^-1 H_n)
    double* sPart = nullptr;       // partial sums of the factor updates
    double* sAcc = nullptr;        // [B][F][N][2][C*C] packed: sum_t lambda Xh^-1, sum_t lambda q q^H
    // loss history of bss_run_record: [n_iter][B]
    double* loss_hist = nullptr;
    size_t loss_hist_elems = 0;
};

// Wait for the handle's stream.  Default: cudaStreamSynchronize (the driver spins, lowest latency).  With
// BSS_OPT_BLOCKING_SYNC the thread sleeps on a blocking event instead: a pipelined whole-job call runs several host threads
// per GPU and several processes per node, and that many spinning waiters starve the threads that still have launches to issue.
static inline cudaError_t bss_wait(bss_handle* h) {
    if (!h->opt_blocking_sync) return cudaStreamSynchronize(h->stream);
    if (!h->ev_block) {
        cudaError_t e = cudaEventCreateWithFlags(&h->ev_block, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaEventRecord(h->ev_block, h->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(h->ev_block);
}

#define BSS_CUDA(h, call)                                                                              \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                            \
            return BSS_ECUDA;                                                                          \
        }                                                                                              \
    } while (0)

#define BSS_TRY(expr)                    \
    do {                                 \
        int rc__ = (expr);               \
        if (rc__ != BSS_OK) return rc__; \
    } while (0)

static inline int bss_fail(bss_handle* h, int code, const std::string& msg) {
    h->err = msg;
    return code;
}

static inline int ensure_staging(bss_handle* h, size_t bytes) {
    if (bytes <= h->staging_bytes) return BSS_OK;
    if (h->staging) cudaFree(h->staging);
    h->staging = nullptr;
    h->staging_bytes = 0;
    BSS_CUDA(h, cudaMalloc(&h->staging, bytes));
    h->staging_bytes = bytes;
    return BSS_OK;
}

static inline int ensure_scratch2(bss_handle* h, size_t bytes) {
    if (bytes <= h->scratch2_bytes) return BSS_OK;
    if (h->scratch2) cudaFree(h->scratch2);
    h->scratch2 = nullptr;
    h->scratch2_bytes = 0;
    BSS_CUDA(h, cudaMalloc(&h->scratch2, bytes));
    h->scratch2_bytes = bytes;
    return BSS_OK;
}

static inline int ensure_pinned(bss_handle* h, size_t bytes) {
    if (bytes <= h->pinned_bytes) return BSS_OK;
    if (h->pinned) cudaFreeHost(h->pinned);
    h->pinned = nullptr;
    h->pinned_bytes = 0;
    BSS_CUDA(h, cudaMallocHost(&h->pinned, bytes));
    h->pinned_bytes = bytes;
    return BSS_OK;
}

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

// Shared-memory plan of a warp-stream kernel: [wpc][stages] mbarriers | [wpc] scratch | [wpc][stages] stages
struct StreamPlan {
    int wpc;
    uint32_t scratch_off, scratch_stride, ring_off;
    size_t smem_bytes;
    unsigned grid;
};
static inline bool plan_stream(const bss_handle* h, const TileGeom& g, int stages, size_t scratch_per_warp,
                               long long n_items, int max_wpc, StreamPlan* out, size_t reserve_bytes = 0) {
    const uint32_t scratch_stride = (uint32_t)round_up((int)scratch_per_warp, 16);
    const size_t per_warp = (size_t)stages * g.stage_bytes + scratch_stride + (size_t)stages * 8;
    if ((size_t)h->max_smem < 1024 + reserve_bytes + per_warp) return false;
    int wpc = (int)(((size_t)h->max_smem - 512 - reserve_bytes) / per_warp);
    if (wpc > max_wpc) wpc = max_wpc;
    if (wpc < 1 || n_items > 0x7fffffffLL) return false;   // the device-side cursors are 32 bit
    const uint32_t bars_bytes = (uint32_t)round_up(wpc * stages * 8, 128);
    out->wpc = wpc;
    out->scratch_off = bars_bytes;
    out->scratch_stride = scratch_stride;
    out->ring_off = (uint32_t)round_up((int)(bars_bytes + wpc * scratch_stride), 128);
    out->smem_bytes = out->ring_off + (size_t)wpc * stages * g.stage_bytes;
    int ctas_per_sm = (int)((size_t)h->max_smem / (out->smem_bytes + reserve_bytes + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const int by_threads = 2048 / (wpc * 32);
    if (ctas_per_sm > by_threads) ctas_per_sm = by_threads;
    if (ctas_per_sm > 4) ctas_per_sm = 4;
    long long grid = cdiv(n_items, wpc);
    const long long cap = (long long)h->n_sm * ctas_per_sm;
    if (grid > cap) grid = cap;
    out->grid = (unsigned)grid;
    return true;
}

// Same plan with room for a per-CTA cache of two mixtures' worth of per-mixture data (`bytes_per_mix`), taken only
// when it costs no warps and a CTA's contiguous item range (cdiv(n_items, grid)) cannot span more than two mixtures.
static inline bool plan_stream_cached(const bss_handle* h, const TileGeom& g, int stages, size_t scratch_per_warp, long long n_items,
                                      int max_wpc, size_t bytes_per_mix, long long items_per_mix, StreamPlan* out, uint32_t* cache_off,
                                      size_t* smem_bytes) {
    StreamPlan full;
    const size_t cache_bytes = 2 * bytes_per_mix;
    if (!plan_stream(h, g, stages, scratch_per_warp, n_items, max_wpc, &full)) return false;
    if (!plan_stream(h, g, stages, scratch_per_warp, n_items, max_wpc, out, cache_bytes + 16)) return false;
    if (out->wpc != full.wpc || out->grid != full.grid) return false;
    if (cdiv(n_items, out->grid) > items_per_mix) return false;
    *cache_off = (uint32_t)round_up((int)out->smem_bytes, 16);
    *smem_bytes = *cache_off + cache_bytes;
    return true;
}

// kernels_*.cu launchers -------------------------------------------------------------------------
// weight modes of the covariance-accumulate kernel
enum { WM_UNIT = 0, WM_ILRMA = 1, WM_FRAME = 2, WM_EXPLICIT = 3 };

struct CovArgs {
    const cf* X;          // [B][F][C][Tp]
    double* U;            // [B][NW][F][C*C]
    int B, F, C, NW, T, Tp;
    int wmode;
    // WM_ILRMA
    const float* basis;   // [B][N][F][K]
    const float* act;     // [B][N][K][Tp]
    int K;
    float expo;           // 2/domain
    float eps;
    // WM_FRAME
    const float* wfr;     // [B][NW][Tp]  inverse weights
    // WM_EXPLICIT
    const float* iw;      // [B][F][NW][Tp] inverse weights
    int wsel[8];          // weight sets to compute (IP2 computes only its pair)
    int n_sel;
};
int launch_covariance(bss_handle* h, const CovArgs& a);
int launch_plain_covariance(bss_handle* h, const cf* X, double* Cx, int B, int F, int C, int T, int Tp);   // fp64 accumulation

struct IpArgs {
    double2* W;           // [B][F][N][C]
    const double* U;      // [B][NW][F][C*C]
    const double* Cx;     // [B][F][C*C]  (may be null)
    int32_t* gate;        // [B][N][F]     (may be null)
    double* pw;           // [B][N][F]     (may be null): w_n^H Cx w_n after the sweep
    int32_t* flags;
    int B, F, C;
    double threshold, eps;
    int use_gate, floor_den;
    int pair_m, pair_n;   // >= 0: IP2 update of this pair instead of the full sweep
    int32_t* order;       // [B][F][2] IP2 eigenvalue order (may be null)
    double2* eigval;      // [B][F][2] IP2 eigenvalues (may be null)
    int variant;          // BSS_OPT_IP_KERNEL: 0 = by problem size, 1 = thread per bin, 2 = lane group per bin
    cf* Wf;               // [B][F][N][C] fp32 mirror of W kept for the streaming kernels (may be null)
};
int launch_ip(bss_handle* h, const IpArgs& a);
int launch_pb_scale(bss_handle* h, const double2* W, const double* Cx, double2* scale, int B, int F, int C, int ref);
int launch_logdet(bss_handle* h, const double2* W, double* out, long long n_bins, int C, int transpose_sq);
int launch_lsq_map(bss_handle* h, const double2* A, const double2* Bm, double2* out, int Ra, int Rb, int F, int T);
int launch_lsq_filter(bss_handle* h, const double2* G, const double* Cx, double2* W, long long n_bins, int C);

struct MuArgs {
    const cf* X;          // [B][F][C][Tp]
    const cf* Y;          // [B][F][N][Tp]; non-null: take |Y|^2 from the stored estimates (ISS)
    const cf* Wf;         // [B][F][N][C]
    const float* basis;   // [B][N][F][K]
    float* basis_out;
    const float* act;     // [B][N][K][Tp]
    int B, F, C, T, Tp, K;
    int mode;             // 0 Gauss (IS divergence), 1 Student-t
    float p_exp, q_exp;   // (d+2)/d and d/(d+2)
    float nu, eps;
    int sel_m, sel_n;     // >= 0: pairwise source model, only these two sources move
    float* raw;           // non-null: basis kernel stores the raw (num, den) sums [B][N][F][K][2] instead of updating
    float* Pout;          // non-null (n_basis == 2, filter-based): the basis kernel also stores |y|^2 as float tiles, block-major
                          // [B][128-frame block][F][N][128] (a block's bins are consecutive), half the bytes of X ...
    const float* Pin;     // ... which the activation kernel then streams instead of X: no second y = W x, half the traffic
};
int launch_mu_basis(bss_handle* h, const MuArgs& a);
int launch_mu_act(bss_handle* h, const MuArgs& a, float* act, int* n_chunks_out = nullptr);
int launch_mu_act_finish(bss_handle* h, const MuArgs& a, float* act, int n_chunks);
int launch_mu_fused(bss_handle* h, const MuArgs& a, int* n_chunks_out, bool* done);   // kernels_mu_fused.cu
int launch_normalize_power(bss_handle* h, double2* W, cf* Wf, float* basis, const double* pw, int B, int N, int C, int F, int K,
                           double domain, double eps, double* aux_out);
int launch_normalize_pb(bss_handle* h, double2* W, cf* Wf, float* basis, const double2* scale, int B, int N, int C, int F, int K,
                        double domain);
int launch_sync_wf(bss_handle* h, const double2* W, cf* Wf, long long n);
int launch_identity_filter(bss_handle* h, double2* W, cf* Wf, long long n_bins, int N, int C);
int launch_widen(bss_handle* h, const cf* in, double2* out, long long n);
int launch_separate(bss_handle* h, const cf* X, const cf* Wf, const double2* scale, cf* Y, cf* out, int B, int C, int F, int T,
                    int Tp);
int launch_export_y(bss_handle* h, const cf* Y, const double2* scale, cf* out, int B, int N, int F, int T, int Tp);
int launch_ilrma_loss(bss_handle* h, const MuArgs& a, float expo, double* terms);
int launch_loss_finish(bss_handle* h, const double* terms, const double* logdet, double coef, int B, int F, double* out);
int launch_loss_append(bss_handle* h, const double* result, double* hist, int* counter, int B, int capacity);
int launch_import_x(bss_handle* h, const void* staged, int dtype, cf* X, int B, int C, int F, int T, int Tp);
int launch_frame_weights(bss_handle* h, const cf* src, const cf* Wf, int from_y, float* winv, float* raw, int B, int C, int F,
                         int T, int Tp, int kind, float eps);
int launch_t_weights(bss_handle* h, const cf* X, const cf* Wf, const float* basis, const float* act, float* iw, int B, int C,
                     int F, int K, int Tp, float nu, float eps);
int launch_import_weights(bss_handle* h, const double* r_dev, float* iw, int N, int F, int T, int Tp);
int launch_import_variance(bss_handle* h, const double* r_dev, float* iw, int B, int N, int F, int T, int Tp, double eps);
int launch_idlma_loss(bss_handle* h, const cf* X, const cf* Wf, const float* iw, double* terms, int B, int C, int F, int T, int Tp);
int launch_iss(bss_handle* h, cf* Y, int mode, const float* basis, const float* act, const float* wfr, double* pw, int B, int N,
               int F, int T, int Tp, int K, float expo, float eps);
int launch_cross_cov(bss_handle* h, const cf* Y, const cf* X, double2* G, long long n_bins, int C, int T, int Tp);
int launch_scale_y(bss_handle* h, cf* Y, float* basis, const double* aux, const double2* scale, int B, int N, int F, int Tp, int K,
                   double domain);
int launch_aux_from_power(bss_handle* h, const double* pw, double* aux, int B, int N, int F, double eps);
// single-channel NMF (kernels_nmf.cu)
struct NmfMath {
    int kind;          // 0 EUC, 1 KL, 2 IS, 3 t, 4 Cauchy
    int alg;           // enum bss_nmf_algorithm
    double a, b, p, q; // exponents of the update (see methods_nmf.cu)
    double nu, eps;
    double loss_expo;  // 2 / domain
    double loss_eps;
};
int launch_nmf_update(bss_handle* h, const NmfMath& m, const double* Z, double* Tm, double* V, int B, int F, int T, int K);
int launch_nmf_fused(bss_handle* h, const NmfMath& m, const double* Z, double* Tm, double* V, double* loss_hist, int B, int F, int T,
                     int K, int n_iter, bool* done);
int launch_nmf_loss(bss_handle* h, const NmfMath& m, const double* Z, const double* Tm, const double* V, double* terms, int B, int F,
                    int T, int K);
// partitioned ILRMA (kernels_part.cu)
int launch_part_expand(bss_handle* h);
int launch_part_latent(bss_handle* h);
int launch_part_basis(bss_handle* h);
int launch_part_act_finish(bss_handle* h, int n_chunks);
int launch_part_normalize(bss_handle* h);
// FastMNMF (kernels_mnmf.cu)
int launch_mnmf_xt(bss_handle* h);
int launch_mnmf_basis(bss_handle* h);
int launch_mnmf_act(bss_handle* h);
int launch_mnmf_scm(bss_handle* h);
int launch_mnmf_weights(bss_handle* h, int tiled);
int launch_covariance8(bss_handle* h, bool* done);   // kernels_cov8.cu: FastMNMF, 8 channels, weights computed in the kernel
int launch_covariance_mma(bss_handle* h, const cf* X, const float* iw_tiled, double* U, int B, int F, int C, int NW, int T, int Tp,
                          bool* done);
int launch_mnmf_loss_terms(bss_handle* h);
int launch_mnmf_separate(bss_handle* h, cf* out);
int launch_mnmf_normalize(bss_handle* h);
// Sawada IS-MNMF (kernels_smnmf.cu)
void smnmf_act_plan(const bss_handle* h, int* bins, int* chunks);
int launch_smnmf_stats(bss_handle* h);
int launch_smnmf_factor(bss_handle* h, int which);
int launch_smnmf_spatial(bss_handle* h, int normalize);
int launch_smnmf_loss_terms(bss_handle* h, double* terms);
int launch_smnmf_separate(bss_handle* h, cf* out);
int launch_sum_frames(bss_handle* h, const float* raw, int B, int N, int T, int Tp, int kind, double coef, double eps, double* out);
