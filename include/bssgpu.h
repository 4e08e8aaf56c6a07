/*
 * bssgpu.h -- C ABI of libbssgpu.so, the B200 (sm_100a) implementation of the
 * STFT-domain blind-source-separation update loop.
 *
 * The reference (tky823/audio_source_separation) is pure Python/NumPy and has no FFI of
 * its own; the boundary it offers is the Python class surface
 *     Model.__init__ / Model.__call__ / Model.update_once / Model.separate /
 *     Model.compute_negative_loglikelihood
 * (src/bss/ilrma.py:183-677, src/bss/iva.py:388-802, src/bss/mnmf.py:637-946,
 *  src/algorithm/nmf.py:10-595).  Each entry point below names the reference function it
 * replaces.  All pointers are plain host pointers unless the name says `_device`; arrays use the
 * reference's own layouts with one extra leading batch axis B (independent mixtures; B = 1
 * reproduces the reference shapes exactly).  Nothing in this header depends on torch.
 *
 * Conventions
 *   - every function returns BSS_OK (0) or a negative bss_status; bss_last_error() gives text
 *   - host pointers are borrowed for the duration of the call only
 *   - a handle is bound to one device and one stream and is not thread-safe
 *   - dtype arguments select the HOST element type; device storage is complex64/float32 for
 *     the big tensors and float64 for the per-bin linear algebra (demixing filters, covariances)
 *
 * Environment switches (measurement aids; every alternative gives the same results, the first two bit for bit)
 *   BSSGPU_NO_GRAPH=1           bss_run queues every iteration eagerly instead of replaying a CUDA graph
 *   BSSGPU_NO_POWER_HANDOFF=1   the activation update recomputes |W x|^2 from the mixture instead of reading the
 *                               powers the basis update stored (n_basis == 2)
 *   BSSGPU_NO_CLUSTER=1         NMF runs its per-phase kernels instead of the single cluster launch
 *   BSSGPU_NO_MMA=1             FastMNMF covariances / spatial update on CUDA cores instead of tensor cores
 */
#ifndef BSSGPU_H
#define BSSGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bss_handle bss_handle;

enum bss_status {
    BSS_OK = 0,
    BSS_EINVAL = -1,       /* bad argument / unsupported size (ValueError)              */
    BSS_ECUDA = -2,        /* CUDA runtime error (RuntimeError)                         */
    BSS_ESINGULAR = -3,    /* exactly singular bin: np.linalg.LinAlgError in the reference */
    BSS_ENOMEM = -4,
    BSS_ESTATE = -5,       /* call order violated (e.g. update before set_input)        */
    BSS_EUNSUPPORTED = -6, /* NotImplementedError in the reference                      */
    BSS_ENCCL = -7         /* NCCL missing or a collective failed (RuntimeError)        */
};

enum bss_method {
    BSS_GAUSS_ILRMA = 0,      /* src/bss/ilrma.py:178  GaussILRMA            */
    BSS_T_ILRMA = 1,          /* src/bss/ilrma.py:713  tILRMA                */
    BSS_AUX_LAPLACE_IVA = 2,  /* src/bss/iva.py:388    AuxLaplaceIVA         */
    BSS_AUX_GAUSS_IVA = 3,    /* src/bss/iva.py:621    AuxGaussIVA           */
    BSS_FAST_MNMF = 4,        /* src/bss/mnmf.py:637   FastMultichannelISNMF */
    BSS_IS_MNMF = 5,          /* src/bss/mnmf.py:116   MultichannelISNMF(author='Sawada'); `normalize` != 0 is normalize=True */
    BSS_GAUSS_IDLMA = 6,      /* src/sss/idlma.py:88   GaussIDLMA: spatial model with source variances supplied by the caller
                                 (BSS_STATE_VARIANCE, the output of the caller's DNN); update_once = update_space_model, bss_normalize = its normalisation */
    BSS_NMF_EUC = 10,         /* src/algorithm/nmf.py:150 EUCNMF             */
    BSS_NMF_KL = 11,          /* src/algorithm/nmf.py:209 KLNMF              */
    BSS_NMF_IS = 12,          /* src/algorithm/nmf.py:268 ISNMF              */
    BSS_NMF_T = 13,           /* src/algorithm/nmf.py:358 tNMF               */
    BSS_NMF_CAUCHY = 14       /* src/algorithm/nmf.py:430 CauchyNMF          */
};

enum bss_spatial {            /* `algorithm_spatial` of the reference constructors */
    BSS_SPATIAL_IP = 0,       /* 'IP' / 'IP1'      */
    BSS_SPATIAL_ISS = 1,      /* 'ISS'             */
    BSS_SPATIAL_IP2 = 2       /* 'IP2' / 'pairwise' */
};

enum bss_normalize {
    BSS_NORMALIZE_NONE = 0,
    BSS_NORMALIZE_POWER = 1,            /* 'power'            src/bss/ilrma.py:304-322 */
    BSS_NORMALIZE_PROJECTION_BACK = 2   /* 'projection-back'  src/bss/ilrma.py:323-330 */
};

enum bss_nmf_algorithm {      /* `algorithm` of the NMF constructors */
    BSS_ALG_MM = 0,
    BSS_ALG_ME = 1,
    BSS_ALG_NAIVE = 2,        /* CauchyNMF 'naive-multipricative' */
    BSS_ALG_MM_FAST = 3       /* CauchyNMF 'mm_fast'              */
};

enum bss_dtype { BSS_F32 = 0, BSS_F64 = 1, BSS_C64 = 2, BSS_C128 = 3, BSS_I32 = 4,
                 BSS_I16 = 5 /* 16-bit PCM waveforms (bss_set_input_waveform): sample / 32768, as the reference's notebooks read wav files */ };

/* State tensors (host layouts, leading B omitted):
 *   DEMIX_FILTER   (F,N,C) complex   model.demix_filter
 *   ESTIMATION     (N,F,T) complex   model.estimation
 *   BASIS          (N,F,K) real      model.basis        ((F,K) when partitioning / NMF)
 *   ACTIVATION     (N,K,T) real      model.activation   ((K,T) when partitioning / NMF)
 *   LATENT         (N,K)   real      model.latent       (partitioning only)
 *   DIAGONALIZER   (F,M,M) complex   FastMNMF model.diagonalizer
 *   SPATIAL        (N,F,M) real      FastMNMF model.spatial_covariance
 *                  (F,N,C,C) complex IS-MNMF model.spatial (Hermitian; basis (F,K), activation (K,T), latent (N,K))
 *   TARGET         (F,T)   real      NMF target
 *   COVARIANCE     (N,F,C,C) complex weighted covariances of the last spatial update (read only)
 *   GATE           (N,F)   int32     condition-number gate decisions of the last IP update (read only)
 *   VARIANCE       (N,F,T) real      GaussIDLMA: R = dnn_output^(2/domain) (floored at eps on the device, src/sss/idlma.py:190) (write only)
 */
enum bss_state {
    BSS_STATE_DEMIX_FILTER = 0,
    BSS_STATE_ESTIMATION = 1,
    BSS_STATE_BASIS = 2,
    BSS_STATE_ACTIVATION = 3,
    BSS_STATE_LATENT = 4,
    BSS_STATE_DIAGONALIZER = 5,
    BSS_STATE_SPATIAL = 6,
    BSS_STATE_TARGET = 7,
    BSS_STATE_COVARIANCE = 8,
    BSS_STATE_GATE = 9,
    BSS_STATE_VARIANCE = 10,
    BSS_STATE_ORDER = 11,     /* (F,2) int32    IP2: order = argsort(eigenvalues)[::-1] of the last pairwise update
                                                (src/bss/ilrma.py:608-611, src/bss/iva.py:574-577) (read only)        */
    BSS_STATE_EIGVAL = 12     /* (F,2) complex  IP2: the two eigenvalues of V_n^-1 V_m that `order` indexes, in the order the
                                                device produced them ('+' root, '-' root of the 2x2 characteristic
                                                polynomial) (read only)                                              */
};

/* Per-handle switches (bss_set_option) and read-outs (bss_get_info).  Every alternative computes the same update; the
 * options exist so that tests can force a small problem onto the kernel a large one would take, and so that
 * measurements can compare them. */
enum bss_option {
    BSS_OPT_IP_KERNEL = 0,    /* iterative-projection sweep: 0 = choose by problem size (default), 1 = one thread per bin
                                 (ip_sweep_kernel), 2 = lane group per bin (ip_sweep_group_kernel)                          */
    BSS_OPT_SOURCE_MODEL = 3, /* NMF source model of (t-)ILRMA: 0 = fused single-pass kernel where it covers the configuration
                                 (Gauss, domain 2, n_basis 2, 4 channels, n_frames <= 512), else three passes (default);
                                 1 = always three passes (basis kernel, power tiles, activation kernel); 2 = same as 0      */
    BSS_OPT_BLOCKING_SYNC = 2,/* 1: host waits of this handle sleep on a blocking CUDA event instead of spinning in
                                 cudaStreamSynchronize (for many host threads / processes per node); 0 = spin (default) */
    BSS_OPT_ASYNC_INPUT = 4,  /* 1: bss_set_input / bss_set_input_waveform return as soon as the copy is queued on the handle's
                                 stream; the caller keeps the (pinned) buffer untouched until the next call that waits for
                                 the handle (bss_synchronize, bss_loss, ...).  0 = they return when the buffer may be
                                 reused (default).  A pipelined job queues the inputs of all its sub-batches back to back
                                 this way, so the host link never idles behind a sub-batch's STFT                          */
    BSS_OPT_ACT_CHUNKS = 1    /* number of bin chunks of the cross-bin (activation) reduction of the source model; 0 = chosen
                                 from batch size and machine (default).  The chunking fixes the summation order, so a single
                                 mixture given the chunk count of a batch reproduces the batch bit for bit                */
};
enum bss_info {
    BSS_INFO_IP_KERNEL = 0,   /* which form the last IP sweep took (values of BSS_OPT_IP_KERNEL; 4 = pairwise ip2_kernel) */
    BSS_INFO_GRAPH_REPLAYS = 1, /* CUDA-graph replays issued by bss_run / bss_run_record so far */
    BSS_INFO_LAUNCHES = 2,    /* same as bss_launch_count */
    BSS_INFO_ACT_CHUNKS = 3,  /* bin chunks the last activation update used (see BSS_OPT_ACT_CHUNKS) */
    BSS_INFO_SOURCE_MODEL = 4 /* form the last source-model update took: 1 = three passes, 2 = fused single pass */
};

typedef struct bss_config {
    int32_t method;        /* enum bss_method                                   */
    int32_t spatial;       /* enum bss_spatial                                  */
    int32_t normalize;     /* enum bss_normalize                                */
    int32_t partitioning;  /* GaussILRMA(partitioning=True)                     */
    int32_t algorithm;     /* enum bss_nmf_algorithm (NMF family only)          */
    int32_t n_batch;       /* B independent mixtures (reference: always 1)      */
    int32_t n_channels;    /* C (FastMNMF: M)                                   */
    int32_t n_sources;     /* N (== C except FastMNMF)                          */
    int32_t n_bins;        /* F                                                 */
    int32_t n_frames;      /* T                                                 */
    int32_t n_basis;       /* K                                                 */
    int32_t reference_id;  /* reference microphone of projection back           */
    int32_t device;        /* CUDA device ordinal                               */
    int32_t stream_priority; /* priority of the handle's own CUDA stream: 0 = default, negative = more urgent (clamped) */
    double domain;         /* 1 <= domain <= 2                                  */
    double nu;             /* degrees of freedom (tILRMA, tNMF)                 */
    double eps;            /* EPS = 1e-12        src/bss/ilrma.py:8             */
    double threshold;      /* THRESHOLD = 1e12   src/bss/ilrma.py:9             */
} bss_config;

/* ---- life cycle ------------------------------------------------------------------------- */
/* replaces Model.__init__ + the allocation half of Model._reset (src/bss/ilrma.py:183,50) */
int bss_create(const bss_config* cfg, bss_handle** out);
void bss_destroy(bss_handle* h);
/* text of the last error on this handle (h == NULL: last bss_create failure of this thread) */
const char* bss_last_error(const bss_handle* h);
/* run all later work of this handle on `cuda_stream` (a cudaStream_t); NULL = the handle's own stream */
int bss_set_stream(bss_handle* h, void* cuda_stream);
int bss_synchronize(bss_handle* h);

/* ---- data in / out ---------------------------------------------------------------------- */
/* `self.input = input` of Model.__call__ (src/bss/ilrma.py:210): x is (B,C,F,T) complex, host.
 * Also precomputes the plain spatial covariance mean_t x x^H used by the algebraic forms of
 * power normalisation and projection back. */
int bss_set_input(bss_handle* h, const void* x, int dtype);
/* the same from the time-domain mixture: x is (B,C,n_samples) float32/float64, or int16 PCM (scaled by 1/32768), on the host; the STFT of
 * src/transform/stft.py:4-8 (scipy.signal.stft: zero boundary extension, tail padding, `window`, onesided,
 * divided by sum(window)) is computed on the device straight into the bin tiles.  The handle must have
 * n_bins = fft_size/2 + 1 and n_frames = bss_stft_frames(n_samples, fft_size, hop_size); fft_size a power of two. */
int bss_set_input_waveform(bss_handle* h, const void* x, int dtype, int n_samples, int fft_size, int hop_size,
                           const double* window);
/* the state-copy half of Model._reset (src/bss/ilrma.py:67-104) and attribute assignment */
int bss_set_state(bss_handle* h, int which, const void* src, int dtype);
/* attribute reads (callbacks, tests) */
int bss_get_state(bss_handle* h, int which, void* dst, int dtype);
/* Model._reset defaults that need no host data: W = I (src/bss/ilrma.py:67-69),
 * FastMNMF Q = I, G = 1e-2 / 1 (src/bss/mnmf.py:660-663), estimation = separate(X, W) */
int bss_reset_spatial(bss_handle* h);

/* ---- the update loop -------------------------------------------------------------------- */
/* IP2 pair of the next update (src/bss/ilrma.py:635-646); advanced by the caller exactly as
 * the reference's __call__ does, not by bss_update_once.  (-1, -1) clears the pair (`update_pair = None`): the next
 * bss_run starts the schedule at (0, 1), bss_update_once without a pair fails with BSS_ESTATE. */
int bss_set_update_pair(bss_handle* h, int m, int n);
/* Model.update_once(): src/bss/ilrma.py:286 / :814, src/bss/iva.py:469 / :702,
 * src/bss/mnmf.py:737, src/algorithm/nmf.py:182,241,302,329,397,461-595 */
int bss_update_once(bss_handle* h);
/* GaussIDLMA only: bss_update_once is update_space_model (src/sss/idlma.py:175-210: covariances from BSS_STATE_VARIANCE,
 * gated IP sweep); bss_normalize is the normalisation tail of GaussIDLMA.update_once (src/sss/idlma.py:150-165,
 * 'projection-back': W <- diag(scale) W).  The source-model half (the DNN) stays with the caller. */
int bss_normalize(bss_handle* h);
/* n_iter x update_once without returning to the host in between (the `for idx in
 * range(iteration)` loop of __call__, src/bss/ilrma.py:233, with recordable_loss=False and no
 * callbacks); advances the IP2 pair schedule itself */
int bss_run(bss_handle* h, int n_iter);
/* the same loop with the loss recorded after every iteration, as the reference does by default
 * (recordable_loss=True: src/bss/ilrma.py:239-241; NMFbase.update: src/algorithm/nmf.py:169-174): the
 * per-iteration losses are reduced on the device and copied back once at the end.  loss[n_iter][B]. */
int bss_run_record(bss_handle* h, int n_iter, double* loss);
/* Model.compute_negative_loglikelihood() (src/bss/ilrma.py:648,993; src/bss/iva.py:604,783;
 * src/bss/mnmf.py:890) or the NMF criterion (src/algorithm/nmf.py:172-174); loss[B] */
int bss_loss(bss_handle* h, double* loss);
/* the tail of Model.__call__ (src/bss/ilrma.py:258-273): Y = separate(X, W), optionally scaled
 * by projection_back(Y, X[reference_id]) (src/algorithm/projection_back.py:3-34).
 * y is (B,N,F,T) complex on the host.  FastMNMF: multichannel Wiener filter (src/bss/mnmf.py:919). */
int bss_separate(bss_handle* h, void* y, int dtype, int apply_projection_back);
/* same, into a device buffer of complex64 (B,N,F,T) -- used to feed the NCCL gather of a
 * sharded batch without a host round trip */
int bss_separate_device(bss_handle* h, void* y_device, int apply_projection_back);
/* the same tail, continued to the time domain on the device: ISTFT (src/transform/stft.py:10-17) of the separated
 * estimates; y is (B,N,bss_istft_length(n_frames, fft_size, hop_size)) float32/float64 on the host */
int bss_separate_waveform(bss_handle* h, void* y, int dtype, int fft_size, int hop_size, const double* window,
                          int apply_projection_back);
/* per-handle switches and read-outs, see enum bss_option / enum bss_info */
int bss_set_option(bss_handle* h, int option, int value);
int bss_get_info(bss_handle* h, int what, int64_t* value);
/* the same with the time-domain estimates left on the device: y_device (B,N,bss_istft_length(...)) float32/float64; queued on
 * the handle's stream (bss_synchronize before another stream reads it) -- feeds the exchange of a sharded batch */
int bss_separate_waveform_device(bss_handle* h, void* y_device, int dtype, int fft_size, int hop_size, const double* window,
                                 int apply_projection_back);
/* The one collective of a sharded batch (BASELINE configs[4]: "NVLink gather only at end"): every rank contributes `bytes`
 * bytes at send_device and receives rank r's contribution at recv_base_device + r * rank_stride_bytes (r = 0 .. n_ranks-1,
 * its own included), on the handle's stream: one NCCL group of broadcasts, each contribution straight into its final place
 * (BSSGPU_GATHER=allgather in the environment: one ncclAllGather through a rank-major scratch buffer and one strided device
 * copy instead) -- so a sub-batch can be gathered into its final place of a (global batch, ...) buffer while later
 * sub-batches are still iterating.  `nccl_comm` is an
 * ncclComm_t of n_ranks ranks created by the caller (libnccl.so.2 is resolved at run time, the library does not link it).
 * Every rank must call it in the same order.  No reference counterpart (the reference has no batch axis). */
int bss_gather_outputs(bss_handle* h, void* nccl_comm, int n_ranks, int rank, const void* send_device, void* recv_base_device,
                       size_t bytes, size_t rank_stride_bytes);
/* The same gather without a collective kernel: every rank pushes its contribution into the result buffers of all its peers
 * with device-to-device copies over NVLink (copy engines: no SM is taken from the update loops of the other sub-batches, which
 * is what a gather that overlaps them needs -- an NCCL kernel beside them measured no gain, profiles/round2_scaling.md).
 * The result buffers are allocated by bss_peer_alloc (cudaMalloc + CUDA IPC handle, 64 bytes, which the caller hands to the
 * other processes), mapped by bss_peer_open in every peer, and all have the same layout: bss_push_outputs copies `bytes`
 * bytes from send_device to peer_bases[r] + dst_offset_bytes for every r != rank on the handle's stream.  The caller
 * synchronises its streams and then all ranks (a barrier) before anybody reads. */
int bss_peer_alloc(int device, size_t bytes, void** dptr, void* ipc_handle_64);
int bss_peer_open(int device, const void* ipc_handle_64, void** dptr);
int bss_peer_close(int device, void* dptr);
int bss_peer_free(int device, void* dptr);
int bss_push_outputs(bss_handle* h, int n_ranks, int rank, void* const* peer_bases, const void* send_device, size_t dst_offset_bytes,
                     size_t bytes);
/* ISS keeps no filter: W = Y X^H (X X^H)^-1 (src/bss/ilrma.py:167-173); result is readable as
 * BSS_STATE_DEMIX_FILTER afterwards */
int bss_compute_demix_filter(bss_handle* h);

/* ---- stateless primitives (one launch each; used by the parity tests and ncu) ----------- */
/* U[n,f] = mean_t x_ft x_ft^H / r[n,f,t]   src/bss/ilrma.py:503-511, src/bss/iva.py:491-499,
 * src/bss/mnmf.py:875.  x (C,F,T) complex128, r (N,F,T) float64 (already floored),
 * u (N,F,C,C) complex128. */
int bss_weighted_covariance(int device, int n_channels, int n_weights, int n_bins, int n_frames,
                            const void* x, const double* r, void* u);
/* Gauss-Seidel iterative-projection sweep over all rows   src/bss/ilrma.py:512-530 (floor_den = 0),
 * src/bss/mnmf.py:872-886 (floor_den = 1).  w (F,N,C) complex128 in/out, u (N,F,C,C) complex128,
 * gate (N,F) int32 out (may be NULL). */
int bss_ip_update(int device, int n_channels, int n_bins, void* w, const void* u, int32_t* gate,
                  double threshold, int floor_den, double eps);
/* scale (N,F) complex128 = projection_back(Y = W X, X[reference_id]) computed from W and the
 * plain covariance of x   src/algorithm/projection_back.py:12-21 */
int bss_projection_back_scale(int device, int n_channels, int n_bins, int n_frames, const void* x,
                              const void* w, int reference_id, void* scale);

/* Per-bin least-squares map  M_f = A_f B_f^H (B_f B_f^H)^-1  for arbitrary arrays:
 *   projection_back(Y, reference)          src/algorithm/projection_back.py:12-21 (2-D reference: n_rows_a = 1) and :25-32
 *                                          (3-D reference: n_rows_a = n_channels):  a = reference, b = Y, out = scale
 *   compute_demix_filter(estimation, input) src/bss/ilrma.py:167-173, src/bss/iva.py:119-125: a = Y, b = X, out[n][c][f] = W[f][n][c]
 * a (n_rows_a,F,T) complex128, b (n_rows_b,F,T) complex128, out (n_rows_a,n_rows_b,F) complex128; 1 <= rows <= 8.
 * An exactly singular B_f B_f^H returns BSS_ESINGULAR (np.linalg.inv raises LinAlgError there). */
int bss_least_squares_map(int device, int n_rows_a, int n_rows_b, int n_bins, int n_frames, const void* a, const void* b,
                          void* out);

/* Model.separate(input, demix_filter)   src/bss/ilrma.py:153-165, src/bss/iva.py:105-117.
 * x (C,F,T) complex128, w (F,C,C) complex128, y (C,F,T) complex128; `flags` is reserved (0). */
int bss_demix(int device, int n_channels, int n_bins, int n_frames, int flags, const void* x, const void* w,
              void* y);

/* ---- STFT feed (src/transform/stft.py:4-17 == scipy.signal.stft / istft with window=window_fn) ------------- */
/* number of frames of stft(x[n_samples]) and length of istft(Z[..., n_frames]) (host arithmetic only) */
int bss_stft_frames(int n_samples, int fft_size, int hop_size);
int bss_istft_length(int n_frames, int fft_size, int hop_size);
/* x (n_signals, n_samples) float64, window (fft_size) float64 -> out (n_signals, fft_size/2+1, n_frames) complex128 */
int bss_stft(int device, int n_signals, int n_samples, int fft_size, int hop_size, const double* window, const double* x,
             void* out);
/* z (n_signals, fft_size/2+1, n_frames) complex128 -> out (n_signals, bss_istft_length(...)) float64 */
int bss_istft(int device, int n_signals, int n_frames, int fft_size, int hop_size, const double* window, const void* z,
              double* out);

/* ---- measurement helpers ---------------------------------------------------------------- */
/* CUDA-event timing on the handle's stream: begin/end bracket a region, elapsed in ms */
int bss_timer_begin(bss_handle* h);
int bss_timer_end(bss_handle* h, float* elapsed_ms);
/* time `repeat` launches of the covariance-accumulate kernel alone on the current state;
 * returns the mean launch duration in ms (CUDA events on the handle's stream) */
int bss_time_covariance(bss_handle* h, int repeat, float* mean_ms);
/* number of kernel launches issued by this handle so far */
int64_t bss_launch_count(const bss_handle* h);
/* raw device buffers for zero-copy interop (documented layouts) */
int bss_device_buffer(bss_handle* h, int which, void** dptr, size_t* bytes);
const char* bss_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BSSGPU_H */
