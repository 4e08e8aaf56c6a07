// Per-method orchestration (which kernels one update_once launches, in which order).
#pragma once
#include "handle.h"

// determined BSS (ILRMA / AuxIVA): methods_bss.cu
int bss_allocate(bss_handle* h);
int bss_reset_filter(bss_handle* h);
int ilrma_update_once(bss_handle* h);
int tilrma_update_once(bss_handle* h);
int auxiva_update_once(bss_handle* h);
int idlma_update_once(bss_handle* h);                     // GaussIDLMA.update_space_model
int idlma_normalize(bss_handle* h);                       // normalisation tail of GaussIDLMA.update_once
int idlma_set_variance(bss_handle* h, const double* r);   // host (B,N,F,T) -> iw = 1 / max(r, eps)
int bss_loss_device(bss_handle* h);                       // result in lossbuf[B*F .. B*F+B)
int bss_separate_to(bss_handle* h, cf* out, int apply_pb); // out: device (B,N,F,T) complex64
int bss_filter_from_estimates(bss_handle* h);
int bss_covariance_only(bss_handle* h);
int bss_refresh_estimates(bss_handle* h);                 // Y <- W X (bin-major device buffer)

// GaussILRMA(partitioning=True): methods_part.cu
int ilrma_partitioned_update_once(bss_handle* h);
int ilrma_partitioned_loss(bss_handle* h);

// FastMNMF: methods_mnmf.cu
int mnmf_allocate(bss_handle* h);
int mnmf_reset(bss_handle* h);
int mnmf_update_once(bss_handle* h);
int mnmf_loss(bss_handle* h);
int mnmf_separate(bss_handle* h, cf* out);
int mnmf_covariance_only(bss_handle* h);

// Sawada IS-MNMF: methods_smnmf.cu
int smnmf_allocate(bss_handle* h);
int smnmf_reset(bss_handle* h);
int smnmf_update_once(bss_handle* h);
int smnmf_loss(bss_handle* h);
int smnmf_separate(bss_handle* h, cf* out);
int smnmf_set_state(bss_handle* h, int which, const void* src, int dtype);
int smnmf_get_state(bss_handle* h, int which, void* dst, int dtype);

// single-channel NMF: methods_nmf.cu
int nmf_allocate(bss_handle* h);
int nmf_update_once(bss_handle* h);
int nmf_run(bss_handle* h, int n_iter, double* loss_hist_device);
int nmf_loss(bss_handle* h);

// STFT feed: kernels_stft.cu
int stft_into_handle(bss_handle* h, const void* x, int dtype, int n_samples, int fft_size, int hop_size, const double* window);
int istft_from_device(bss_handle* h, const cf* z, int n_signals, int fft_size, int hop_size, const double* window, void* y, int dtype,
                      int y_on_device = 0);
