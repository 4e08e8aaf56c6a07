// Single-channel NMF multiplicative updates: src/algorithm/nmf.py:150-595.
#include "methods.h"

int nmf_allocate(bss_handle* h) { return bss_fail(h, BSS_EUNSUPPORTED, "NMF is not implemented on the GPU path yet"); }
int nmf_update_once(bss_handle* h) { return bss_fail(h, BSS_EUNSUPPORTED, "NMF is not implemented on the GPU path yet"); }
int nmf_loss(bss_handle* h) { return bss_fail(h, BSS_EUNSUPPORTED, "NMF is not implemented on the GPU path yet"); }
