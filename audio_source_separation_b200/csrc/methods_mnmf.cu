// FastMNMF: src/bss/mnmf.py:637-946.
#include "methods.h"

int mnmf_allocate(bss_handle* h) { return bss_fail(h, BSS_EUNSUPPORTED, "FastMNMF is not implemented on the GPU path yet"); }
int mnmf_reset(bss_handle* h) { return bss_fail(h, BSS_EUNSUPPORTED, "FastMNMF is not implemented on the GPU path yet"); }
int mnmf_update_once(bss_handle* h) { return bss_fail(h, BSS_EUNSUPPORTED, "FastMNMF is not implemented on the GPU path yet"); }
int mnmf_loss(bss_handle* h) { return bss_fail(h, BSS_EUNSUPPORTED, "FastMNMF is not implemented on the GPU path yet"); }
int mnmf_separate(bss_handle* h, cf*) { return bss_fail(h, BSS_EUNSUPPORTED, "FastMNMF is not implemented on the GPU path yet"); }
