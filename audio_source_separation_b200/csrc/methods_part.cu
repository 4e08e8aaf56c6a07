// GaussILRMA with a shared (partitioned) basis: src/bss/ilrma.py:368-408, :313-320, :493-495.
#include "methods.h"

int ilrma_partitioned_update_once(bss_handle* h) {
    return bss_fail(h, BSS_EUNSUPPORTED, "partitioning=True is not implemented on the GPU path yet");
}

int ilrma_partitioned_loss(bss_handle* h) {
    return bss_fail(h, BSS_EUNSUPPORTED, "partitioning=True is not implemented on the GPU path yet");
}
