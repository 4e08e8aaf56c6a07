"""Index helpers of src/utils/utils_linalg.py that the update loop uses."""
import numpy as np


def parallel_sort(x, order, axis=-2):
    """Batched gather along `axis` by integer `order` (src/utils/utils_linalg.py:33-52).

    Host-side integer index arithmetic, used by callers that post-process eigen-decompositions
    (the IP2 update itself orders its eigenvectors on the device).  Results are bit-identical to the
    reference for any input.
    """
    x = np.asarray(x)
    order = np.asarray(order)
    if axis < 0:
        axis = x.ndim + axis
    lead = x.shape[:axis]
    n_elem = x.shape[axis]
    tail = x.shape[axis + 1:]
    n_pick = order.shape[-1]
    flat = x.reshape(-1, *tail)
    base = np.repeat(n_elem * np.arange(int(np.prod(lead, dtype=np.int64))), n_pick)
    picked = flat[order.reshape(-1) + base]
    return picked.reshape(*lead, n_pick, *tail)
