"""grayskull_b200 -- B200-native (sm_100a) implementation of the zserge/grayskull stencil and
sliding-window hot path, behind the reference's gs_* C API.

The product is `libgrayskull_b200.so` (hand-written CUDA + an extern "C" shim, see include/*.h).
This package is the thin host-side mirror used by the tests and bench.py:

    grayskull_b200.lib()        the loaded C ABI (ctypes); raises if the library is not built
    grayskull_b200.api          gs_* on numpy arrays (host pointers), *_batch on torch CUDA tensors
    grayskull_b200.shard        frame-batch sharding across ranks (torch.distributed)
"""
from ._lib import lib, Image, Keypoint, Rect, Cascade, KP_DTYPE, RECT_DTYPE, MATCH_DTYPE, BLOB_DTYPE, load_cascade  # noqa: F401
