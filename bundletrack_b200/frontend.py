"""Host-side mirror of the reference's per-frame depth front end (`Frame::processDepth` + `Frame::depthToCloudAndNormals`,
/root/reference/src/Frame.cpp:152-233) on top of the C-ABI `bt_frames_preprocess`.  All arithmetic runs in
lib/libbundletrack_b200.so on the GPU; this file only marshals device pointers."""
from __future__ import annotations

import ctypes
from typing import Sequence

from . import _lib
from .config import depth_params


def _ptr(x) -> int:
    return int(x.data_ptr()) if hasattr(x, "data_ptr") else int(x)


class FrameFrontEnd:
    def __init__(self, yml=None, device: int = 0, stream: int = 0, ctx=None):
        self.lib = _lib.load()
        self.params = depth_params(yml)
        self.stream = ctypes.c_void_p(stream)
        self._own = ctx is None
        self.ctx = ctypes.c_void_p() if ctx is None else ctx
        if self._own:
            _lib.check(self.lib.bt_ctx_create(ctypes.byref(self.ctx), ctypes.c_int(device)), "bt_ctx_create")

    def close(self):
        if self._own and self.ctx:
            self.lib.bt_ctx_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def process(self, depth_raw: Sequence, H: int, W: int, K, depth_out: Sequence, normal_out: Sequence, xyz_out: Sequence = None):
        """depth_raw/depth_out: per-frame device float32 [H,W]; normal_out/xyz_out: device float32 [H,W,4].
        K = (fx, fy, cx, cy) or a 3x3 matrix.  Outputs are written in place; returns None."""
        import numpy as np
        Kn = np.asarray(K, np.float64)
        fx, fy, cx, cy = (Kn[0, 0], Kn[1, 1], Kn[0, 2], Kn[1, 2]) if Kn.ndim == 2 else Kn
        n = len(depth_raw)
        arr = lambda seq: (ctypes.c_void_p * n)(*[_ptr(t) for t in seq])
        a_in, a_out, a_n = arr(depth_raw), arr(depth_out), arr(normal_out)
        a_x = arr(xyz_out) if xyz_out is not None else None
        _lib.check(self.lib.bt_frames_preprocess(self.ctx, ctypes.c_int(n), a_in, ctypes.c_int(H), ctypes.c_int(W),
                                                 ctypes.c_float(fx), ctypes.c_float(fy), ctypes.c_float(cx), ctypes.c_float(cy),
                                                 ctypes.byref(self.params), a_out, a_x, a_n, self.stream), "bt_frames_preprocess")
