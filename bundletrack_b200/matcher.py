"""Host-side mirror of the matcher boundary (SiftManager::findCorresbyNN's two knnMatch calls,
/root/reference/src/FeatureManager.cpp:271-273), batched over frame pairs, on top of the C-ABI."""
from __future__ import annotations

import ctypes
from typing import List, Sequence, Tuple

from . import _lib


class KnnMatcher:
    """knnMatch(query, train, k) for many (A, B) pairs at once, both directions from one tensor-core pass per direction."""

    def __init__(self, device: int = 0, max_pairs: int = 64, max_feats: int = 4096, dim: int = 256, stream: int = 0, ctx=None):
        self.lib = _lib.load()
        self.stream = ctypes.c_void_p(stream)
        self._own = ctx is None
        self.ctx = ctypes.c_void_p() if ctx is None else ctx
        if self._own:
            _lib.check(self.lib.bt_ctx_create(ctypes.byref(self.ctx), ctypes.c_int(device)), "bt_ctx_create")
        _lib.check(self.lib.bt_matcher_reserve(self.ctx, ctypes.c_int(max_pairs), ctypes.c_int(max_feats), ctypes.c_int(dim)), "bt_matcher_reserve")
        self.dim = dim

    def close(self):
        if self._own and self.ctx:
            self.lib.bt_ctx_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def knn_match_pairs(self, pairs: Sequence[Tuple["torch.Tensor", "torch.Tensor"]], k: int = 5):
        """pairs: list of (A [nA,dim] fp32 cuda, B [nB,dim] fp32 cuda); rows may be pitched (stride(0) >= dim).
        Returns (idxAB, distAB, idxBA, distBA): lists (one entry per pair) of [n,k] int32 / float32 cuda tensors.
        distance = sqrt(sum (a-b)^2) ascending, ties -> lower train index (cv::DMatch semantics)."""
        import torch
        n = len(pairs)
        A = (_lib.DescView * n)()
        B = (_lib.DescView * n)()
        na = nb = 0
        for i, (a, b) in enumerate(pairs):
            for view, t in ((A[i], a), (B[i], b)):
                assert t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1
                view.dev, view.n, view.dim, view.pitch_bytes = t.data_ptr(), t.shape[0], t.shape[1], t.stride(0) * 4 if t.shape[0] > 1 else t.shape[1] * 4
            na += a.shape[0]
            nb += b.shape[0]
        dev = pairs[0][0].device
        idxAB = torch.empty((max(na, 1), k), dtype=torch.int32, device=dev)
        distAB = torch.empty((max(na, 1), k), dtype=torch.float32, device=dev)
        idxBA = torch.empty((max(nb, 1), k), dtype=torch.int32, device=dev)
        distBA = torch.empty((max(nb, 1), k), dtype=torch.float32, device=dev)
        _lib.check(self.lib.bt_knn_match_pairs(self.ctx, ctypes.c_int(n), A, B, ctypes.c_int(k),
                                               ctypes.c_void_p(idxAB.data_ptr()), ctypes.c_void_p(distAB.data_ptr()),
                                               ctypes.c_void_p(idxBA.data_ptr()), ctypes.c_void_p(distBA.data_ptr()), self.stream), "bt_knn_match_pairs")
        oAB, oBA, ia, ib = [], [], 0, 0
        for a, b in pairs:
            oAB.append((idxAB[ia:ia + a.shape[0]], distAB[ia:ia + a.shape[0]]))
            oBA.append((idxBA[ib:ib + b.shape[0]], distBA[ib:ib + b.shape[0]]))
            ia += a.shape[0]
            ib += b.shape[0]
        return [x[0] for x in oAB], [x[1] for x in oAB], [x[0] for x in oBA], [x[1] for x in oBA]

    def enable_timing(self, on=True):
        _lib.check(self.lib.bt_knn_enable_timing(self.ctx, ctypes.c_int(1 if on else 0)), "bt_knn_enable_timing")

    def timing(self):
        ms = (ctypes.c_float * 4)()
        info = (ctypes.c_int * 3)()
        _lib.check(self.lib.bt_knn_get_timing(self.ctx, ms, info), "bt_knn_get_timing")
        return {"prep_ms": ms[0], "tc_ms": ms[1], "rerank_ms": ms[2], "fallback_ms": ms[3], "items": info[0], "rows": info[1], "fallback_rows": info[2]}


class Ransac:
    """ransacMultiPairGPU (/root/reference/src/cuda/cuda_ransac.h:50) for a batch of pairs."""

    def __init__(self, device: int = 0, max_pairs: int = 64, max_pts: int = 8192, max_trials: int = 2000, stream: int = 0, ctx=None):
        self.lib = _lib.load()
        self.stream = ctypes.c_void_p(stream)
        self._own = ctx is None
        self.ctx = ctypes.c_void_p() if ctx is None else ctx
        if self._own:
            _lib.check(self.lib.bt_ctx_create(ctypes.byref(self.ctx), ctypes.c_int(device)), "bt_ctx_create")
        _lib.check(self.lib.bt_ransac_reserve(self.ctx, ctypes.c_int(max_pairs), ctypes.c_int(max_pts), ctypes.c_int(max_trials)), "bt_ransac_reserve")
        self.max_trials = max_trials

    def close(self):
        if self._own and self.ctx:
            self.lib.bt_ctx_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()

    def ransac_pairs(self, ptsA: List["torch.Tensor"], ptsB: List["torch.Tensor"], n_trials: int, dist_thresh: float, seed: int = 0):
        """ptsA[p], ptsB[p]: [n_p,4] float32 cuda (w ignored).  Returns list of int32 cuda tensors of inlier ids (ascending)."""
        import torch
        n = len(ptsA)
        pa = (ctypes.c_void_p * n)(*[t.data_ptr() for t in ptsA])
        pb = (ctypes.c_void_p * n)(*[t.data_ptr() for t in ptsB])
        cnt = (ctypes.c_int * n)(*[t.shape[0] for t in ptsA])
        total = sum(t.shape[0] for t in ptsA)
        dev = ptsA[0].device
        ids = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        nin = torch.empty(n, dtype=torch.int32, device=dev)
        _lib.check(self.lib.bt_ransac_pairs(self.ctx, ctypes.c_int(n), pa, pb, cnt, ctypes.c_int(n_trials), ctypes.c_float(dist_thresh), ctypes.c_uint64(seed),
                                            ctypes.c_void_p(ids.data_ptr()), ctypes.c_void_p(nin.data_ptr()), self.stream), "bt_ransac_pairs")
        nin_h = nin.cpu().numpy()
        out, o = [], 0
        for p in range(n):
            out.append(ids[o:o + int(nin_h[p])])
            o += ptsA[p].shape[0]
        return out

    def debug(self, n_trials: int, n_pairs: int):
        import numpy as np
        u3 = np.zeros((n_trials, 3), np.float32)
        bt = np.zeros(n_pairs, np.int32)
        _lib.check(self.lib.bt_ransac_debug(self.ctx, u3.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n_trials), bt.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n_pairs)), "bt_ransac_debug")
        return u3, bt
