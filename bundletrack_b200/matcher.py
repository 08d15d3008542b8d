"""Host-side mirror of the matcher boundary (SiftManager::findCorresbyNN's two knnMatch calls,
/root/reference/src/FeatureManager.cpp:271-273), batched over frame pairs, on top of the C-ABI."""
from __future__ import annotations

import ctypes
from typing import List, Sequence, Tuple

from . import _lib


class KnnMatcher:
    """knnMatch(query, train, k) for many (A, B) pairs at once; ONE tensor-core contraction per pair serves both directions."""

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

    # ---- persistent descriptor pool: a frame's descriptors are converted once (Lfnet::detectFeature, FeatureManager.cpp:907) ----
    def pool_reserve(self, n_slots: int):
        _lib.check(self.lib.bt_desc_pool_reserve(self.ctx, ctypes.c_int(n_slots)), "bt_desc_pool_reserve")

    def pool_store(self, slot: int, desc: "torch.Tensor"):
        v = _lib.DescView()
        v.dev, v.n, v.dim, v.pitch_bytes = desc.data_ptr(), desc.shape[0], desc.shape[1], (desc.stride(0) * 4 if desc.shape[0] > 1 else desc.shape[1] * 4)
        _lib.check(self.lib.bt_desc_pool_store(self.ctx, ctypes.c_int(slot), ctypes.byref(v), self.stream), "bt_desc_pool_store")

    def knn_match_slots(self, slot_pairs: Sequence[Tuple[int, int]], sizes: Sequence[Tuple[int, int]], k: int = 5, device=None):
        """slot_pairs: (slotA, slotB) per pair; sizes: (nA, nB) per pair (what was stored).  Same outputs as knn_match_pairs."""
        import torch
        n = len(slot_pairs)
        sa = (ctypes.c_int32 * n)(*[int(a) for a, _ in slot_pairs])
        sb = (ctypes.c_int32 * n)(*[int(b) for _, b in slot_pairs])
        na, nb = sum(a for a, _ in sizes), sum(b for _, b in sizes)
        dev = device or torch.device("cuda", torch.cuda.current_device())
        idxAB = torch.empty((max(na, 1), k), dtype=torch.int32, device=dev); distAB = torch.empty((max(na, 1), k), dtype=torch.float32, device=dev)
        idxBA = torch.empty((max(nb, 1), k), dtype=torch.int32, device=dev); distBA = torch.empty((max(nb, 1), k), dtype=torch.float32, device=dev)
        _lib.check(self.lib.bt_knn_match_slots(self.ctx, ctypes.c_int(n), sa, sb, ctypes.c_int(k), ctypes.c_void_p(idxAB.data_ptr()), ctypes.c_void_p(distAB.data_ptr()),
                                               ctypes.c_void_p(idxBA.data_ptr()), ctypes.c_void_p(distBA.data_ptr()), self.stream), "bt_knn_match_slots")
        oAB, oBA, ia, ib = [], [], 0, 0
        for a, b in sizes:
            oAB.append((idxAB[ia:ia + a], distAB[ia:ia + a])); oBA.append((idxBA[ib:ib + b], distBA[ib:ib + b]))
            ia += a; ib += b
        return [x[0] for x in oAB], [x[1] for x in oAB], [x[0] for x in oBA], [x[1] for x in oBA]

    def force_fallback(self, every_nth: int):
        """Test knob: every n-th query row goes through the exact brute-force kernel (0 = off)."""
        _lib.check(self.lib.bt_knn_debug_force_fallback(self.ctx, ctypes.c_int(every_nth)), "bt_knn_debug_force_fallback")

    def enable_timing(self, on=True):
        _lib.check(self.lib.bt_knn_enable_timing(self.ctx, ctypes.c_int(1 if on else 0)), "bt_knn_enable_timing")

    def timing(self):
        ms = (ctypes.c_float * 4)()
        info = (ctypes.c_int * 3)()
        _lib.check(self.lib.bt_knn_get_timing(self.ctx, ms, info), "bt_knn_get_timing")
        return {"prep_ms": ms[0], "tc_ms": ms[1], "rerank_ms": ms[2], "fallback_ms": ms[3], "units": info[0], "rows": info[1], "fallback_rows": info[2]}


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


class MatchPipeline:
    """The matcher half of the hot path, device-resident: SiftManager::findCorres for a batch of frame pairs
    (kNN both directions -> pruneMatches -> collectMutualMatches -> runRansacBetween) ending in the EntryJ records
    Bundler::optimizeGPU would build (/root/reference/src/FeatureManager.cpp:173-368,561-741; src/Bundler.cpp:298-324)."""

    def __init__(self, yml=None, device: int = 0, max_pairs: int = 64, max_feats: int = 4096, dim: int = 256, stream: int = 0):
        import math
        from .config import load_yml
        self.lib = _lib.load()
        self.stream = ctypes.c_void_p(stream)
        self.ctx = ctypes.c_void_p()
        _lib.check(self.lib.bt_ctx_create(ctypes.byref(self.ctx), ctypes.c_int(device)), "bt_ctx_create")
        y = load_yml(yml)
        fc, rs = y["feature_corres"], y["ransac"]
        self.prune = _lib.PruneParams(float(fc["max_dist_no_neighbor"]), math.cos(float(fc["max_normal_no_neighbor"]) / 180.0 * math.pi),
                                      float(fc["max_dist_neighbor"]), math.cos(float(fc["max_normal_neighbor"]) / 180.0 * math.pi))
        self.ransac_trials = int(rs["max_iter"])
        self.ransac_inlier_dist = float(rs["inlier_dist"])
        _lib.check(self.lib.bt_pipeline_reserve(self.ctx, ctypes.c_int(max_pairs), ctypes.c_int(max_feats), ctypes.c_int(dim), ctypes.c_int(max(self.ransac_trials, 1))), "bt_pipeline_reserve")
        self.max_pairs, self.max_feats = max_pairs, max_feats

    def close(self):
        if self.ctx:
            self.lib.bt_ctx_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()

    @staticmethod
    def _frame(fr) -> "_lib.MatchFrame":
        """fr: dict with kpts [n,2] cuda f32, depth [H,W] cuda f32, normal [H,W,4] cuda f32, pose [4,4] numpy, id, window_index."""
        import numpy as np
        m = _lib.MatchFrame()
        m.kpts_dev, m.n = fr["kpts"].data_ptr(), fr["kpts"].shape[0]
        m.depth_dev, m.normal_dev = fr["depth"].data_ptr(), fr["normal"].data_ptr()
        P = np.ascontiguousarray(fr["pose"], np.float32).reshape(16)
        for k in range(16):
            m.pose[k] = float(P[k])
        m.frame_id, m.window_index = int(fr["id"]), int(fr.get("window_index", fr["id"]))
        return m

    def _views(self, pairs):
        n = len(pairs)
        A = (_lib.MatchFrame * n)()
        B = (_lib.MatchFrame * n)()
        dA = (_lib.DescView * n)()
        dB = (_lib.DescView * n)()
        for i, (fa, fb) in enumerate(pairs):
            A[i], B[i] = self._frame(fa), self._frame(fb)
            for view, t in ((dA[i], fa.get("desc")), (dB[i], fb.get("desc"))):
                if t is not None:      # (frames whose descriptors live in the pool need no view)
                    view.dev, view.n, view.dim, view.pitch_bytes = t.data_ptr(), t.shape[0], t.shape[1], (t.stride(0) * 4 if t.shape[0] > 1 else t.shape[1] * 4)
        return A, B, dA, dB

    def prune_mutual(self, pairs, idxAB, idxBA, H, W, K, k=5):
        """pairs: list of (frameA newer, frameB older) dicts; idxAB/idxBA: concatenated [sum n, k] int32 cuda kNN indices.
        Returns list of [m,10] float32 cuda tensors (uA,vA,uB,vB,ptA_cam,ptB_cam)."""
        import torch
        A, B, _, _ = self._views(pairs)
        n = len(pairs)
        cap = sum(fa["kpts"].shape[0] + fb["kpts"].shape[0] for fa, fb in pairs)
        dev = pairs[0][0]["kpts"].device
        corr = torch.empty((max(cap, 1), 10), dtype=torch.float32, device=dev)
        cnt = torch.empty(n, dtype=torch.int32, device=dev)
        _lib.check(self.lib.bt_prune_mutual_pairs(self.ctx, ctypes.c_int(n), A, B, ctypes.c_int(H), ctypes.c_int(W),
                                                  ctypes.c_float(K[0]), ctypes.c_float(K[1]), ctypes.c_float(K[2]), ctypes.c_float(K[3]),
                                                  ctypes.c_void_p(idxAB.data_ptr()), ctypes.c_void_p(idxBA.data_ptr()), ctypes.c_int(k), ctypes.byref(self.prune),
                                                  ctypes.c_void_p(corr.data_ptr()), ctypes.c_void_p(cnt.data_ptr()), self.stream), "bt_prune_mutual_pairs")
        ch = cnt.cpu().numpy()
        out, o = [], 0
        for p, (fa, fb) in enumerate(pairs):
            out.append(corr[o:o + int(ch[p])])
            o += fa["kpts"].shape[0] + fb["kpts"].shape[0]
        return out

    def pool_reserve(self, n_slots: int):
        _lib.check(self.lib.bt_desc_pool_reserve(self.ctx, ctypes.c_int(n_slots)), "bt_desc_pool_reserve")

    def pool_store(self, slot: int, desc: "torch.Tensor"):
        v = _lib.DescView()
        v.dev, v.n, v.dim, v.pitch_bytes = desc.data_ptr(), desc.shape[0], desc.shape[1], (desc.stride(0) * 4 if desc.shape[0] > 1 else desc.shape[1] * 4)
        _lib.check(self.lib.bt_desc_pool_store(self.ctx, ctypes.c_int(slot), ctypes.byref(v), self.stream), "bt_desc_pool_store")

    def match_pairs(self, pairs, H, W, K, seed: int = 0, capacity: int = 0, keep_on_device: bool = False, slots=None):
        """Fused pipeline.  slots: optional list of (slotA, slotB) naming descriptor sets stored with pool_store (the frames' "desc" entries
        are then not read).  Returns (entries [total] EntryJ structured numpy array, n_entry [n_pairs], entry_off [n_pairs]).
        keep_on_device=True: the entries stay on the GPU - returns (device int32 tensor [cap, 8], n_entry, entry_off) and only the
        two small count arrays are read back; hand them to SolveWindow(corr_dev=..., blocks=...) (see `solver_blocks`)."""
        import numpy as np
        import torch
        from .synth import ENTRYJ_DTYPE
        A, B, dA, dB = self._views(pairs)
        n = len(pairs)
        cap = capacity or sum(fa["kpts"].shape[0] + fb["kpts"].shape[0] for fa, fb in pairs)
        dev = pairs[0][0]["kpts"].device
        ent = torch.empty((max(cap, 1), 8), dtype=torch.int32, device=dev)     # 32-byte EntryJ records
        n_ent = torch.empty(n, dtype=torch.int32, device=dev)
        off = torch.empty(n, dtype=torch.int32, device=dev)
        tot = torch.zeros(4, dtype=torch.int32, device=dev)
        tail = (ctypes.c_int(H), ctypes.c_int(W), ctypes.c_float(K[0]), ctypes.c_float(K[1]), ctypes.c_float(K[2]), ctypes.c_float(K[3]),
                ctypes.byref(self.prune), ctypes.c_int(self.ransac_trials), ctypes.c_float(self.ransac_inlier_dist), ctypes.c_uint64(seed),
                ctypes.c_void_p(ent.data_ptr()), ctypes.c_int(cap), ctypes.c_void_p(n_ent.data_ptr()), ctypes.c_void_p(off.data_ptr()),
                ctypes.c_void_p(tot.data_ptr()), self.stream)
        if slots is None:
            _lib.check(self.lib.bt_match_pairs(self.ctx, ctypes.c_int(n), A, B, dA, dB, *tail), "bt_match_pairs")
        else:
            sa = (ctypes.c_int32 * n)(*[int(a) for a, _ in slots]); sb = (ctypes.c_int32 * n)(*[int(b) for _, b in slots])
            _lib.check(self.lib.bt_match_pairs_pool(self.ctx, ctypes.c_int(n), A, B, sa, sb, *tail), "bt_match_pairs_pool")
        self.last = (ent, n_ent, off, tot)
        if keep_on_device:
            cnt = torch.stack([n_ent, off]).cpu().numpy()      # one small D2H (2 x n_pairs ints), no entry leaves the device
            return ent, cnt[0], cnt[1]
        total = int(tot[0].item())
        host = ent[:total].cpu().numpy().view(np.uint8).reshape(-1).view(ENTRYJ_DTYPE)
        return host, n_ent.cpu().numpy(), off.cpu().numpy()

    @staticmethod
    def solver_blocks(pairs, n_entry, entry_off):
        """blocks=(off, n, i, j) for SolveWindow from match_pairs(keep_on_device=True): block p carries (imgIdx_i, imgIdx_j) =
        (older frame B's window index, newer frame A's), exactly what k_emit_entryj writes into the entries."""
        import numpy as np
        bi = np.array([fb["window_index"] for _, fb in pairs], np.uint32)
        bj = np.array([fa["window_index"] for fa, _ in pairs], np.uint32)
        return np.asarray(entry_off, np.int32), np.asarray(n_entry, np.int32), bi, bj


class FindCorres:
    """`SiftManager::findCorres` (/root/reference/src/FeatureManager.cpp:173-242) on the device-resident chain: per pair (A newer)
    kNN both ways -> prune -> mutual union -> [non-neighbour: map-point propagation] -> RANSAC -> EntryJ, with the `_matches` cache
    (a pair is matched once), the map-point bookkeeping (bt_tracks_*) and the Frame::FAIL outcome.  Descriptors live in the pool."""

    def __init__(self, pipeline: "MatchPipeline", max_pairs_cached: int = 512, max_entries_per_pair: int = 0):
        from .policy import Tracks
        self.mp = pipeline
        self.lib = pipeline.lib
        self.ctx = pipeline.ctx
        self.tracks = Tracks()
        self.slot_entries = max_entries_per_pair or 3 * pipeline.max_feats
        _lib.check(self.lib.bt_match_cache_reserve(self.ctx, ctypes.c_int(max_pairs_cached), ctypes.c_int(self.slot_entries)), "bt_match_cache_reserve")
        self.failed = set()          # ids of frames the reference would mark Frame::FAIL

    def has(self, id_a: int, id_b: int) -> bool:
        return bool(self.lib.bt_match_cache_has(self.ctx, ctypes.c_int(id_a), ctypes.c_int(id_b)))

    def find_corres(self, pairs, slots, H, W, K, seed: int = 0):
        """pairs: list of (frameA newer, frameB older) dicts as for MatchPipeline.match_pairs; slots: their descriptor-pool slots.
        Pairs already in the cache are skipped (FeatureManager.cpp:176).  The pairs of ONE call must not depend on each other's map
        points (the reference matches them one after the other): hand a new frame's neighbour pair first, then its other pairs.
        Returns {(idA, idB): (n_entries, status)} of the pairs processed in this call."""
        import numpy as np
        import torch
        todo = [(k, p) for k, p in enumerate(pairs) if not self.has(p[0]["id"], p[1]["id"])]
        if not todo:
            return {}
        prs = [p for _, p in todo]
        sl = [slots[k] for k, _ in todo]
        n = len(prs)
        A, B, _, _ = self.mp._views(prs)      # (descriptor views unused: the pool is named by slot)
        extra = (_lib.MatchExtra * n)()
        keep = []
        for i, (a, b) in enumerate(prs):
            if abs(a["id"] - b["id"]) != 1:
                uv = np.ascontiguousarray(self.tracks.propagate(a["id"], b["id"], np.zeros((0, 4), np.float32)), np.float32)
                keep.append(uv)
                extra[i].uv, extra[i].n = (uv.ctypes.data if len(uv) else None), len(uv)
        cap = sum(a["kpts"].shape[0] + b["kpts"].shape[0] + extra[i].n for i, (a, b) in enumerate(prs))
        dev = prs[0][0]["kpts"].device
        ent = torch.empty((max(cap, 1), 8), dtype=torch.int32, device=dev)
        uvo = torch.empty((max(cap, 1), 4), dtype=torch.float32, device=dev)
        meta = torch.zeros((3, n), dtype=torch.int32, device=dev)              # n_entry, entry_off, status
        tot = torch.zeros(4, dtype=torch.int32, device=dev)
        sa = (ctypes.c_int32 * n)(*[int(a) for a, _ in sl]); sb = (ctypes.c_int32 * n)(*[int(b) for _, b in sl])
        mp = self.mp
        _lib.check(self.lib.bt_match_pairs_ex(self.ctx, ctypes.c_int(n), A, B, None, None, sa, sb, ctypes.c_int(H), ctypes.c_int(W),
                                              ctypes.c_float(K[0]), ctypes.c_float(K[1]), ctypes.c_float(K[2]), ctypes.c_float(K[3]), ctypes.byref(mp.prune),
                                              ctypes.c_int(mp.ransac_trials), ctypes.c_float(mp.ransac_inlier_dist), ctypes.c_uint64(seed), extra,
                                              ctypes.c_void_p(ent.data_ptr()), ctypes.c_int(cap), ctypes.c_void_p(meta[0].data_ptr()), ctypes.c_void_p(meta[1].data_ptr()),
                                              ctypes.c_void_p(tot.data_ptr()), ctypes.c_void_p(meta[2].data_ptr()), ctypes.c_void_p(uvo.data_ptr()), mp.stream), "bt_match_pairs_ex")
        m = meta.cpu().numpy()                                                  # 3 x n ints: the only per-call read-back besides the inliers' pixel coordinates
        uv_h = uvo[: int((m[0] + m[1]).max())].cpu().numpy() if m[0].sum() else np.zeros((0, 4), np.float32)
        ida = (ctypes.c_int32 * n)(*[int(a["id"]) for a, _ in prs]); idb = (ctypes.c_int32 * n)(*[int(b["id"]) for _, b in prs])
        n_h = np.ascontiguousarray(m[0]); off_h = np.ascontiguousarray(m[1]); st_h = np.ascontiguousarray(m[2])
        _lib.check(self.lib.bt_match_cache_put(self.ctx, ctypes.c_int(n), ida, idb, ctypes.c_void_p(ent.data_ptr()), n_h.ctypes.data_as(ctypes.c_void_p),
                                               off_h.ctypes.data_as(ctypes.c_void_p), st_h.ctypes.data_as(ctypes.c_void_p), mp.stream), "bt_match_cache_put")
        out = {}
        for i, (a, b) in enumerate(prs):
            if n_h[i] > 0:        # updateFramePairMapPoints on the surviving matches (all inliers)
                self.tracks.update_pair(a["id"], b["id"], uv_h[off_h[i]:off_h[i] + n_h[i]])
            if st_h[i] == 2:
                self.failed.add(a["id"])
            out[(a["id"], b["id"])] = (int(n_h[i]), int(st_h[i]))
        self._keep = (ent, uvo, keep)
        return out

    def window_entries(self, frames, device):
        """The EntryJ list Bundler::optimizeGPU builds for the window `frames` (oldest first; window index = position), straight from
        the cache: returns (device EntryJ tensor, blocks) for SolveWindow(corr_dev=..., blocks=...)."""
        import numpy as np
        import torch
        prs = [(frames[j]["id"], frames[i]["id"], i, j) for i in range(len(frames)) for j in range(i + 1, len(frames))]
        n = len(prs)
        ida = (ctypes.c_int32 * n)(*[p[0] for p in prs]); idb = (ctypes.c_int32 * n)(*[p[1] for p in prs])
        wi = np.array([p[2] for p in prs], np.uint32); wj = np.array([p[3] for p in prs], np.uint32)
        cap = n * self.slot_entries
        dst = torch.empty((max(min(cap, 1 << 22), 1), 8), dtype=torch.int32, device=device)
        off = np.zeros(n, np.int32); cnt = np.zeros(n, np.int32)
        _lib.check(self.lib.bt_match_cache_gather(self.ctx, ctypes.c_int(n), ida, idb, wi.ctypes.data_as(ctypes.c_void_p), wj.ctypes.data_as(ctypes.c_void_p),
                                                  ctypes.c_void_p(dst.data_ptr()), ctypes.c_int(dst.shape[0]), off.ctypes.data_as(ctypes.c_void_p), cnt.ctypes.data_as(ctypes.c_void_p),
                                                  self.mp.stream), "bt_match_cache_gather")
        return dst, (off, cnt, wi, wj)

    def forget_frame(self, frame_id: int):
        _lib.check(self.lib.bt_match_cache_forget_frame(self.ctx, ctypes.c_int(frame_id)), "bt_match_cache_forget_frame")
        self.tracks.forget_frame(frame_id)
