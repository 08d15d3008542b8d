"""One fused matcher call (45 pairs x 2000 feats: kNN -> prune -> mutual -> RANSAC -> EntryJ) for launch-list captures."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bundletrack_b200 import synth
from bundletrack_b200.matcher import MatchPipeline
dev = torch.device("cuda:0")
w = synth.make_window(77, n_frames=10, n_corr=10)
fr = synth.make_feature_frames(w, 2000, seed=77, n_surface=12000)
devf = [{"kpts": torch.from_numpy(fr[k]["kpts"]).to(dev), "desc": torch.from_numpy(fr[k]["desc"]).to(dev), "depth": torch.from_numpy(w.depth[k]).to(dev),
         "normal": torch.from_numpy(w.normal[k]).to(dev), "pose": w.poses_init[k], "id": k, "window_index": k} for k in range(10)]
mp = MatchPipeline(None, max_pairs=48, max_feats=2048)
prs = [(devf[j], devf[i]) for i in range(10) for j in range(i + 1, 10)]
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ent, n_ent, off = mp.match_pairs(prs, w.H, w.W, w.K)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    ent, n_ent, off = mp.match_pairs(prs, w.H, w.W, w.K)
torch.cuda.synchronize()
print("call ms", (time.perf_counter() - t0) / 10 * 1e3, "entries", int(np.sum(n_ent)), "per pair", n_ent.tolist()[:10])
