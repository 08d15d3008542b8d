"""Run on a GPU box: dumps the three curand_uniform() values per trial after curand_init(0, trial, 0) — the sampling
sequence of the reference's ransacEstimateModelKernel (/root/reference/src/cuda/cuda_ransac.cu:1154-1161) — through the
library's own table kernel (which makes exactly that curand call).  -> tests/golden/curand_xorwow_seed0.npy"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bundletrack_b200.matcher import Ransac
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden"
os.makedirs(out, exist_ok=True)
dev = torch.device("cuda:0")
r = Ransac(max_pairs=1, max_pts=16, max_trials=4096)
A = torch.rand((8, 4), device=dev)
r.ransac_pairs([A], [A.clone()], 4096, 0.01)
u3, _ = r.debug(4096, 1)
np.save(os.path.join(out, "curand_xorwow_seed0.npy"), u3)
print(u3[:3])
