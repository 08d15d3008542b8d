"""cfg2-shaped matcher calls (45 pairs x 2000 feats, descriptors in the pool) and one cfg5 pair (5000 x 5000) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bundletrack_b200 import synth
from bundletrack_b200.matcher import KnnMatcher
dev = torch.device("cuda:0")
m = KnnMatcher(max_pairs=48, max_feats=5120)
m.pool_reserve(12)
frames = [torch.from_numpy(synth.make_descriptors(100 + f, 2000, 8)[0]).to(dev) for f in range(10)]
for f in range(10):
    m.pool_store(f, frames[f])
idx = [(j, i) for i in range(10) for j in range(i + 1, 10)]
for _ in range(3):
    m.knn_match_slots(idx, [(2000, 2000)] * 45, device=dev)
torch.cuda.synchronize()
a, b, _, _ = synth.make_descriptors(5, 5000, 5000)
m.pool_store(10, torch.from_numpy(a).to(dev)); m.pool_store(11, torch.from_numpy(b).to(dev))
for _ in range(2):
    m.knn_match_slots([(10, 11)], [(5000, 5000)], device=dev)
torch.cuda.synchronize()
