"""One cfg2-shaped matcher call (45 pairs x 2000 feats) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bundletrack_b200 import synth
from bundletrack_b200.matcher import KnnMatcher
dev = torch.device("cuda:0")
m = KnnMatcher(max_pairs=48, max_feats=5120)
frames = [torch.from_numpy(synth.make_descriptors(100 + f, 2000, 8)[0]).to(dev) for f in range(10)]
pairs = [(frames[j], frames[i]) for i in range(10) for j in range(i + 1, 10)]
for _ in range(3):
    m.knn_match_pairs(pairs)
torch.cuda.synchronize()
a, b, _, _ = synth.make_descriptors(5, 5000, 5000)
ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
for _ in range(2):
    m.knn_match_pairs([(ta, tb)])
torch.cuda.synchronize()
