"""Where the MMA-issuing thread of k_knn_tc waits (developer aid): cycles per CTA by cause, for the 45-pair batch."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bundletrack_b200 import synth, _lib
from bundletrack_b200.matcher import KnnMatcher
dev = torch.device("cuda:0")
m = KnnMatcher(max_pairs=48, max_feats=2048)
m.pool_reserve(10)
for f in range(10):
    m.pool_store(f, torch.from_numpy(synth.make_descriptors(100 + f, 2000, 8)[0]).to(dev))
idx = [(j, i) for i in range(10) for j in range(i + 1, 10)]
for _ in range(3):
    m.knn_match_slots(idx, [(2000, 2000)] * 45, device=dev)
torch.cuda.synchronize()
out = np.zeros((160, 6), np.int64)
lib = _lib.load()
lib.bt_knn_debug_prof.restype = ctypes.c_int
assert lib.bt_knn_debug_prof(out.ctypes.data_as(ctypes.c_void_p)) == 0
act = out[out[:, 4] > 0]
print(f"CTAs issuing: {len(act)}; units/CTA {act[:,4].mean():.1f}; cycles total {act[:,3].mean():.0f}; waiting on tm_empty {act[:,0].mean():.0f}, b_full {act[:,1].mean():.0f}, a full {act[:,2].mean():.0f}")
print(f"  per unit: total {(act[:,3]/act[:,4]).mean():.0f}, tm_empty {(act[:,0]/act[:,4]).mean():.0f}, b_full {(act[:,1]/act[:,4]).mean():.0f}, a full {(act[:,2]/act[:,4]).mean():.0f}")
