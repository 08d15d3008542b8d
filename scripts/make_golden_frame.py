"""Generates tests/golden/ref_frame_*.npz ON A GPU BOX (run under gpurun): raw depth images pushed through the REFERENCE'S
OWN front-end kernels (oracle/_ref/libbt_ref.so = /root/reference/src/cuda/CUDAImageUtil.cu compiled verbatim with the
reference's flags; host sequence of Frame.cpp:62-81,152-233 in oracle/ref_harness.cu).  The CPU suite pins
oracle/frame_oracle.py against these vectors without a GPU.

    gpurun -- 'python scripts/make_golden_frame.py gpurun_out/golden'   &&  cp gpurun_out/golden/ref_frame_*.npz tests/golden/
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle
from oracle import frame_oracle
from bundletrack_b200 import synth

out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden"
os.makedirs(out_dir, exist_ok=True)
cases = [  # (seed, H, W, depth-processing overrides)
    (201, 240, 320, {}),
    (202, 150, 200, {"erode_radius": 2, "erode_diff": 0.004, "erode_ratio": 0.5, "bf_radius": 1, "sigma_D": 1.5, "sigma_R": 0.05}),
]
for idx, (seed, H, W, dp) in enumerate(cases):
    raw, K = synth.make_raw_depth(seed, H, W)
    d, xyz, n, t_ms = oracle.ref_frame_preprocess(raw, K, dp)
    od, ox, on = frame_oracle.preprocess(raw, K, dp)
    bad_d = np.mean(np.abs(od - d) > 2e-6)
    bad_n = np.mean(np.abs(on - n).max(-1) > 1e-3)
    print(f"case {idx}: {W}x{H} valid {int((d > 0).sum())}  oracle vs reference kernels: depth max|diff| {np.abs(od - d).max():.2e} (frac > 2e-6: {bad_d:.2e}), "
          f"normal frac > 1e-3: {bad_n:.2e}, ref time {t_ms:.2f} ms")
    np.savez_compressed(os.path.join(out_dir, f"ref_frame_{idx}.npz"), raw=raw, K=np.asarray(K, np.float32), depth=d, normal=n[..., :3].copy(), xyz=xyz[..., :3].copy(),
                        **{k: np.float32(v) for k, v in dict(frame_oracle.DEFAULTS, **dp).items()})
