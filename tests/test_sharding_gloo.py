"""world_size-2 gloo test of the multi-GPU plumbing (windows sharded over ranks, poses gathered): runs on CPU with the
oracle standing in for the per-rank solver, and must reproduce the single-process result exactly."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _windows():
    from bundletrack_b200 import synth
    return [synth.make_window(200 + k, n_frames=2 + (k % 3), n_corr=60 + 10 * k, H=120, W=160, K=tuple(v * 0.25 for v in synth.NOCS_K)) for k in range(5)]


def _solve(w):
    import oracle
    return oracle.solve_window(w.depth, w.normal, w.K, w.corr, w.poses_init)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from bundletrack_b200.sharding import shard_indices, gather_poses
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ws = _windows()
    mine = shard_indices(len(ws), rank, world)
    local = [_solve(ws[i]) for i in mine]
    full = gather_poses(local, len(ws), [w.n_frames for w in ws], rank, world)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.concatenate([p.reshape(-1) for p in full]))
    dist.barrier(); dist.destroy_process_group()


def test_shard_and_gather_world2(tmp_path):
    from bundletrack_b200.sharding import shard_indices
    assert shard_indices(5, 0, 2) == [0, 2, 4] and shard_indices(5, 1, 2) == [1, 3]
    assert sorted(shard_indices(7, 0, 3) + shard_indices(7, 1, 3) + shard_indices(7, 2, 3)) == list(range(7))
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    want = np.concatenate([_solve(w).reshape(-1) for w in _windows()])
    for r in range(2):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert np.array_equal(got, want)
