"""world_size-2 gloo test of the frame-sharding host logic (the N>1 path has no data-path collective; the only
exchange is the end-of-stream gather of fixed-size per-frame records)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from pytorch_openpose_b200.shard import frames_for_rank, gather_pose_mats


def test_partition_is_disjoint_and_complete():
    for mode in ("interleave", "chunk"):
        for n, world in ((0, 2), (1, 2), (17, 2), (64, 8), (5, 8)):
            parts = [frames_for_rank(n, r, world, mode) for r in range(world)]
            allidx = np.sort(np.concatenate(parts)) if n else np.array([])
            assert np.array_equal(allidx, np.arange(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, n_frames, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = frames_for_rank(n_frames, rank, world)
    mats = np.stack([np.full((60, 3), float(i)) for i in idx]) if len(idx) else np.zeros((0, 60, 3))
    full = gather_pose_mats(idx, mats, n_frames)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), full)
    dist.destroy_process_group()


def test_gather_two_ranks(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 7, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(a, b) and a.shape == (7, 60, 3)
    assert np.array_equal(a[:, 0, 0], np.arange(7.0))
