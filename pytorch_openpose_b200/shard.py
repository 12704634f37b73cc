"""Frame sharding across GPUs: frames are independent, so every rank owns a full replica of the weights and a
disjoint subset of the frames; no collective runs on the data path.  Mirrors the reference's only
scale-out mechanism -- one process per GPU on disjoint work (srcmx/Batch_motion_Estimation.py:143-156,194-200)."""
import numpy as np


def frames_for_rank(n_frames, rank, world, mode="interleave"):
    """Indices of the frames rank `rank` of `world` processes.  'interleave': i % world == rank (balanced for
    streams); 'chunk': contiguous blocks (keeps video decode sequential per rank)."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world %d" % (rank, world))
    if mode == "interleave":
        return np.arange(rank, n_frames, world)
    if mode == "chunk":
        bounds = np.linspace(0, n_frames, world + 1).astype(int)
        return np.arange(bounds[rank], bounds[rank + 1])
    raise ValueError(mode)


def gather_pose_mats(local_indices, local_mats, n_frames, group=None):
    """Assemble the per-frame fixed-size records (e.g. PoseMat (60,3), srcmx/MotionEstimation.py:141) of all ranks
    into one (n_frames, ...) array on every rank.  One end-of-stream all_gather, off the per-frame path."""
    import torch
    import torch.distributed as dist
    local_mats = np.asarray(local_mats, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()):
        out = np.zeros((n_frames,) + local_mats.shape[1:])
        out[np.asarray(local_indices, dtype=int)] = local_mats
        return out
    world = dist.get_world_size(group)
    payload = (np.asarray(local_indices, dtype=np.int64), local_mats)
    gathered = [None] * world
    dist.all_gather_object(gathered, payload, group=group)
    shape = next(m.shape[1:] for _, m in gathered if len(m))
    out = np.zeros((n_frames,) + shape)
    for idx, mats in gathered:
        if len(idx):
            out[idx] = mats
    return out
