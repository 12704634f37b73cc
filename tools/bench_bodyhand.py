"""BASELINE config 4: 720p video stream, body (4 scales) + two hand crops (4 scales) per frame, one GPU per rank.

Random-init weights find no person, so the two hand boxes are fixed 184x184 crops; the hand net
always sees 184/368/552/736-pixel inputs regardless of the crop size (src/hand.py:32).  Frames go through
Body.submit_batch, the 2*B crops of a batch through one Hand.submit; several sessions keep the GPU busy.
Prints one JSON line (rank 0)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_openpose_b200.model import random_checkpoint      # noqa: E402  (random-init weights; no checkpoints offline)
from pytorch_openpose_b200 import Body, Hand       # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--frames-per-step", type=int, default=32)
ap.add_argument("--streams", type=int, default=2)
args = ap.parse_args()
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H, W, B, F = 720, 1280, args.batch, args.frames_per_step
body = Body(random_checkpoint("body", 0), scale_search=[0.5, 1.0, 1.5, 2.0], device=local)
hand = Hand(random_checkpoint("hand", 0), device=local)
bs = [body.net.session() for _ in range(args.streams)]
hs = [hand.net.session() for _ in range(args.streams)]
rng = np.random.default_rng(rank)
pool = torch.from_numpy(np.repeat(np.repeat(rng.integers(0, 256, (64, H // 8, W // 8, 3), dtype=np.uint8), 8, 1), 8, 2)).pin_memory()
frames = pool.numpy()
boxes = [(700, 300, 184), (400, 300, 184)]          # fixed right / left hand boxes (x, y, w)


def step(i):
    inflight = [False] * args.streams
    for b in range(F // B):
        si = b % args.streams
        if inflight[si]:
            body.collect_batch(bs[si])
            hand.collect(hs[si])
        idx = (i * F + b * B) % 64
        fr = frames[idx:idx + B]
        body.submit_batch(fr, bs[si], where=2)
        crops = np.stack([fr[f, y:y + w, x:x + w] if k == 0 else fr[f, y:y + w, x:x + w][:, ::-1]
                          for f in range(B) for k, (x, y, w) in enumerate(boxes)])      # left hand mirrored
        hand.submit(crops, hs[si])
        inflight[si] = True
    for si in range(args.streams):
        if inflight[si]:
            body.collect_batch(bs[si])
            hand.collect(hs[si])


for i in range(args.warmup):
    step(i)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
bs[0].mark(0)
for i in range(args.steps):
    step(args.warmup + i)
for s in bs + hs:
    s.mark(1)
torch.cuda.synchronize()
ms = max(bs[0].elapsed_ms(0, s, 1) for s in bs + hs)
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
if rank == 0:
    n = F * args.steps * world
    print(json.dumps({"metric": "body_plus_two_hands_frames_per_sec_720p_4scale", "value": n / (ms * 1e-3), "unit": "frames/s",
                      "n_gpus": world, "ms_per_frame_per_gpu": ms / (F * args.steps), "gflop_per_frame": 6730.4,
                      "tflops": 6730.4 * n / ms,
                      "config": {"frames_per_batch": B, "streams": args.streams, "hand_crop": 184, "data": "synthetic, pinned host frames"}}))
if world > 1:
    dist.destroy_process_group()
