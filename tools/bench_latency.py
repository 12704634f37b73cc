"""Drop-in usage latency: what a caller of the reference sees when it swaps the import and keeps calling
`body_estimation(oriImg)` / `hand_estimation(crop)` synchronously, one frame at a time (srcmx/MotionEstimation.py:139).
Host numpy arrays in, results out, wall clock per call."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_openpose_b200.model import random_checkpoint      # noqa: E402  (random-init weights; no checkpoints offline)
from pytorch_openpose_b200 import Body, Hand       # noqa: E402

rng = np.random.default_rng(0)
res = {}


def lat(fn, inputs, reps=60):
    for x in inputs[:4]:
        fn(x)
    torch.cuda.synchronize()
    t = []
    for i in range(reps):
        t0 = time.perf_counter()
        fn(inputs[i % len(inputs)])
        t.append((time.perf_counter() - t0) * 1e3)
    return {"ms_median": float(np.median(t)), "ms_p90": float(np.percentile(t, 90)), "calls_per_s": 1e3 / float(np.median(t))}


sd = random_checkpoint("body", 0)
res["body_c1_640x480_scale0.5"] = lat(Body(sd), [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(8)])
res["body_c2_720p_4scale"] = lat(Body(sd, scale_search=[0.5, 1.0, 1.5, 2.0]),
                                 [rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8) for _ in range(8)])
hand = Hand(random_checkpoint("hand", 0))
res["hand_184_4scale"] = lat(hand, [rng.integers(0, 256, (184, 184, 3), dtype=np.uint8) for _ in range(8)])
res["hand_changing_size_4scale"] = lat(hand, [rng.integers(0, 256, (w, w, 3), dtype=np.uint8) for w in range(150, 214, 8)])
print(json.dumps({"metric": "synchronous_call_latency", "results": res}))
