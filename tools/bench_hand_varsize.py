"""Hand() on crops whose size changes every call (what srcmx/MotionEstimation.py does: the box width follows the arm):
first-visit vs revisit latency per call and device memory held by the per-size plans."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_openpose_b200.model import random_checkpoint      # noqa: E402  (random-init weights; no checkpoints offline)
from pytorch_openpose_b200 import Hand             # noqa: E402

rng = np.random.default_rng(0)
sizes = list(range(150, 270, 2))
crops = [rng.integers(0, 256, (w, w, 3), dtype=np.uint8) for w in sizes]
hand = Hand(random_checkpoint("hand", 0))
hand(crops[0])
torch.cuda.synchronize()
free0 = torch.cuda.mem_get_info()[0]
out = {}
for tag in ("first_visit", "revisit"):
    t = []
    for c in crops[1:]:
        t0 = time.perf_counter()
        hand(c)
        t.append((time.perf_counter() - t0) * 1e3)
    out[tag + "_ms_median"] = float(np.median(t))
    out[tag + "_ms_max"] = float(np.max(t))
out["device_MB_held_by_%d_sizes" % (len(sizes) - 1)] = (free0 - torch.cuda.mem_get_info()[0]) / 2**20
print(json.dumps(out))
