"""Small fixed workload for ncu: N frames of the BASELINE C2 body path (720p, 4 scales) on one stream."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_openpose_b200.model import random_checkpoint      # noqa: E402  (random-init weights; no checkpoints offline)
from pytorch_openpose_b200 import Body             # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
body = Body(random_checkpoint("body", 0), scale_search=[0.5, 1.0, 1.5, 2.0])
rng = np.random.default_rng(0)
for i in range(n):
    img = rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    cand, subset = body(img)
print("ok", len(cand), len(subset))
