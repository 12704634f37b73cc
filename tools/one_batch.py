"""Small fixed workload for ncu: one batch of B 720p frames through the 4-scale body path on one stream."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_openpose_b200 import model                # noqa: E402
from pytorch_openpose_b200 import Body             # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
body = Body(model.random_checkpoint("body", 0), scale_search=[0.5, 1.0, 1.5, 2.0])
frames = np.random.default_rng(0).integers(0, 256, (B, 720, 1280, 3), dtype=np.uint8)
out = body.batch(frames)
print("ok", [len(c) for c, s in out])
