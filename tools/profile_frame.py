"""Per-launch device times (CUDA events between launches) of one synchronous Body() call.
usage: profile_frame.py [H W scale...]   default: 480 640 0.5 (the reference's default configuration)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_openpose_b200 import Body                         # noqa: E402
from pytorch_openpose_b200.model import random_checkpoint      # noqa: E402

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (480, 640)
scales = [float(x) for x in sys.argv[3:]] or [0.5]
body = Body(random_checkpoint("body", 0), scale_search=scales)
img = np.random.default_rng(0).integers(0, 256, (H, W, 3), dtype=np.uint8)
for _ in range(4):
    body(img)
s = body._session
s.set_profiling(True)
acc = {}
n = 10
for _ in range(n):
    body(img)
    for name, ms, gf in s.profile():
        a = acc.setdefault(name, [0.0, 0.0])
        a[0] += ms / n
        a[1] = gf
tot = sum(v[0] for v in acc.values())
print("step,ms,gflop,tflops")
for k, (ms, gf) in acc.items():
    print("%s,%.4f,%.3f,%.1f" % (k, ms, gf, gf / ms if ms > 0 else 0))
print("total,%.4f" % tot)
