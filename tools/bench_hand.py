"""Hand path timings (BASELINE config 3 style): batched 368x368 crops, 4 scales, per-stage CUDA-event profile."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_openpose_b200.model import random_checkpoint      # noqa: E402  (random-init weights; no checkpoints offline)
from pytorch_openpose_b200 import Hand             # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
scales = [0.5, 1.0, 1.5, 2.0] if len(sys.argv) < 3 else [float(s) for s in sys.argv[2].split(",")]
hand = Hand(random_checkpoint("hand", 0), scale_search=scales)
crops = np.random.default_rng(0).integers(0, 256, (batch, 368, 368, 3), dtype=np.uint8)
s = hand._session
for _ in range(2):
    hand(crops)
s.set_profiling(True)
t0 = time.perf_counter()
hand(crops)
dt = time.perf_counter() - t0
prof = s.profile()
agg = {}
for name, ms, gf in prof:
    k = name.split(":")[0]
    a = agg.setdefault(k, [0.0, 0.0])
    a[0] += ms
    a[1] += gf
tot = sum(a[0] for a in agg.values())
print("batch %d scales %s: wall %.2f ms, device %.2f ms -> %.1f crops/s" % (batch, scales, dt * 1e3, tot, batch / (tot * 1e-3)))
for k, (ms, gf) in agg.items():
    print("  %-14s %8.3f ms  %8.1f GFLOP  %7.1f TFLOP/s" % (k, ms, gf, gf / ms if ms > 0 and gf > 0 else 0.0))
