"""Per-stage device times of one Hand() batch (32 crops of 368x368, 4 scales)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_openpose_b200 import Hand                         # noqa: E402
from pytorch_openpose_b200.model import random_checkpoint      # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
hand = Hand(random_checkpoint("hand", 0))
crops = np.random.default_rng(0).integers(0, 256, (B, 368, 368, 3), dtype=np.uint8)
for _ in range(3):
    hand(crops)
s = hand._session
s.set_profiling(True)
acc = {}
for _ in range(3):
    hand(crops)
    for name, ms, gf in s.profile():
        key = name.split(":")[0]
        acc[key] = acc.get(key, 0.0) + ms / 3
print({k: round(v, 3) for k, v in acc.items()}, "sum", round(sum(acc.values()), 3), "ms per batch of", B)
