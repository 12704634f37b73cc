"""Per-launch device times of one Batch_body call (720p float frames, batch 16)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_openpose_b200 import Batch_body                   # noqa: E402
from pytorch_openpose_b200.model import random_checkpoint      # noqa: E402
import torch                                                    # noqa: E402

B = 16
est = Batch_body(random_checkpoint("body", 0))
fr = torch.from_numpy(np.random.default_rng(0).random((B, 3, 720, 1280), dtype=np.float32)).cuda()
for _ in range(4):
    est.submit(fr)
    est.collect()
s = est._session
s.set_profiling(True)
acc = {}
for _ in range(5):
    est.submit(fr)
    est.collect()
    for name, ms, gf in s.profile():
        key = name.split(":")[0]
        acc[key] = acc.get(key, 0.0) + ms / 5 / B
print({k: round(v, 4) for k, v in acc.items()}, "sum", round(sum(acc.values()), 4), "ms per frame")
