"""BASELINE config 3: Hand() on 256 synthetic 368x368 crops, one B200 (4-scale like src/hand.py, and single-scale)."""
import json
import os
import sys
import time

import numpy as np
import torch
import faulthandler
faulthandler.dump_traceback_later(int(os.environ.get("OPB_DBG_TIMEOUT", "50")), exit=True)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_openpose_b200.model import random_checkpoint      # noqa: E402  (random-init weights; no checkpoints offline)
from pytorch_openpose_b200 import Hand             # noqa: E402

crops = np.random.default_rng(0).integers(0, 256, (256, 368, 368, 3), dtype=np.uint8)
res = {}
for tag, scales, gflop in (("4scale", [0.5, 1.0, 1.5, 2.0], 1547.82), ("1scale", [1.0], 206.38)):
    hand = Hand(random_checkpoint("hand", 0), scale_search=scales)
    hand(crops)                                     # warm-up: builds the plans
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        peaks = hand(crops)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    res[tag] = {"crops_per_s": 256 / dt, "ms_per_batch_of_256": dt * 1e3, "tflops": gflop * 256 / dt * 1e-3}
    del hand
print(json.dumps({"metric": "hand_crops_per_sec_368x368_batch256", "results": res, "timing": "host wall clock around Hand()(crops) incl. H2D/D2H"}))
