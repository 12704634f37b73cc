"""Small fixed workload for ncu: batches of B 720p frames through the device-resident body + two-hands pipeline
(motion.PoseEstimator, fixed 184x184 hand boxes as in BASELINE config 4)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_openpose_b200 import Body, Hand, motion             # noqa: E402
from pytorch_openpose_b200.model import random_checkpoint        # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
body = Body(random_checkpoint("body", 0), scale_search=[0.5, 1.0, 1.5, 2.0])
hand = Hand(random_checkpoint("hand", 0))
est = motion.PoseEstimator(body, hand)
frames = np.random.default_rng(0).integers(0, 256, (B, 720, 1280, 3), dtype=np.uint8)
boxes = np.tile(np.array([[400, 300, 184], [700, 300, 184]], dtype=np.int32), (B, 1, 1))
for _ in range(reps):
    est.submit_batch(frames, fixed_boxes=boxes)
    pose = est.collect()
print("ok", pose.shape, int((pose[:, 18:, 2] > 0).sum()))
