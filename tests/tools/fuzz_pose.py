"""Randomised parity run of the device-resident per-frame caller: motion.PoseEstimator (person selection, handDetect,
ragged hand batch on the device) vs motion.pose_mat_every_frame (host pipeline over Body / Hand) on random frames.
usage: fuzz_pose.py [cases] [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import openpose_oracle as O                    # noqa: E402
from pytorch_openpose_b200 import Body, Hand, motion        # noqa: E402
import cv2                                                   # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
body_sd = O.make_weights("body", 2, "kaiming")
hand = Hand(O.make_weights("hand", 5, "kaiming"), scale_search=[0.5, 1.0])
bad = with_hands = 0
for i in range(cases):
    H, W = int(rng.integers(100, 300)), int(rng.integers(150, 480))
    scales = [float(rng.choice([0.75, 1.0, 1.25]))]
    n = int(rng.integers(1, 4))
    frames = np.stack([cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), (0, 0), float(rng.uniform(1, 3)))
                       for _ in range(n)])
    body = Body(body_sd, scale_search=scales)
    est = motion.PoseEstimator(body, hand)
    pose = est(frames)
    for f in range(n):
        try:
            ref, cand, sub = motion.pose_mat_every_frame(frames[f], body, hand, "bodyhand")
        except ZeroDivisionError:                         # the reference raises on an empty hand box; the device gives zero rows
            print("case %d frame %d: empty hand box (reference raises)" % (i, f))
            continue
        hands = int((ref[18:, 2] > 0).any())
        with_hands += hands
        if not np.array_equal(pose[f], ref):
            bad += 1
            print("MISMATCH case %d frame %d (%dx%d scales %s): max diff %g" % (i, f, H, W, scales, np.abs(pose[f] - ref).max()))
    print("case %d ok: %dx%d scales %s frames %d persons(last) %d" % (i, H, W, scales, n, len(sub)), flush=True)
    del est, body
print("pose fuzz done: %d cases, %d mismatches, %d frames with hand key points" % (cases, bad, with_hands))
sys.exit(1 if bad else 0)
