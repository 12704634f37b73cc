"""Randomised parity run of Hand() / Batch_body() / Batch_hand(): random crop and frame sizes, batches; discrete results
must equal the oracle's post-processing of the device-produced maps.  usage: fuzz_hand.py [cases] [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import openpose_oracle as O                                  # noqa: E402
from pytorch_openpose_b200 import Hand, Batch_body, Batch_hand           # noqa: E402
import cv2                                                                 # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
hand = Hand(O.make_weights("hand", 5, "kaiming"), scale_search=[0.5, 1.0])
bb = Batch_body(O.make_weights("body", 2, "kaiming"))
bh = Batch_hand(O.make_weights("hand", 5, "kaiming"))
bad = 0
for i in range(cases):
    h, w = int(rng.integers(8, 260)), int(rng.integers(8, 260))
    n = int(rng.integers(1, 4))
    crops = np.stack([cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), float(rng.uniform(0.5, 3)))
                      for _ in range(n)])
    peaks = hand(crops)
    maps = hand.last_maps(crops.shape)
    for f in range(n):
        if not np.array_equal(peaks[f], O.hand_postprocess(maps[f].astype(np.float64))):
            bad += 1
            print("MISMATCH Hand case %d crop %d (%dx%d)" % (i, f, h, w))
    # batched estimators: float frames in [0, 1]
    H, W = int(rng.integers(40, 300)), int(rng.integers(40, 400))
    fr = rng.random((n, 3, H, W), dtype=np.float32)
    fr = np.stack([cv2.GaussianBlur(f.transpose(1, 2, 0), (0, 0), 2.0).transpose(2, 0, 1) for f in fr]).astype(np.float32)
    res = bb(fr)
    blurred, paf = bb.last_maps()
    for f in range(n):
        rc, rs = O.batch_body_postprocess(blurred[f], paf[f])
        if not (np.array_equal(np.asarray(res[f][0]).reshape(-1, 4), rc.reshape(-1, 4)) and np.array_equal(res[f][1], rs)):
            bad += 1
            print("MISMATCH Batch_body case %d frame %d (%dx%d)" % (i, f, H, W))
    s = int(rng.integers(2, 30)) * 8
    cr = rng.random((n, 3, s, s), dtype=np.float32)
    pk = bh(cr)
    bl = bh.last_maps()
    for f in range(n):
        if not np.array_equal(pk[f], O.batch_hand_postprocess(bl[f])):
            bad += 1
            print("MISMATCH Batch_hand case %d crop %d (%d)" % (i, f, s))
    print("case %d ok: hand %dx%d x%d, batch_body %dx%d, batch_hand %d" % (i, h, w, n, H, W, s), flush=True)
print("hand fuzz done: %d cases, %d mismatches" % (cases, bad))
sys.exit(1 if bad else 0)
