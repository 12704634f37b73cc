"""Randomised parity run: Body() on frames of random sizes / scale lists / weights; the discrete results must equal the
oracle's post-processing of the device-produced maps (north_star criterion 2).  usage: fuzz_body.py [cases] [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import openpose_oracle as O            # noqa: E402
from pytorch_openpose_b200 import Body, Batch_body  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
import cv2                                           # noqa: E402
weights = [O.make_weights("body", 2, "kaiming"), O.make_weights("body", 0)]
bad = 0
for i in range(cases):
    H, W = int(rng.integers(12, 420)), int(rng.integers(12, 520))
    k = int(rng.integers(1, 4))
    scales = sorted(float(s) for s in rng.choice([0.25, 0.5, 0.75, 1.0, 1.5, 2.0], k, replace=False))
    if max(scales) * 368 * max(W / H, 1.0) > 2600:            # keep the net input small
        scales = [s for s in scales if s <= 1.0] or [0.5]
    sd = weights[i % 2]
    img = cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), (0, 0), float(rng.uniform(0.5, 4)))
    n = int(rng.integers(1, 4))
    body = Body(sd, scale_search=scales)
    try:
        if n == 1:
            out = [body(img)]
            heat, paf = body.last_maps(img.shape)
            heat, paf = heat[None], paf[None]
        else:
            frames = np.stack([np.roll(img, 3 * f, axis=1) for f in range(n)])
            out = body.batch(frames)
            heat, paf = body.last_maps(frames.shape)
        for f in range(n):
            rc, rs = O.body_postprocess(heat[f].astype(np.float64), paf[f].astype(np.float64), H)
            ok = out[f][0].shape == rc.shape and np.array_equal(out[f][0], rc) and np.array_equal(out[f][1], rs)
            if not ok:
                bad += 1
                print("MISMATCH case %d frame %d: H=%d W=%d scales=%s: %d vs %d candidates" % (i, f, H, W, scales, len(out[f][0]), len(rc)))
    except IndexError:
        try:
            O.body_postprocess(heat[0].astype(np.float64), paf[0].astype(np.float64), H)
            print("case %d: device raised IndexError, oracle did not" % i)
            bad += 1
        except IndexError:
            pass
    print("case %d ok: %dx%d scales %s frames %d candidates %d persons %d" % (i, H, W, scales, n, len(out[0][0]), len(out[0][1])), flush=True)
    del body
print("fuzz done: %d cases, %d mismatches" % (cases, bad))
sys.exit(1 if bad else 0)
