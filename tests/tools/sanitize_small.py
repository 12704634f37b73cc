"""Small run of every post-processing path for compute-sanitizer (memcheck / racecheck): Body single + batch, maps on
request, Batch_body, stage-level entry points on a synthetic scene."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import openpose_oracle as O                     # noqa: E402
from pytorch_openpose_b200 import Body, Batch_body          # noqa: E402
from pytorch_openpose_b200.model import random_checkpoint   # noqa: E402

sd = random_checkpoint("body", 0)
rng = np.random.default_rng(0)
for shape, scales in (((120, 160, 3), [0.5, 1.0]), ((97, 131, 3), [0.5, 1.0]), ((200, 300, 3), [0.5])):
    body = Body(sd, scale_search=scales)
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    for _ in range(4):                      # eager, eager, capture, replay
        cand, subset = body(img)
    heat, paf = body.last_maps(img.shape)
    rc, rs = O.body_postprocess(heat.astype(np.float64), paf.astype(np.float64), shape[0])
    assert np.array_equal(cand, rc) and np.array_equal(subset, rs), shape
    frames = rng.integers(0, 256, (3,) + shape, dtype=np.uint8)
    out = body.batch(frames)
    assert len(out) == 3
est = Batch_body(sd)
fr = rng.random((2, 3, 120, 160), dtype=np.float32)
est(fr)
from tests import gpu_util as G                              # noqa: E402
heat, paf, _ = O.synthetic_scene(240, 320, (2, 1), seed=0)
cand, pb, cand_dev = G.find_peaks(heat.transpose(2, 0, 1))
subset, conns, cc = G.group_limbs(paf.transpose(2, 0, 1), cand_dev, pb)
rc, rs = O.body_postprocess(heat, paf, 240)
assert np.array_equal(cand, rc) and np.array_equal(subset, rs)
print("sanitize run ok")
