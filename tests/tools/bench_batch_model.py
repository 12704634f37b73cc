"""Row N2 measurement: the reference author's throughput path (srcmx/Batch_model.py, batch 16-32, single scale).

Batch_body on 720p float frames (scale 0.5 -> 327x184 net input) and Batch_hand on 368x368 float crops, one B200,
through the public classes with HOST float batches (H2D of the float frames and D2H of the results inside the timed
region, wall clock), double-buffered over two sessions; next to it the CPU oracle port on a bounded sample."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import openpose_oracle as O                        # noqa: E402
from pytorch_openpose_b200 import Batch_body, Batch_hand       # noqa: E402

rng = np.random.default_rng(0)
res = {}


def run(est, batches, reps, submit=None):
    submit = submit or est.submit
    ss = [est._session, est.net.session()]
    for s, b in zip(ss, batches):                  # warm-up: plans
        submit(b, s)
    for s in ss:
        est.collect(s)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pending = []
    for i in range(reps):
        s = ss[i % 2]
        if len(pending) == 2:
            est.collect(pending.pop(0))
        submit(batches[i % len(batches)], s)
        pending.append(s)
    for s in pending:
        est.collect(s)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


B = 16
frames = [torch.from_numpy(rng.random((B, 3, 720, 1280), dtype=np.float32)).pin_memory() for _ in range(2)]
body = Batch_body(O.make_weights("body", 0))
dt = run(body, frames, 12)
res["batch_body_720p_b16"] = {"frames_per_s": B / dt, "ms_per_batch": dt * 1e3, "gflop_per_frame": 121.16}
dev = [f.cuda() for f in frames]
dt = run(body, dev, 12)
res["batch_body_720p_b16"]["frames_per_s_device_resident_input"] = B / dt
del dev
# decoded uint8 frames as they come out of cv2 (what the extraction job feeds): 4x fewer bytes over PCIe than float
# frames (16 x 11 MB instead of 16 x 11 MB x 4 = 177 MB per batch, which is what bounds the float-input figure above)
u8 = [torch.from_numpy(rng.integers(0, 256, (B, 720, 1280, 3), dtype=np.uint8)).pin_memory() for _ in range(2)]
dt = run(body, [u.numpy() for u in u8], 12, submit=lambda b, s: body.submit_frames(b, s, where=2))
res["batch_body_720p_b16"]["frames_per_s_uint8_frames_from_pinned_host"] = B / dt
del u8
t0 = time.perf_counter()
O.batch_body_call(frames[0][:2].numpy(), O.make_weights("body", 0))
res["batch_body_720p_b16"]["cpu_port_frames_per_s"] = 2 / (time.perf_counter() - t0)
del body

Bh = 64
crops = [torch.from_numpy(rng.random((Bh, 3, 368, 368), dtype=np.float32)).pin_memory() for _ in range(2)]
hand = Batch_hand(O.make_weights("hand", 0))
dt = run(hand, crops, 12)
res["batch_hand_368_b64"] = {"crops_per_s": Bh / dt, "ms_per_batch": dt * 1e3, "tflops": 206.38 * Bh / dt * 1e-3}
t0 = time.perf_counter()
O.batch_hand_call(crops[0][:2].numpy(), O.make_weights("hand", 0))
res["batch_hand_368_b64"]["cpu_port_crops_per_s"] = 2 / (time.perf_counter() - t0)
print(json.dumps({"metric": "batched_estimators_throughput", "results": res,
                  "timing": "host wall clock, pinned float host batches in, results out, 2 sessions", "cpu_cores": os.cpu_count()}))
