"""End-to-end extraction job on a synthetic 720p MJPG video: cv2 decode thread -> pinned
batch ring -> Body (4 scales) on 3 sessions -> pose track -> joblib file.  Reports frames/s of the whole job (wall
clock, decode included) and of the decode alone, i.e. where the host becomes the limit."""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2                                              # noqa: E402
from pytorch_openpose_b200.model import random_checkpoint      # noqa: E402  (random-init weights; no checkpoints offline)
from pytorch_openpose_b200 import Body, Batch_body, extract   # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 384
d = tempfile.mkdtemp()
path = os.path.join(d, "v.avi")
wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 25, (1280, 720))
rng = np.random.default_rng(0)
base = [cv2.GaussianBlur(rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8), (0, 0), 5) for _ in range(8)]
for i in range(N):
    wr.write(np.roll(base[i % 8], 7 * i, axis=1))
wr.release()

W = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dec = {}
for k in (1, W):
    sum(len(f) for f, _ in extract.FrameBatches(path, None, batch=8, depth=5, pinned=True, workers=k))     # warm-up
    t0 = time.perf_counter()
    n = sum(len(f) for f, _ in extract.FrameBatches(path, None, batch=8, depth=5, pinned=True, workers=k))
    dec[k] = n / (time.perf_counter() - t0)

body = Body(random_checkpoint("body", 0), scale_search=[0.5, 1.0, 1.5, 2.0])
extract.extract_motion_from_video(path, os.path.join(d, "w.pkl"), None, body, mode="body", batch=8, sessions=3,
                                  log=lambda m: None)                      # warm-up: plans
job = {}
for k in (1, W):
    extract.extract_motion_from_video(path, os.path.join(d, "w.pkl"), None, body, mode="body", batch=8, sessions=3,
                                      log=lambda m: None, decode_workers=k)        # pinned rings of this shape cached
    t0 = time.perf_counter()
    st = {}
    mat = extract.extract_motion_from_video(path, os.path.join(d, "o.pkl"), None, body, mode="body", batch=8, sessions=3,
                                            log=lambda m: None, decode_workers=k, stats=st)
    job[k] = len(mat) / (time.perf_counter() - t0)
    job["seconds_%d" % k] = {a: round(b, 3) for a, b in st.items()}

# ---- body + two hands per frame (the reference's 'bodyhand' mode).  Random-init weights find nobody, so every frame
# gets one synthetic person whose arm joints (and therefore both hand boxes, ~150-200 px, a new size every frame)
# move with the frame index; the hand network and its post-processing really run on those crops.
from pytorch_openpose_b200 import Hand          # noqa: E402


class BodyP(Body):
    n = 0

    def _person(self):
        k = BodyP.n
        BodyP.n += 1
        joints = {2: (500, 200), 3: (430 + k % 17, 330), 4: (380 + 2 * (k % 13), 460 + k % 19),
                  5: (780, 200), 6: (850 - k % 16, 330), 7: (900 - 2 * (k % 11), 460 + k % 23)}
        cand = np.zeros((8, 4))
        row = -np.ones(20)
        for i, (j, (x, y)) in enumerate(sorted(joints.items())):
            cand[i] = (x, y, 0.9, i)
            row[j] = i
        row[18], row[19] = 6.0, 6
        return cand, row[None].copy()

    def collect_batch(self, session=None):
        return [self._person() for _ in Body.collect_batch(self, session)]


bodyp = BodyP(body.net_weights if hasattr(body, "net_weights") else random_checkpoint("body", 0), scale_search=[0.5, 1.0, 1.5, 2.0])
hand = Hand(random_checkpoint("hand", 0))
bh = {}
for k in (W,):
    extract.extract_motion_from_video(path, os.path.join(d, "wh.pkl"), None, bodyp, hand, mode="bodyhand", batch=8,
                                      sessions=3, log=lambda m: None, decode_workers=k)
    t0 = time.perf_counter()
    mat = extract.extract_motion_from_video(path, os.path.join(d, "oh.pkl"), None, bodyp, hand, mode="bodyhand", batch=8,
                                            sessions=3, log=lambda m: None, decode_workers=k)
    bh[k] = len(mat) / (time.perf_counter() - t0)
    bh["frames_with_both_hands"] = int(((mat[:, 18:39, 2] > 0).any(1) & (mat[:, 39:, 2] > 0).any(1)).sum())
del bodyp, hand
del body
bb = Batch_body(random_checkpoint("body", 0))
extract.batch_body_extraction(path, os.path.join(d, "wb.pkl"), 16, None, bb, log=lambda m: None)
bjob = {}
for k in (1, W):
    extract.batch_body_extraction(path, os.path.join(d, "wb.pkl"), 16, None, bb, log=lambda m: None, decode_workers=k)
    t0 = time.perf_counter()
    mat = extract.batch_body_extraction(path, os.path.join(d, "ob.pkl"), 16, None, bb, log=lambda m: None, decode_workers=k)
    bjob[k] = len(mat) / (time.perf_counter() - t0)
print(json.dumps({"metric": "extraction_job_frames_per_sec_720p", "frames": N, "decode_only_fps_by_workers": dec,
                  "body_4scale_job_fps_by_workers": job, "bodyhand_4scale_job_fps_by_workers": bh, "batch_body_job_fps_by_workers": bjob, "host_cores": os.cpu_count(),
                  "note": "cv2.VideoCapture decode threads (MJPG 720p); wall clock incl. decode, H2D, D2H, file write"}))
