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
from oracle import openpose_oracle as O                # noqa: E402
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

body = Body(O.make_weights("body", 0), scale_search=[0.5, 1.0, 1.5, 2.0])
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
del body
bb = Batch_body(O.make_weights("body", 0))
extract.batch_body_extraction(path, os.path.join(d, "wb.pkl"), 16, None, bb, log=lambda m: None)
bjob = {}
for k in (1, W):
    extract.batch_body_extraction(path, os.path.join(d, "wb.pkl"), 16, None, bb, log=lambda m: None, decode_workers=k)
    t0 = time.perf_counter()
    mat = extract.batch_body_extraction(path, os.path.join(d, "ob.pkl"), 16, None, bb, log=lambda m: None, decode_workers=k)
    bjob[k] = len(mat) / (time.perf_counter() - t0)
print(json.dumps({"metric": "extraction_job_frames_per_sec_720p", "frames": N, "decode_only_fps_by_workers": dec,
                  "body_4scale_job_fps_by_workers": job, "batch_body_job_fps_by_workers": bjob, "host_cores": os.cpu_count(),
                  "note": "cv2.VideoCapture decode threads (MJPG 720p); wall clock incl. decode, H2D, D2H, file write"}))
