"""The per-frame caller on the device (opb_pose_*): person selection, util.handDetect, hand crops
cut from the frame in device memory, one ragged Hand batch, PoseMat as the only result -- equal to the host pipeline
(pytorch_openpose_b200/motion.py::pose_mat_every_frame == srcmx/MotionEstimation.py:126-216) bit for bit."""
import ctypes

import numpy as np
import pytest

from oracle import openpose_oracle as O

pytestmark = pytest.mark.gpu


def _device_select(candidate, subset, H, W):
    from pytorch_openpose_b200 import _lib
    cand = np.ascontiguousarray(candidate, dtype=np.float64).reshape(-1, 4)
    sub = np.ascontiguousarray(subset, dtype=np.float64).reshape(-1, 20)
    pose = np.zeros((60, 3))
    boxes = np.zeros((2, 4), dtype=np.int32)
    _lib.check(_lib.lib().opb_pose_select(_lib.context(0), cand.ctypes.data, len(cand), sub.ctypes.data, len(sub), H, W,
                                          pose.ctypes.data, boxes.ctypes.data))
    return pose, boxes


def test_person_selection_and_hand_boxes_equal_the_host_functions():
    """Random people (missing joints, arms leaving the frame, ties in the shoulder position) -> body rows of PoseMat and
    the two boxes equal select_person + util.handDetect of the host pipeline (themselves checked against the live
    reference in tests/test_properties.py)."""
    from pytorch_openpose_b200 import motion, util
    rng = np.random.default_rng(5)
    H, W = 240, 320
    img = np.zeros((H, W, 3), np.uint8)
    for trial in range(300):
        n_people = int(rng.integers(0, 4))
        cand, rows = [], []
        for p in range(n_people):
            row = -np.ones(20)
            for part in range(18):
                if rng.random() < 0.8:
                    row[part] = len(cand)
                    x = float(rng.integers(0, W)) if rng.random() < 0.9 else float(W - 1)
                    cand.append([x, float(rng.integers(0, H)), float(rng.random()), float(len(cand))])
            row[18], row[19] = float(rng.random() * 10), float((row[:18] >= 0).sum())
            rows.append(row)
        candidate = np.array(cand).reshape(-1, 4)
        subset = np.array(rows).reshape(-1, 20)
        pose, boxes = _device_select(candidate, subset, H, W)
        ref = np.zeros((60, 3))
        chosen = motion.select_person(candidate, subset) if len(subset) else None
        want = {True: (0, 0, 0, 0), False: (0, 0, 0, 0)}
        if chosen is not None:
            for k in range(18):
                idx = int(subset[chosen][k])
                if idx != -1:
                    ref[k] = candidate[idx][:3]
            only = -np.ones_like(subset)
            only[chosen] = subset[chosen]
            for x, y, w, is_left in util.handDetect(candidate, only, img):
                want[bool(is_left)] = (x, y, w, int(w > 0))
        assert np.array_equal(pose, ref), trial
        for k, left in ((0, True), (1, False)):
            if want[left][3]:
                assert tuple(boxes[k]) == want[left], (trial, k)
            else:
                assert boxes[k][3] == 0, (trial, k)


def test_ragged_hand_batch_equals_hand_on_every_crop():
    """fixed boxes of different sizes (one touching the border, one empty, left ones mirrored): the hand rows of the
    device PoseMat equal Hand()(crop) + the caller's coordinate arithmetic, the body rows equal the Body results."""
    import cv2
    from pytorch_openpose_b200 import Body, Hand, extract, motion
    rng = np.random.default_rng(8)
    H, W, n = 240, 320, 3
    frames = np.stack([cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), (0, 0), 2) for _ in range(n)])
    body = Body(O.make_weights("body", 0), scale_search=[0.5])
    hand = Hand(O.make_weights("hand", 5, "kaiming"), scale_search=[0.5, 1.0])
    boxes = np.array([[[10, 20, 97], [200, 100, 120]],
                      [[0, 0, 64], [150, 30, 0]],             # right hand: empty box -> zero rows
                      [[180, 90, 140], [33, 77, 151]]], dtype=np.int32)
    est = motion.PoseEstimator(body, hand)
    for rep in range(2):                                          # second round: cached plans / tables
        pose = est.submit_batch(frames, fixed_boxes=boxes) or est.collect()
        assert pose.shape == (n, 60, 3)
        results = body.batch(frames)
        for f in range(n):
            ref = np.zeros((60, 3))
            ref[:18] = extract.body_pose(*results[f])[0]
            for k, is_left in ((0, True), (1, False)):
                x, y, w = (int(v) for v in boxes[f, k])
                if w <= 0:
                    continue
                crop = frames[f, y:y + w, x:x + w]
                peaks = hand(np.ascontiguousarray(crop[:, ::-1]) if is_left else np.ascontiguousarray(crop))
                extract._apply_hand(ref, peaks, x, y, w, is_left)
            assert np.array_equal(pose[f], ref), (rep, f)
        assert (pose[:, 18:, 2] > 0).any()                        # hands were really estimated


def test_pose_single_frame_other_size_and_bad_boxes():
    """est(frame) for one frame; a second frame size on the same estimators (the crop-size tables are rebuilt for the
    larger frame); fixed boxes that leave the frame are treated as missing, not read out of bounds."""
    from pytorch_openpose_b200 import Body, Hand, extract, motion
    rng = np.random.default_rng(11)
    body = Body(O.make_weights("body", 0), scale_search=[0.5])
    hand = Hand(O.make_weights("hand", 5, "kaiming"), scale_search=[0.5, 1.0])
    est = motion.PoseEstimator(body, hand)
    small = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    pose = est(small)
    assert pose.shape == (60, 3) and np.array_equal(pose[:18], extract.body_pose(*body(small))[0])
    big = rng.integers(0, 256, (2, 300, 260, 3), dtype=np.uint8)
    boxes = np.array([[[200, 10, 100], [0, 0, 260]],            # left: x + w > W -> missing; right: the largest possible box
                      [[-5, 20, 50], [100, 250, 60]]], dtype=np.int32)      # negative x; y + w > H
    est.submit_batch(big, fixed_boxes=boxes)
    pose = est.collect()
    assert not pose[0, 18:39].any() and not pose[1, 18:].any()
    ref = np.zeros((60, 3))
    extract._apply_hand(ref, hand(np.ascontiguousarray(big[0, 0:260, 0:260])), 0, 0, 260, False)
    assert np.array_equal(pose[0, 39:], ref[39:])


def test_bodyhand_video_job_on_device(tmp_path):
    """mode='bodyhand' with this package's own Body and Hand runs the device pipeline; its track equals the frame by
    frame host caller (random weights find nobody, so this pins the body rows and the plumbing; the hand path is pinned
    by the test above)."""
    import cv2
    import joblib
    from pytorch_openpose_b200 import Body, Hand, extract, motion
    from tests.test_gpu_extract import _video
    path = str(tmp_path / "v.avi")
    _video(path, 7, 320, 240, 5)
    body = Body(O.make_weights("body", 2, "kaiming"), scale_search=[0.5, 1.0])
    hand = Hand(O.make_weights("hand", 5, "kaiming"), scale_search=[0.5, 1.0])
    mat = extract.extract_motion_from_video(path, str(tmp_path / "o.pkl"), None, body, hand, "bodyhand", batch=4, sessions=2,
                                            log=lambda m: None)
    assert mat.shape == (7, 60, 3) and np.array_equal(joblib.load(str(tmp_path / "o.pkl")), mat)
    cap = cv2.VideoCapture(path)
    for i in range(7):
        ok, frame = cap.read()
        assert ok
        pose, _, _ = motion.pose_mat_every_frame(frame, body, hand, "bodyhand")
        assert np.array_equal(mat[i], pose)


def test_pose_pipeline_grows_result_buffers():
    """Frames whose peaks / limb pairs / person rows overflow the (here: tiny) initial buffers: the device pipeline grows
    them, repeats the body post-processing of those frames and the hand half of the batch, and still equals the host
    pipeline -- also on the following calls (re-captured graph)."""
    import os
    import subprocess
    import sys
    code = r'''
import numpy as np, cv2
from oracle import openpose_oracle as O
from pytorch_openpose_b200 import Body, Hand, motion
rng = np.random.default_rng(21)
frames = np.stack([cv2.GaussianBlur(rng.integers(0, 256, (169, 439, 3), dtype=np.uint8), (0, 0), 1.5) for _ in range(3)])
body = Body(O.make_weights("body", 2, "kaiming"), scale_search=[1.0])
hand = Hand(O.make_weights("hand", 5, "kaiming"), scale_search=[0.5, 1.0])
est = motion.PoseEstimator(body, hand)
persons = 0
for rep in range(4):
    pose = est(frames)
    for f in range(3):
        ref, cand, sub = motion.pose_mat_every_frame(frames[f], body, hand, "bodyhand")
        assert len(cand) > 64, len(cand)
        persons += len(sub)
        assert np.array_equal(pose[f], ref), (rep, f)
print("POSE-GROWTH-OK persons", persons)
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OPB_TEST_SMALL_BUFFERS="1", PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert "POSE-GROWTH-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
