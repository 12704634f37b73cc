"""Host glue of the per-frame body + hands record against the reference's own caller code
(srcmx/MotionEstimation.py:126-216), with deterministic stand-in estimators so that the test runs on CPU."""
import numpy as np
import pytest

from oracle import openpose_oracle as O
from oracle import reference_loader as RL
from pytorch_openpose_b200.motion import pose_mat_every_frame, select_person


class FakeBody(object):
    def __init__(self, candidate, subset):
        self.c, self.s = candidate, subset

    def __call__(self, img):
        return self.c.copy(), self.s.copy()


class FakeHand(object):
    def __call__(self, crop):
        h, w = crop.shape[:2]
        rng = np.random.default_rng(int(crop.sum()) % 1000)
        peaks = np.zeros((21, 3))
        peaks[:, 0] = rng.integers(0, w, 21)
        peaks[:, 1] = rng.integers(0, h, 21)
        peaks[:, 2] = rng.random(21)
        peaks[5] = 0
        return peaks


def _scene(golden):
    g = golden("body_postproc")
    return g["cand_p8"].copy(), g["subset_p8"].copy()


def test_person_selection_and_body_rows(golden):
    cand, subset = _scene(golden)
    img = np.random.default_rng(0).integers(0, 256, (360, 640, 3), dtype=np.uint8)
    pose, c2, s2 = pose_mat_every_frame(img, FakeBody(cand, subset), mode="body")
    chosen = select_person(cand, subset)
    xs = [cand[int(p[5])][0] for p in subset]
    assert chosen == int(np.argmax(xs))
    for k in range(18):
        idx = int(subset[chosen][k])
        assert np.array_equal(pose[k], cand[idx][:3] if idx != -1 else np.zeros(3))
    assert (pose[18:] == 0).all()
    assert all((s2[i] == -1).all() for i in range(len(s2)) if i != chosen)


def test_bodyhand_matches_restated_caller(golden):
    cand, subset = _scene(golden)
    img = np.random.default_rng(1).integers(0, 256, (360, 640, 3), dtype=np.uint8)
    hand = FakeHand()
    pose, _, s2 = pose_mat_every_frame(img, FakeBody(cand, subset), hand, mode="bodyhand")
    # restatement of srcmx/MotionEstimation.py:164-194 with the oracle's handDetect
    exp = pose.copy()
    exp[18:] = 0
    for x, y, w, is_left in O.hand_detect(cand, s2, 360, 640):
        crop = img[y:y + w, x:x + w, :]
        if is_left:
            pk = hand(np.ascontiguousarray(crop[:, ::-1, :]))
            pk[:, 0] = np.where(pk[:, 0] == 0, pk[:, 0], w - pk[:, 0] - 1 + x)
            pk[:, 1] = np.where(pk[:, 1] == 0, pk[:, 1], pk[:, 1] + y)
            exp[18:39] = pk
        else:
            pk = hand(crop)
            pk[:, 0] = np.where(pk[:, 0] == 0, pk[:, 0], pk[:, 0] + x)
            pk[:, 1] = np.where(pk[:, 1] == 0, pk[:, 1], pk[:, 1] + y)
            exp[39:60] = pk
    assert np.array_equal(pose, exp)
    assert (pose[18:39, 2] > 0).any() and (pose[39:60, 2] > 0).any()


def test_no_person():
    img = np.zeros((64, 64, 3), np.uint8)
    pose, c, s = pose_mat_every_frame(img, FakeBody(np.array([]), -np.ones((0, 20))), FakeHand(), mode="bodyhand")
    assert pose.shape == (60, 3) and not pose.any()
