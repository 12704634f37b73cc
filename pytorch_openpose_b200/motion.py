"""Per-frame body + two-hand key-point record (`PoseMat`), the call pattern of the reference's primary caller
(srcmx/MotionEstimation.py:126-216) on top of the B200 `Body` / `Hand`.

PoseMat rows: 0-17 body, 18-38 left hand, 39-59 right hand; columns x, y, score; zeros mean "missing"."""
import numpy as np

from . import util


def select_person(candidate, subset):
    """Index of the person with the largest left-shoulder x (srcmx/MotionEstimation.py:144-150); a missing
    shoulder (-1) indexes the LAST candidate, exactly like the reference's candidate[-1]."""
    if len(subset) < 1:
        return None
    xs = np.array([candidate[int(person[5])][0] for person in subset])
    return int(np.argmax(xs))


def pose_mat_every_frame(oriImg, body_estimation, hand_estimation=None, mode="body"):
    """-> (PoseMat (60,3) float64, candidate, subset).  `mode` is 'body' or 'bodyhand'."""
    candidate, subset = body_estimation(oriImg)
    pose = np.zeros((60, 3))
    chosen = select_person(candidate, subset)
    if chosen is not None:
        for k in range(18):
            idx = int(subset[chosen][k])
            if idx != -1:
                pose[k, :] = candidate[idx][:3]
    for i in range(len(subset)):                      # blank the other persons (MotionEstimation.py:160-162)
        if i != chosen:
            subset[i, :] = -1
    if mode == "bodyhand":
        if hand_estimation is None:
            raise ValueError("mode='bodyhand' needs a hand estimator")
        for x, y, w, is_left in util.handDetect(candidate, subset, oriImg):
            crop = oriImg[y:y + w, x:x + w, :]
            if is_left:
                # the hand net detects right hands: mirror, then un-mirror x (MotionEstimation.py:191-194)
                peaks = hand_estimation(np.ascontiguousarray(crop[:, ::-1, :]))
                peaks[:, 0] = np.where(peaks[:, 0] == 0, peaks[:, 0], w - peaks[:, 0] - 1 + x)
                peaks[:, 1] = np.where(peaks[:, 1] == 0, peaks[:, 1], peaks[:, 1] + y)
                pose[18:39, :] = peaks
            else:
                peaks = hand_estimation(crop)
                peaks[:, 0] = np.where(peaks[:, 0] == 0, peaks[:, 0], peaks[:, 0] + x)
                peaks[:, 1] = np.where(peaks[:, 1] == 0, peaks[:, 1], peaks[:, 1] + y)
                pose[39:60, :] = peaks
    return pose, candidate, subset
