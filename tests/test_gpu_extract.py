"""The extraction job on the real estimators (cuda:0): decode ring -> batched Body on several sessions -> hands -> file
equals the frame-by-frame caller (pytorch_openpose_b200/motion.py, srcmx/MotionEstimation.py:126-216)."""
import numpy as np
import pytest

from oracle import openpose_oracle as O

pytestmark = pytest.mark.gpu


def _video(path, n, w, h, seed):
    import cv2
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 25, (w, h))
    assert wr.isOpened()
    rng = np.random.default_rng(seed)
    for _ in range(n):
        wr.write(cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 4))
    wr.release()


def _with_person(result, k):
    """Random-init networks find nobody; give frame results a person with both arms (joints move with k) so that the
    hand path of the job really runs."""
    candidate, subset = result
    if len(subset):
        return result
    joints = {2: (150, 60), 3: (120 + k % 7, 100), 4: (100 + 2 * (k % 5), 150 + k % 9),
              5: (200, 60), 6: (230 - k % 6, 100), 7: (250 - 2 * (k % 4), 150 + k % 8)}
    cand = np.zeros((8, 4))
    row = -np.ones(20)
    for i, (j, (x, y)) in enumerate(sorted(joints.items())):
        cand[i] = (x, y, 0.9, i)
        row[j] = i
    row[18], row[19] = 6.0, 6
    return cand, row[None].copy()


@pytest.mark.parametrize("mode", ["body", "bodyhand"])
def test_video_job_equals_frame_by_frame(tmp_path, mode):
    import cv2
    import joblib
    from pytorch_openpose_b200 import Body, Hand, extract, motion

    class BodyP(Body):                                   # same estimator, plus the injected person (keyed by the pixels)
        def collect_batch(self, session=None):
            s = session or self._session
            keys = [int(f[5, 7, 0]) for f in np.asarray(s._keepalive)]
            return [_with_person(r, k) for r, k in zip(Body.collect_batch(self, session), keys)]

        def __call__(self, oriImg):
            return _with_person(Body.__call__(self, oriImg), int(oriImg[5, 7, 0]))

    path = str(tmp_path / "v.avi")
    _video(path, 11, 320, 240, 3)
    rec = [(16, 8), (304, 232)]
    body = BodyP(O.make_weights("body", 2, "kaiming"), scale_search=[0.5, 1.0])
    hand = Hand(O.make_weights("hand", 5, "kaiming"), scale_search=[0.5, 1.0])
    mat = extract.extract_motion_from_video(path, str(tmp_path / "o.pkl"), rec, body, hand, mode, batch=4, sessions=2,
                                            log=lambda m: None)
    assert mat.shape == (11, 60 if mode == "bodyhand" else 18, 3)
    assert np.array_equal(joblib.load(str(tmp_path / "o.pkl")), mat)
    if mode == "bodyhand":
        assert (mat[:, 18:39, 2] > 0).any() and (mat[:, 39:, 2] > 0).any()      # both hands were really estimated
    cap = cv2.VideoCapture(path)
    for i in range(11):
        ok, frame = cap.read()
        assert ok
        pose, _, _ = motion.pose_mat_every_frame(frame[8:232, 16:304], body, hand, mode)
        assert np.array_equal(mat[i], pose[:mat.shape[1]])


def test_batch_body_job(tmp_path):
    import joblib
    from pytorch_openpose_b200 import Batch_body, extract
    path = str(tmp_path / "v.avi")
    _video(path, 9, 320, 240, 4)
    est = Batch_body(O.make_weights("body", 2, "kaiming"))
    mat = extract.batch_body_extraction(path, str(tmp_path / "b.pkl"), 4, None, est, log=lambda m: None)
    assert mat.shape == (9, 18, 3) and np.array_equal(joblib.load(str(tmp_path / "b.pkl")), mat)
    import cv2
    cap = cv2.VideoCapture(path)
    for i in range(9):
        ok, frame = cap.read()
        (cand, sub), = est(extract.to_tensor(frame[None]))
        assert np.array_equal(mat[i], extract.body_pose(cand, sub)[0])
