"""Stage-level parity of the CUDA path against the CPU oracle (oracle/openpose_oracle.py), through the C ABI.
Integer / index work must be bit-exact; floating-point work carries its tolerance in the test."""
import numpy as np
import pytest
import torch

from oracle import openpose_oracle as O
from oracle.make_golden import smooth_noise_maps

pytestmark = pytest.mark.gpu


# ---- (b) preprocessing: bit-exact with the open-source cv2 algorithm -------------------------------
@pytest.mark.parametrize("shape,scale", [((480, 640), 0.5), ((720, 1280), 0.5), ((720, 1280), 2.0), ((97, 131), 1.5),
                                         ((40, 40), 2.0), ((368, 368), 1.0)])
def test_preprocess_bit_exact(shape, scale):
    from tests import gpu_util as G
    img = np.random.default_rng(3).integers(0, 256, shape + (3,), dtype=np.uint8)
    out, (h, w, hp, wp) = G.preprocess(img, scale)
    mult = scale * 368 / shape[0]
    padded, _, pad = O.preprocess(img, mult, use_cv2=False)
    assert out.shape == padded.shape
    assert np.array_equal(out, padded)
    # and within 1 LSB of whatever cv2 build is installed (IPP path)
    padded_cv, _, _ = O.preprocess(img, mult, use_cv2=True)
    assert np.abs(out.astype(int) - padded_cv.astype(int)).max() <= 1


# ---- (b) upsample + average: |err| <= 1e-5 absolute on O(1) maps ---------------------------------------
@pytest.mark.parametrize("H,W,scales", [(480, 640, (0.5,)), (720, 1280, (0.5, 1.0, 1.5, 2.0)), (97, 131, (0.5, 1.5)),
                                        (40, 40, (0.5, 1.0, 1.5, 2.0))])
def test_upsample_avg(H, W, scales):
    from tests import gpu_util as G
    rng = np.random.default_rng(4)
    plan = O.scale_plan(H, W, scales)
    maps = [rng.standard_normal((p["ho"], p["wo"], 19)).astype(np.float32) for p in plan]
    ref = O.upsample_avg([m.transpose(2, 0, 1) for m in maps], plan, H, W, use_cv2=True)       # (H,W,C) f64
    out = G.upsample_avg(maps, scales, H, W).transpose(1, 2, 0)
    assert np.abs(out - ref).max() <= 1e-5


# ---- (c) smoothing is bit-identical to scipy, peaks identical to the reference loop -----------------
def test_gaussian_bit_exact():
    from tests import gpu_util as G
    m = smooth_noise_maps(75, 133, 3, 2, 0.3, 8).transpose(2, 0, 1)
    got = G.smooth(m)
    for c in range(3):
        assert np.array_equal(got[c], O.gaussian_sigma3(m[c].astype(np.float32).astype(np.float64)))
    tiny = np.random.default_rng(1).random((2, 5, 9)).astype(np.float32)       # smaller than the filter radius
    got = G.smooth(tiny)
    for c in range(2):
        assert np.array_equal(got[c], O.gaussian_sigma3(tiny[c].astype(np.float64)))


@pytest.mark.parametrize("tag,H,W,grid", [("p1", 240, 320, (1, 1)), ("p8", 360, 640, (4, 2)), ("p50", 720, 1280, (10, 5))])
def test_postproc_scenes_match_golden(golden, tag, H, W, grid):
    """Device NMS + PAF grouping + assembly on synthetic scenes == the REAL reference's output (golden)."""
    from tests import gpu_util as G
    g = golden("body_postproc")
    heat, paf, _ = O.synthetic_scene(H, W, grid, seed=0)
    cand, pb, cand_dev = G.find_peaks(heat.transpose(2, 0, 1))
    assert np.array_equal(cand, g["cand_" + tag])
    subset, conns, cc = G.group_limbs(paf.transpose(2, 0, 1), cand_dev, pb)
    assert np.array_equal(subset, g["subset_" + tag])


def test_postproc_noise_matches_golden(golden):
    from tests import gpu_util as G
    g = golden("body_postproc")
    heat = smooth_noise_maps(240, 320, 19, 4, 0.12, 11)
    paf = smooth_noise_maps(240, 320, 38, 6, 0.30, 12)
    cand, pb, cand_dev = G.find_peaks(heat.transpose(2, 0, 1))
    assert np.array_equal(cand, g["cand_noise"])
    subset, conns, cc = G.group_limbs(paf.transpose(2, 0, 1), cand_dev, pb)
    assert np.array_equal(subset, g["subset_noise"])
    # every per-limb connection list equals the oracle's (idA, idB, score, i, j), bit for bit
    peaks = O.find_peaks(heat)
    ref_conns, special = O.match_limbs(peaks, paf, 240)
    for k in range(19):
        if k in special:
            assert cc[k] == -1
        else:
            assert cc[k] == len(ref_conns[k]) and np.array_equal(conns[k, :cc[k]], ref_conns[k])


def test_postproc_empty_and_ragged():
    from tests import gpu_util as G
    heat = np.zeros((19, 64, 80), np.float32)
    cand, pb, cand_dev = G.find_peaks(heat)
    assert len(cand) == 0 and pb[18] == 0
    subset, _, cc = G.group_limbs(np.zeros((38, 64, 80), np.float32), cand_dev, pb)
    assert subset.shape == (0, 20) and all(c == -1 for c in cc)
    # a single part present: peaks but every limb has an empty side
    heat[3, 20, 30] = 5.0
    cand, pb, cand_dev = G.find_peaks(heat)
    ref = O.find_peaks(heat.transpose(1, 2, 0).astype(np.float64))
    assert np.array_equal(cand, np.concatenate(ref))
    subset, _, cc = G.group_limbs(np.zeros((38, 64, 80), np.float32), cand_dev, pb)
    assert subset.shape == (0, 20)


def test_peak_capacity_overflow_is_reported():
    from tests import gpu_util as G
    from pytorch_openpose_b200 import _lib
    heat = smooth_noise_maps(240, 320, 19, 1.0, 0.5, 5).transpose(2, 0, 1)
    with pytest.raises(_lib.OpbError) as e:
        G.find_peaks(heat, capacity=64)
    assert e.value.code == _lib.OPB_ERR_CAPACITY


# ---- hand peaks ------------------------------------------------------------------------------------
def test_hand_peaks_match_golden(golden):
    from tests import gpu_util as G
    hm = smooth_noise_maps(184, 184, 22, 5, 0.03, 21)
    hm[:, :, 3] = -1.0
    got = G.hand_peaks(hm.transpose(2, 0, 1))
    assert np.array_equal(got, golden("hand_postproc")["peaks"])


def test_hand_peaks_rectangular_and_empty():
    from tests import gpu_util as G
    hm = smooth_noise_maps(61, 95, 22, 3, 0.04, 33)
    got = G.hand_peaks(hm.transpose(2, 0, 1))
    assert np.array_equal(got, O.hand_postprocess(hm))
    assert np.array_equal(G.hand_peaks(np.zeros((22, 30, 30), np.float32)), np.zeros((21, 3)))


def test_hand_peaks_giant_and_many_components():
    """One component covering the whole map (run-based union-find worst case) and a map with hundreds of
    small components: identical to the oracle."""
    from tests import gpu_util as G
    rng = np.random.default_rng(12)
    hm = np.zeros((120, 150, 22))
    hm[:, :, :11] = 0.05 + smooth_noise_maps(120, 150, 11, 3, 0.002, 40)           # everywhere above 0.03
    hm[:, :, 11:] = smooth_noise_maps(120, 150, 11, 1.0, 0.06, 41)                  # speckle: many components
    hm = hm.astype(np.float32).astype(np.float64)
    got = G.hand_peaks(hm.transpose(2, 0, 1))
    ref = O.hand_postprocess(hm)
    assert np.array_equal(got, ref)
    assert (ref[:, 2] > 0).sum() >= 15
