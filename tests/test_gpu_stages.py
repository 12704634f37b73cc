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


# ---- crowds beyond the default buffer sizes --------------------------------------------------------------------------
def _candidates_from_joints(people, score=0.9):
    """part-major candidate list [x, y, score, id] (ids cumulative, like src/body.py:88-92) from joint positions."""
    peaks, nid = [], 0
    for part in range(18):
        pts = np.rint(people[:, part, :]).astype(np.float64)
        order = np.lexsort((pts[:, 0], pts[:, 1]))                     # np.nonzero order: y, then x
        arr = np.zeros((len(pts), 4))
        arr[:, :2] = pts[order]
        arr[:, 2] = score + 0.0001 * np.arange(len(pts))
        arr[:, 3] = nid + np.arange(len(pts))
        nid += len(pts)
        peaks.append(arr)
    return peaks


def _group_on_device(paf, peaks, subset_cap, conn_cap):
    from tests import gpu_util as G
    cand = np.concatenate(peaks)
    pb = [0]
    for p in peaks:
        pb.append(pb[-1] + len(p))
    cand_dev = torch.from_numpy(cand).cuda()
    return G.group_limbs(paf.transpose(2, 0, 1), cand_dev, pb, subset_cap=subset_cap, conn_cap=conn_cap)


def test_grouping_more_persons_than_shared_memory_rows():
    """1056 people in one frame: more person rows than the 1024 the assembly kernel keeps in shared memory -> the rows
    spill to a global work buffer; subsets and every connection list equal the oracle (the reference has no limit)."""
    _, paf, people = O.synthetic_scene(1656, 1452, (44, 24), seed=3, jitter=0.3)
    peaks = _candidates_from_joints(people)
    conns_ref, special = O.match_limbs(peaks, paf, 1656)
    _, subset_ref = O.assemble(peaks, conns_ref, special)
    subset, conns, cc = _group_on_device(paf, peaks, subset_cap=4096, conn_cap=2048)
    assert len(subset_ref) > 1024
    for k in range(19):
        assert cc[k] == len(conns_ref[k]) and np.array_equal(conns[k, :cc[k]], conns_ref[k])
    assert subset.shape == subset_ref.shape and np.array_equal(subset, subset_ref)


def test_grouping_survivor_lists_grow_and_sort_across_the_gpu():
    """A limb field that accepts almost every pair pointing its way: tens of thousands of scored survivors per limb
    (> the 16384 the lists start with) -> the lists grow, the segmented sort ranks them on many CTAs, the greedy walk
    keeps min(nA, nB) of them; equal to the oracle's stable sort + walk."""
    H, W, n = 400, 600, 220
    rng = np.random.default_rng(17)
    people = np.zeros((n, 18, 2))
    people[:, :, 0] = rng.integers(5, W - 5, (n, 18))
    people[:, :, 1] = rng.integers(5, H - 5, (n, 18))
    # distinct positions per part (peaks of one map are distinct pixels)
    for part in range(18):
        flat = rng.choice((H - 10) * (W - 10), n, replace=False)
        people[:, part, 0] = 5 + flat % (W - 10)
        people[:, part, 1] = 5 + flat // (W - 10)
    peaks = _candidates_from_joints(people)
    paf = np.zeros((H, W, 38))
    paf[:, :, 0::2] = 1.0                                              # every limb's field points along +x
    conns_ref, special = O.match_limbs(peaks, paf, H)
    survivors = [len(O.score_pairs(paf, peaks[a - 1], peaks[b - 1], k, H)) for k, (a, b) in enumerate(O.LIMB_SEQ[:2])]
    assert max(survivors) > 16384, survivors
    try:
        _, subset_ref = O.assemble(peaks, conns_ref, special)
        raised = False
    except IndexError:
        raised = True
    from pytorch_openpose_b200 import _lib
    if raised:
        with pytest.raises(IndexError):
            _group_on_device(paf, peaks, subset_cap=4096, conn_cap=512)
    else:
        subset, conns, cc = _group_on_device(paf, peaks, subset_cap=4096, conn_cap=512)
        for k in range(19):
            assert cc[k] == len(conns_ref[k]) and np.array_equal(conns[k, :cc[k]], conns_ref[k])
        assert np.array_equal(subset, subset_ref)
