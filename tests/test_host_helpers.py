"""Host-side helpers that need no GPU."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import openpose_oracle as O


def test_random_checkpoint_is_the_reference_default_init():
    """`model.random_checkpoint(kind, seed)` == what `torch.manual_seed(seed); bodypose_model()` holds, in the caffe-keyed
    checkpoint format (the oracle's make_weights is pinned to the live reference in test_oracle_vs_reference.py)."""
    import torch
    from pytorch_openpose_b200 import model
    for kind in ("body", "hand"):
        a, b = model.random_checkpoint(kind, 0), O.make_weights(kind, 0)
        assert set(a.keys()) == set(b.keys())
        assert all(torch.equal(a[k], b[k]) for k in a)


def test_bench_line_helpers_parse():
    """bench.py's module-level helpers work without a GPU (the driver imports nothing, but the reference arm runs here)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    peak, hbm, how = mod.measured_peaks()
    assert peak > 100 and hbm > 1000 and isinstance(how, str)
    frames = mod.synth_frames(2, 1)
    assert frames.shape == (2, 720, 1280, 3) and frames.dtype == np.uint8


def test_wide_pixel_weights_describe_the_same_layer():
    """conv1_2 runs as a 128 -> 128 channel 3x3 layer on column PAIRS.  The library's packed weights
    (host code of the .so, no device involved) applied to the re-described image must give the original layer's output:
    checked here with an independent numpy convolution on bf16-representable data."""
    import ctypes
    from pytorch_openpose_b200 import _lib
    rng = np.random.default_rng(3)

    def bf16(a):                                     # round to bf16, keep float32
        u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
        u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
        return u.astype(np.uint32).view(np.float32)

    w = bf16(rng.standard_normal((64, 64, 3, 3)).astype(np.float32) * 0.1)
    b = rng.standard_normal(64).astype(np.float32)
    ww = np.zeros((128, 9, 128), dtype=np.uint16)
    bw = np.zeros(128, dtype=np.float32)
    _lib.check(_lib.lib().opb_wide_pool_weights(w.ctypes.data, b.ctypes.data, ww.ctypes.data, bw.ctypes.data))
    wwf = (ww.astype(np.uint32) << 16).view(np.float32).reshape(128, 3, 3, 2, 64)     # (po co, dy, di, pi, c)
    assert np.array_equal(bw, np.concatenate([b, b]))
    # all-zero chunks the kernel skips, and the two half chunks
    assert not wwf[:, :, 0, 0, :].any() and not wwf[:, :, 2, 1, :].any()
    assert not wwf[64:, :, 0, 1, :].any() and not wwf[:64, :, 2, 0, :].any()

    H, W = 6, 10
    x = bf16(rng.standard_normal((H, W, 64)).astype(np.float32))
    xp = np.pad(x, ((1, 1), (1, 1), (0, 0)))
    ref = np.zeros((H, W, 64))
    for dy in range(3):
        for dx in range(3):
            ref += xp[dy:dy + H, dx:dx + W, :].astype(np.float64) @ w[:, :, dy, dx].astype(np.float64).T
    ref += b
    xw = x.reshape(H, W // 2, 2, 64)                 # wide pixels: (parity, channel)
    xwp = np.pad(xw, ((1, 1), (1, 1), (0, 0), (0, 0)))
    out = np.zeros((H, W // 2, 128))
    for dy in range(3):
        for di in range(3):
            a = xwp[dy:dy + H, di:di + W // 2].reshape(H, W // 2, 128).astype(np.float64)
            out += a @ wwf[:, dy, di].reshape(128, 128).astype(np.float64).T
    out += bw
    assert np.allclose(out.reshape(H, W // 2, 2, 64).reshape(H, W, 64), ref, rtol=0, atol=1e-9)


def _pair_tiles(n, h, w, n_tiles_n=1, small=0):
    import ctypes
    from pytorch_openpose_b200 import _lib
    cap = 8 * 4 * (n * ((h + 7) // 8) * ((w + 15) // 16) + 8) * n_tiles_n
    out = np.zeros(cap, dtype=np.int32)
    written = ctypes.c_int(0)
    _lib.check(_lib.lib().opb_debug_pair_tiles(n, h, w, n_tiles_n, small, out.ctypes.data, cap, ctypes.byref(written)))
    return out[:written.value * 8].reshape(-1, 2, 8)           # (pair, rank, fields)


def _check_tiling(n, h, w, n_tiles_n=1, small=0):
    """Every pixel of every image is stored by exactly one (tile, half) per N tile; the two tiles of a pair share N tile and
    orientation (the pair's single A descriptor); the halves the pair multiplies are a superset of what each tile stores."""
    t = _pair_tiles(n, h, w, n_tiles_n, small)
    cover = np.zeros((n_tiles_n, n, h, w), dtype=np.int32)
    multiplied = own = 0
    for pair in t:
        live = [r for r in pair if r[4]]
        assert live, "a pair of two padding tiles"
        assert len({(r[3], r[6]) for r in pair}) == 1, "tiles of a pair differ in N tile or orientation"
        halves = pair[0][5] | pair[1][5]
        multiplied += 2 * bin(halves).count("1") * 128
        for img, x0, y0, n0, real, hv, vs, full in pair:
            if not real:
                assert hv == 0
                continue
            own += bin(hv).count("1") * 128
            th = 8 if small else 16
            for hh in range(2):
                if not (hv >> hh) & 1:
                    continue
                if small:
                    xs, ys, ww, hhh = x0, y0, 16, 8
                elif vs:
                    xs, ys, ww, hhh = x0, y0 + 8 * hh, 16, 8
                else:
                    xs, ys, ww, hhh = x0 + 8 * hh, y0, 8, 16
                cover[n0 // 128, img, ys:min(ys + hhh, h), xs:min(xs + ww, w)] += 1
            assert y0 % th == 0 and x0 % 16 == 0
    assert (cover == 1).all(), "pixels stored %d..%d times" % (cover.min(), cover.max())
    total = float(n_tiles_n * n * h * w)
    return own / total, multiplied / total       # rows of the halves that hold pixels / rows the pairs actually multiply


def test_conv_pair_tiling_covers_every_pixel_once():
    """The tile decode of the CTA-pair kernel (host copy of the device function), over the shapes the networks produce and
    a sweep of small ones: exact cover, consistent pairs, and the padding the design document quotes."""
    body = [(23, 41), (46, 82), (69, 123), (92, 164)]
    hand = [(23, 23), (46, 46), (69, 69), (92, 92)]
    for n in (1, 3, 8):
        for h, w in body + hand:
            _check_tiling(n, h, w)
    pix = lambda shapes: float(sum(h * w for h, w in shapes))
    ratio = lambda shapes, n, k: sum(_check_tiling(n, h, w)[k] * h * w for h, w in shapes) / pix(shapes)
    assert abs(ratio(body, 1, 0) - 1.086) < 2e-3 and abs(ratio(hand, 1, 0) - 1.097) < 2e-3
    # what pairing adds on top (a half one tile of a pair needs is multiplied for both; class padding tiles)
    assert ratio(body, 8, 1) < 1.10 and ratio(hand, 16, 1) < 1.10
    for h in range(1, 41):
        for w in range(1, 41):
            _check_tiling(2, h, w)
    for h, w in [(8, 8), (9, 17), (16, 16), (24, 24), (33, 8), (8, 33)]:
        _check_tiling(3, h, w, n_tiles_n=2)
        _check_tiling(1, h, w, small=1)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 9), st.integers(1, 200), st.integers(1, 200), st.integers(1, 4), st.booleans())
def test_conv_pair_tiling_property(n, h, w, n_tiles_n, small):
    _check_tiling(n, h, w, n_tiles_n, int(small))


def _lib_resize_taps(src, dst, scale):
    from pytorch_openpose_b200 import _lib
    first = np.zeros(dst, dtype=np.int32)
    coef = np.zeros((dst, 4), dtype=np.float32)
    _lib.check(_lib.lib().opb_debug_resize_taps(src, dst, float(scale), first.ctypes.data, coef.ctypes.data))
    return first, coef


@settings(max_examples=60, deadline=None)
@given(src=st.integers(1, 1500), f=st.floats(0.05, 9.0))
def test_library_cubic_taps_equal_the_oracles_bit_for_bit(src, f):
    """The tables the kernels apply are built by host code of the .so; they must be the oracle's restatement of cv2's
    float32 tap arithmetic exactly (the oracle itself is pinned against cv2 in test_oracle_golden / test_properties)."""
    from pytorch_openpose_b200 import _lib
    dst = O.resize_dsize(src, f)
    assert _lib.lib().opb_debug_resize_dsize(src, float(f)) == dst
    if dst < 1:
        return
    first, coef = _lib_resize_taps(src, dst, 1.0 / f)
    ofirst, ocoef = O.cubic_taps(src, dst, 1.0 / f)
    assert np.array_equal(first, ofirst)
    assert np.array_equal(coef.view(np.uint32), np.asarray(ocoef, dtype=np.float32).view(np.uint32))


@settings(max_examples=60, deadline=None)
@given(n_net=st.integers(1, 170), crop=st.integers(0, 7), orig=st.integers(1, 1400))
def test_library_composite_operator_equals_the_oracles(n_net, crop, orig):
    """x8 cubic upsample -> crop -> cubic resize as ONE banded operator per axis (what the fused peak kernel, the
    materialiser and the PAF sampler evaluate): equal to the product of the oracle's two cubic matrices."""
    from pytorch_openpose_b200 import _lib
    n_resized = 8 * n_net - crop
    if n_resized < 1:
        return
    # the network sees the frame resized by 184..736 / H: the way back is a resize by 0.25..4 (src/body.py:57)
    n_orig = min(max(orig, (n_resized + 3) // 4), 4 * n_resized)
    first = np.zeros(n_orig, dtype=np.int32)
    w6 = np.zeros((n_orig, 6), dtype=np.float32)
    _lib.check(_lib.lib().opb_debug_composite_taps(n_net, n_resized, n_orig, first.ctypes.data, w6.ctypes.data))
    dense = np.zeros((n_orig, n_net + 6))
    for o in range(n_orig):
        dense[o, first[o]:first[o] + 6] = w6[o]
    assert not dense[:, n_net:].any(), "weights beyond the last net column"
    M = O.composite_upsample_matrix(n_net, n_resized, n_orig)
    assert np.abs(dense[:, :n_net] - M).max() <= 2e-7                     # the library rounds the float64 products to float32
    assert np.abs(dense[:, :n_net].sum(1) - 1).max() < 1e-5
