"""Host-side helpers that need no GPU."""
import numpy as np

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
