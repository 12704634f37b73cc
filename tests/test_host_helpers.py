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
