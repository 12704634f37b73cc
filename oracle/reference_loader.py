"""TEST INFRASTRUCTURE ONLY -- loads the *real* reference (hitmaxiang/pytorch-openpose) from
/root/reference so that the oracle restatement in `oracle/openpose_oracle.py` can be pinned against it
and golden fixtures can be generated (`oracle/make_golden.py`).

/root/reference exists only in the build container, never on the GPU box: nothing under tests marked
`gpu`, `__graft_entry__.smoke()` or `bench.py` may call into this module.  Nothing is copied from the
reference; its files are executed from where they lie.

Recipe:
  * `matplotlib` and `skimage` are imported at module top by src/body.py:6-7, src/hand.py:7-10,
    src/util.py:4-8 but are absent here -> stub modules are registered in sys.modules first.
    `skimage.measure.label` is backed by `scipy.ndimage.label` with a full (8-connected in 2-D)
    structuring element; both number components in raster order of their first pixel.
  * src/body.py:26 hard-codes `scale_search = [0.5]`; the 4-scale BASELINE config needs that single
    line turned into an attribute lookup, which is done on the source text at exec time.
  * The post-processing halves (src/body.py:70-212, src/hand.py:59-75) are inline code, not functions;
    they are sliced out of the source text at run time and wrapped into callables so that the
    reference's own post-processing can run on injected (device-produced) maps.
"""
import os
import sys
import types
import textwrap
import warnings

REFERENCE_ROOT = os.environ.get("OPENPOSE_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "body.py"))


def _install_stubs():
    import numpy as np
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            backends = types.ModuleType("matplotlib.backends")
            agg = types.ModuleType("matplotlib.backends.backend_agg")
            fig = types.ModuleType("matplotlib.figure")
            agg.FigureCanvasAgg = object
            fig.Figure = object
            mpl.pyplot = plt
            mpl.backends = backends
            mpl.figure = fig
            backends.backend_agg = agg
            sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt,
                                "matplotlib.backends": backends,
                                "matplotlib.backends.backend_agg": agg,
                                "matplotlib.figure": fig})
    if "skimage" not in sys.modules:
        try:
            import skimage.measure  # noqa: F401
        except Exception:
            from scipy import ndimage as ndi
            sk = types.ModuleType("skimage")
            measure = types.ModuleType("skimage.measure")

            def label(binary, return_num=False, connectivity=None):
                binary = np.asarray(binary)
                if connectivity is None:
                    connectivity = binary.ndim
                st = ndi.generate_binary_structure(binary.ndim, connectivity)
                lab, n = ndi.label(binary, structure=st)
                return (lab, n) if return_num else lab

            measure.label = label
            sk.measure = measure
            sys.modules.update({"skimage": sk, "skimage.measure": measure})


_cache = {}


def load():
    """Returns a namespace with the reference's Body, Hand, util, model (scale_search patched)."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # `src` might already name something else: make sure it is the reference package
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        if not getattr(sys.modules[k], "__file__", "").startswith(REFERENCE_ROOT):
            del sys.modules[k]
    import src.model as ref_model
    import src.util as ref_util
    import src.hand as ref_hand

    body_path = os.path.join(REFERENCE_ROOT, "src", "body.py")
    text = open(body_path).read()
    needle = "        scale_search = [0.5]\n"
    assert text.count(needle) == 1, "reference body.py changed: cannot patch scale_search"
    text = text.replace(needle, "        scale_search = getattr(self, 'scale_search', [0.5])\n")
    text = text.split("if __name__")[0]
    ref_body = types.ModuleType("src.body")
    ref_body.__file__ = body_path
    exec(compile(text, body_path, "exec"), ref_body.__dict__)

    ns = types.SimpleNamespace(Body=ref_body.Body, Hand=ref_hand.Hand, util=ref_util, model=ref_model,
                               body_module=ref_body, hand_module=ref_hand)
    _cache["ns"] = ns
    return ns


def _slice(path, first, last):
    lines = open(path).read().split("\n")
    return textwrap.dedent("\n".join(lines[first - 1:last]))


def body_postproc():
    """The reference's own src/body.py:70-212 as `f(heatmap_avg, paf_avg, oriImg) -> (candidate, subset)`."""
    if "bpp" in _cache:
        return _cache["bpp"]
    ns = load()
    path = os.path.join(REFERENCE_ROOT, "src", "body.py")
    body = _slice(path, 70, 212)
    assert body.lstrip().startswith("all_peaks = []") and body.rstrip().endswith("return candidate, subset")
    src = ("def postproc(heatmap_avg, paf_avg, oriImg, thre1=0.1, thre2=0.05):\n"
           + textwrap.indent(body, "    ") + "\n")
    g = dict(ns.body_module.__dict__)
    exec(compile(src, path + ":70-212", "exec"), g)
    _cache["bpp"] = g["postproc"]
    return _cache["bpp"]


def hand_postproc():
    """The reference's own src/hand.py:59-75 as `f(heatmap_avg, thre=0.03) -> peaks (21,3)`.
    NOTE: like the reference it zeroes parts of heatmap_avg in place (src/hand.py:71)."""
    if "hpp" in _cache:
        return _cache["hpp"]
    ns = load()
    path = os.path.join(REFERENCE_ROOT, "src", "hand.py")
    body = _slice(path, 59, 75)
    assert body.lstrip().startswith("all_peaks = []") and body.rstrip().endswith("return np.array(all_peaks)")
    src = "def postproc(heatmap_avg, thre=0.03):\n" + textwrap.indent(body, "    ") + "\n"
    g = dict(ns.hand_module.__dict__)
    exec(compile(src, path + ":59-75", "exec"), g)
    _cache["hpp"] = g["postproc"]
    return _cache["hpp"]


def save_checkpoint(model, path):
    """Write a state dict in the caffe-key format util.transfer expects (src/util.py:36-40)."""
    import torch
    sd = {k.split(".", 1)[1]: v.detach().clone() for k, v in model.state_dict().items()}
    torch.save(sd, path)
    return sd


def load_batch():
    """The reference's batched estimators `Batch_body` / `Batch_hand` (srcmx/Batch_model.py:107-406, SURVEY.md 8f row
    N2) and `utilmx`.  srcmx imports numba, h5py and tslearn at module top (Batch_model.py:22, utilmx.py:17,26) for code
    that is not on this path; they are absent here and stubbed."""
    if "batch" in _cache:
        return _cache["batch"]
    load()
    for name in ("numba", "h5py", "tslearn", "tslearn.metrics"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["numba"], "jit"):
        sys.modules["numba"].jit = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))
    if not hasattr(sys.modules["tslearn"], "metrics"):
        sys.modules["tslearn"].metrics = sys.modules["tslearn.metrics"]
    srcmx = os.path.join(REFERENCE_ROOT, "srcmx")
    if srcmx not in sys.path:
        sys.path.insert(0, srcmx)
    import Batch_model
    import utilmx
    ns = types.SimpleNamespace(Batch_body=Batch_model.Batch_body, Batch_hand=Batch_model.Batch_hand, utilmx=utilmx,
                               module=Batch_model)
    _cache["batch"] = ns
    return ns


def batch_body_postproc():
    """The reference's own srcmx/Batch_model.py:173-204 (peak search on the blurred maps, id/score bookkeeping,
    FindBody_frame per frame) as `f(estimator, blurred_heat (B,19,h,w) torch float32, paf (B,38,h,w) numpy float32)
    -> [(candidate, subset)]`, so that it can run on injected (device-produced) maps."""
    if "bbpp" in _cache:
        return _cache["bbpp"]
    b = load_batch()
    path = os.path.join(REFERENCE_ROOT, "srcmx", "Batch_model.py")
    head, tail = _slice(path, 173, 178), _slice(path, 182, 204)
    assert head.lstrip().startswith("batch_peaks = utilmx.findpeaks_torch") and tail.rstrip().endswith("return results")
    src = ("def postproc(self, b_heatmap, b_paf):\n    batch_size = len(b_heatmap)\n"
           + textwrap.indent(head, "    ") + "\n" + textwrap.indent(tail, "    ") + "\n")
    g = dict(b.module.__dict__)
    exec(compile(src, path + ":173-204", "exec"), g)
    _cache["bbpp"] = g["postproc"]
    return _cache["bbpp"]


def batch_hand_postproc():
    """srcmx/Batch_model.py:387-406 as `f(estimator, heatmap (B,h,w,22) numpy float32) -> (B,21,3)`; like the
    reference it zeroes parts of `heatmap` in place (Batch_model.py:400)."""
    if "bhpp" in _cache:
        return _cache["bhpp"]
    b = load_batch()
    path = os.path.join(REFERENCE_ROOT, "srcmx", "Batch_model.py")
    body = _slice(path, 387, 406)
    assert body.lstrip().startswith("Batch_peaks = []") and body.rstrip().endswith("return np.array(Batch_peaks)")
    src = "def postproc(self, heatmap):\n    batch_size = len(heatmap)\n" + textwrap.indent(body, "    ") + "\n"
    g = dict(b.module.__dict__)
    exec(compile(src, path + ":387-406", "exec"), g)
    _cache["bhpp"] = g["postproc"]
    return _cache["bhpp"]


def load_motion_estimation(body_factory, hand_factory):
    """srcmx/MotionEstimation.py with its module-level estimator singletons (`Body('../model/...')`, :18-19) built by
    the given factories instead of from checkpoint files, so that the reference's own extraction job
    (`Extract_MotionData_from_Video`, :25-77) can run on fake estimators.  Returns the freshly imported module."""
    ns = load()
    load_batch()                                   # stubs + srcmx on sys.path
    import src.body
    import src.hand
    keep = (src.body.Body, src.hand.Hand)
    sys.modules.pop("MotionEstimation", None)
    src.body.Body, src.hand.Hand = body_factory, hand_factory
    try:
        import MotionEstimation
    finally:
        src.body.Body, src.hand.Hand = keep
    assert ns.Hand is keep[1]
    return MotionEstimation


def load_batch_motion_estimation():
    """srcmx/Batch_motion_Estimation.py (its estimators are module globals assigned by the caller, :22,60)."""
    load_batch()
    import Batch_motion_Estimation
    return Batch_motion_Estimation
