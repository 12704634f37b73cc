"""`Body(model_path)(oriImg) -> (candidate, subset)` -- drop-in for the reference's src/body.py:15-212, with every
stage on the GPU behind libopenpose_b200.so (C ABI in include/openpose_b200.h)."""
import ctypes

import numpy as np

from . import _lib

# src/body.py:25-26: the reference hard-codes [0.5] and keeps the 4-scale list in a comment
DEFAULT_SCALE_SEARCH = (0.5,)
FOUR_SCALE_SEARCH = (0.5, 1.0, 1.5, 2.0)


def _load_checkpoint(model_path):
    import torch
    return torch.load(model_path, map_location="cpu")      # src/body.py:20


class Body(object):
    """Same constructor / call contract as the reference class.  Extras that do not break drop-in use:
    `scale_search=` (also settable as an attribute, like the oracle's patched reference) and `device=`."""

    def __init__(self, model_path, scale_search=None, device=None):
        weights = model_path if isinstance(model_path, dict) else _load_checkpoint(model_path)
        self.scale_search = list(scale_search) if scale_search is not None else list(DEFAULT_SCALE_SEARCH)
        self.net = _lib.Net(_lib.NET_BODY, weights, device)
        self._session = self.net.session()

    # ---- asynchronous halves (used by the frame-sharded video pipeline and bench.py) ----
    def submit(self, oriImg, session=None, where=0):
        """Enqueue one frame.  where: 0 pageable host array, 1 device pointer (int), 2 pinned host array."""
        s = session or self._session
        if where == 1:
            ptr, (H, W) = oriImg
        else:
            if oriImg.shape[0] == 0:
                raise ZeroDivisionError("float division by zero")        # x * boxsize / oriImg.shape[0]
            img = np.ascontiguousarray(oriImg, dtype=np.uint8)
            if img.ndim != 3 or img.shape[2] != 3:
                raise ValueError("expected an (H, W, 3) uint8 BGR image")
            s._keepalive = img
            ptr, (H, W) = img.ctypes.data, img.shape[:2]
        arr, n = _lib.scales_array(self.scale_search)
        _lib.check(_lib.lib().opb_body_submit(s.handle, ptr, where, H, W, arr, n))

    def collect(self, session=None):
        s = session or self._session
        L = _lib.lib()
        nc, ns = ctypes.c_int(), ctypes.c_int()
        _lib.check(L.opb_body_wait(s.handle, ctypes.byref(nc), ctypes.byref(ns)))
        candidate = np.empty((nc.value, 4), dtype=np.float64)
        subset = np.empty((ns.value, 20), dtype=np.float64)
        _lib.check(L.opb_body_fetch(s.handle, candidate.ctypes.data, nc.value, subset.ctypes.data, ns.value))
        if nc.value == 0:
            candidate = np.array([])          # np.array([]) has shape (0,), src/body.py:160
        return candidate, subset

    def __call__(self, oriImg):
        self.submit(oriImg)
        return self.collect()

    # ---- batches of equally sized frames: one launch per CNN layer for the whole batch ----
    def submit_batch(self, frames, session=None, where=0):
        """frames: (n, H, W, 3) uint8 array (where 0 / 2) or (device pointer, (n, H, W)) (where 1)."""
        s = session or self._session
        if where == 1:
            ptr, (n, H, W) = frames
        else:
            arr = np.ascontiguousarray(frames, dtype=np.uint8)
            if arr.ndim != 4 or arr.shape[3] != 3:
                raise ValueError("expected an (n, H, W, 3) uint8 BGR array")
            if arr.shape[1] == 0:
                raise ZeroDivisionError("float division by zero")
            s._keepalive = arr
            ptr, (n, H, W) = arr.ctypes.data, arr.shape[:3]
        s._batch = n
        arr_s, ns = _lib.scales_array(self.scale_search)
        _lib.check(_lib.lib().opb_body_submit_batch(s.handle, ptr, where, n, H, W, arr_s, ns))

    def collect_batch(self, session=None):
        """-> list of (candidate, subset), one per frame.  Raises IndexError if any frame hit the reference's
        src/body.py:173 edge (like calling the reference frame by frame would)."""
        s = session or self._session
        L = _lib.lib()
        n = s._batch
        nc, ns, st = (ctypes.c_int * n)(), (ctypes.c_int * n)(), (ctypes.c_int * n)()
        _lib.check(L.opb_body_wait_batch(s.handle, nc, ns, st))
        out = []
        for f in range(n):
            candidate = np.empty((nc[f], 4), dtype=np.float64)
            subset = np.empty((ns[f], 20), dtype=np.float64)
            _lib.check(L.opb_body_fetch_frame(s.handle, f, candidate.ctypes.data, nc[f], subset.ctypes.data, ns[f]))
            out.append((np.array([]) if nc[f] == 0 else candidate, subset))
        return out

    def batch(self, frames):
        self.submit_batch(frames)
        return self.collect_batch()

    def last_maps(self, shape, session=None):
        """(heatmap_avg (H,W,19), paf_avg (H,W,38)) float32 of the last finished frame (src/body.py:67-68)."""
        s = session or self._session
        if len(shape) == 4:                                  # (n, H, W, 3): maps of every frame of the last batch
            n, H, W = shape[:3]
            heat = np.empty((n, 19, H, W), dtype=np.float32)
            paf = np.empty((n, 38, H, W), dtype=np.float32)
            _lib.check(_lib.lib().opb_body_maps(s.handle, heat.ctypes.data, paf.ctypes.data))
            return np.ascontiguousarray(heat.transpose(0, 2, 3, 1)), np.ascontiguousarray(paf.transpose(0, 2, 3, 1))
        H, W = shape[:2]
        heat = np.empty((19, H, W), dtype=np.float32)
        paf = np.empty((38, H, W), dtype=np.float32)
        _lib.check(_lib.lib().opb_body_maps(s.handle, heat.ctypes.data, paf.ctypes.data))
        return np.ascontiguousarray(heat.transpose(1, 2, 0)), np.ascontiguousarray(paf.transpose(1, 2, 0))
