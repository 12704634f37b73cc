"""`Hand(model_path)(oriImg) -> peaks (21, 3)` -- drop-in for the reference's src/hand.py:16-75 on the GPU."""
import numpy as np

from . import _lib
from .body import _load_checkpoint

DEFAULT_SCALE_SEARCH = (0.5, 1.0, 1.5, 2.0)        # src/hand.py:26


class Hand(object):
    def __init__(self, model_path, scale_search=None, device=None):
        weights = model_path if isinstance(model_path, dict) else _load_checkpoint(model_path)
        self.scale_search = list(scale_search) if scale_search is not None else list(DEFAULT_SCALE_SEARCH)
        self.net = _lib.Net(_lib.NET_HAND, weights, device)
        self._session = self.net.session()

    def submit(self, crops, session=None, where=0):
        """crops: (h, w, 3) or a batch (n, h, w, 3) of equally sized uint8 BGR crops."""
        s = session or self._session
        if where == 1:
            ptr, (n, h, w) = crops
        else:
            arr = np.ascontiguousarray(crops, dtype=np.uint8)
            if arr.ndim == 3:
                arr = arr[None]
            if arr.shape[1] == 0:
                raise ZeroDivisionError("float division by zero")        # src/hand.py:32
            if arr.ndim != 4 or arr.shape[3] != 3:
                raise ValueError("expected (h, w, 3) or (n, h, w, 3) uint8 BGR crops")
            s._keepalive = arr
            ptr, (n, h, w) = arr.ctypes.data, arr.shape[:3]
        s._n = n
        sc, ns = _lib.scales_array(self.scale_search)
        _lib.check(_lib.lib().opb_hand_submit(s.handle, ptr, where, n, h, w, sc, ns))

    def collect(self, session=None):
        s = session or self._session
        peaks = np.empty((s._n, 21, 3), dtype=np.float64)
        _lib.check(_lib.lib().opb_hand_wait(s.handle, peaks.ctypes.data))
        return peaks

    def last_maps(self, shape, session=None):
        """heatmap_avg (n, h, w, 22) float32 of the last finished batch (src/hand.py:57)."""
        s = session or self._session
        h, w = shape[-3:-1]
        heat = np.empty((s._n, 22, h, w), dtype=np.float32)
        _lib.check(_lib.lib().opb_hand_maps(s.handle, heat.ctypes.data))
        return np.ascontiguousarray(heat.transpose(0, 2, 3, 1))

    MAX_BATCH = 32          # crops per submit: ~0.7 GB of activations per 4-scale crop

    def __call__(self, oriImg):
        batched = np.ndim(oriImg) == 4
        if batched and len(oriImg) > self.MAX_BATCH:
            # large batches (BASELINE config 3: 256 crops) run as chunks, double-buffered over two sessions
            if not hasattr(self, "_session2"):
                self._session2 = self.net.session()
            sessions = (self._session, self._session2)
            out, pending = [], []
            for i, lo in enumerate(range(0, len(oriImg), self.MAX_BATCH)):
                s = sessions[i % 2]
                if len(pending) == 2:
                    out.append(self.collect(pending.pop(0)))
                self.submit(oriImg[lo:lo + self.MAX_BATCH], s)
                pending.append(s)
            for s in pending:
                out.append(self.collect(s))
            return np.concatenate(out, 0)
        self.submit(oriImg)
        peaks = self.collect()
        return peaks if batched else peaks[0]
