"""`Batch_body(model_path)(batch_images)` / `Batch_hand(model_path)(batch_imgs)` -- drop-ins for the reference's batched
estimators (srcmx/Batch_model.py:107-406) on libopenpose_b200.so.

Same call contract as the reference: `batch_images` is a float (B, 3, h, w) tensor/array in [0, 1] (what
`transforms.ToTensor()` gives); Batch_body returns `[(candidates, subset)] * B`, Batch_hand an array (B, 21, 3).
The numerics differ from `Body` / `Hand` on purpose, exactly as in the reference: torch-bicubic float resizes, one
scale (0.5), a 5x5 blur instead of the sigma-3 Gaussian, peaks found AND scored on the blurred map, hand threshold
0.035."""
import ctypes

import numpy as np

from . import _lib
from .body import _load_checkpoint


def _as_frames_u8(frames, where):
    """Decoded frames (B, h, w, 3) uint8 as they come out of cv2 -> (pointer, keepalive, (B, h, w))."""
    arr = np.ascontiguousarray(frames, dtype=np.uint8)
    if arr.ndim != 4 or arr.shape[3] != 3:
        raise ValueError("expected (B, h, w, 3) uint8 frames")
    return arr.ctypes.data, arr, arr.shape[:3]


def _as_float_batch(batch):
    """-> (pointer, where, keepalive, shape).  torch CUDA tensors are read in place (where=1), pinned CPU tensors are
    copied straight from their pages (where=2); anything else goes through the library's pinned staging buffer."""
    if hasattr(batch, "detach"):
        t = batch.detach()
        if t.dim() != 4 or t.shape[1] != 3:
            raise ValueError("expected a (B, 3, h, w) float batch")
        import torch
        if t.is_cuda or t.is_pinned():
            t = t.to(torch.float32).contiguous()
            if t.is_cuda:
                torch.cuda.current_stream(t.device).synchronize()       # producer work is on torch's stream
            return t.data_ptr(), (1 if t.is_cuda else 2), t, tuple(t.shape)
        batch = t.numpy()
    arr = np.ascontiguousarray(batch, dtype=np.float32)
    if arr.ndim != 4 or arr.shape[1] != 3:
        raise ValueError("expected a (B, 3, h, w) float batch")
    return arr.ctypes.data, 0, arr, arr.shape


class Batch_body(object):
    MAX_BATCH = 16

    def __init__(self, model_path, device=None):
        weights = model_path if isinstance(model_path, dict) else _load_checkpoint(model_path)
        self.net = _lib.Net(_lib.NET_BODY, weights, device)
        self._session = self.net.session()
        self.scale_search = 0.5                        # srcmx/Batch_model.py:118

    def submit(self, batch_images, session=None):
        s = session or self._session
        ptr, where, s._keepalive, (B, _, h, w) = _as_float_batch(batch_images)
        s._batch, s._shape = B, (B, h, w)
        _lib.check(_lib.lib().opb_batch_body_submit(s.handle, ptr, where, B, h, w, float(self.scale_search)))

    def submit_frames(self, frames_u8, session=None, where=0):
        """Same as `submit(ToTensor(frames))` for decoded (B, h, w, 3) uint8 frames, without the host-side float
        conversion: the division by 255 happens on the device (bit-identical).  where=2: `frames_u8` is pinned."""
        s = session or self._session
        ptr, s._keepalive, (B, h, w) = _as_frames_u8(frames_u8, where)
        s._batch, s._shape = B, (B, h, w)
        _lib.check(_lib.lib().opb_batch_body_submit_u8(s.handle, ptr, where, B, h, w, float(self.scale_search)))

    def collect(self, session=None):
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

    def __call__(self, batch_images):
        results = []
        for lo in range(0, len(batch_images), self.MAX_BATCH):          # activations of 16 frames per launch set
            self.submit(batch_images[lo:lo + self.MAX_BATCH])
            results.extend(self.collect())
        return results

    def last_maps(self, session=None):
        """(blurred heat (B,h,w,19), paf (B,h,w,38)) float32 of the last submitted chunk (Batch_model.py:182-183)."""
        s = session or self._session
        B, h, w = s._shape
        blurred = np.empty((B, 19, h, w), dtype=np.float32)
        paf = np.empty((B, 38, h, w), dtype=np.float32)
        _lib.check(_lib.lib().opb_batch_maps(s.handle, blurred.ctypes.data))
        _lib.check(_lib.lib().opb_body_maps(s.handle, None, paf.ctypes.data))
        return np.ascontiguousarray(blurred.transpose(0, 2, 3, 1)), np.ascontiguousarray(paf.transpose(0, 2, 3, 1))


class Batch_hand(object):
    MAX_BATCH = 64

    def __init__(self, model_path, device=None):
        weights = model_path if isinstance(model_path, dict) else _load_checkpoint(model_path)
        self.net = _lib.Net(_lib.NET_HAND, weights, device)
        self._session = self.net.session()

    def submit(self, batch_imgs, session=None):
        s = session or self._session
        ptr, where, s._keepalive, (B, _, h, w) = _as_float_batch(batch_imgs)
        s._n, s._shape = B, (B, h, w)
        if h % 8 or w % 8:
            raise ValueError("Batch_hand crops must have sides that are multiples of 8: the reference upsamples the "
                             "stride-8 maps by exactly 8 (srcmx/Batch_model.py:377)")
        _lib.check(_lib.lib().opb_batch_hand_submit(s.handle, ptr, where, B, h, w))

    def submit_frames(self, crops_u8, session=None, where=0):
        """`submit(ToTensor(crops))` for (B, h, w, 3) uint8 crops; /255 on the device."""
        s = session or self._session
        ptr, s._keepalive, (B, h, w) = _as_frames_u8(crops_u8, where)
        s._n, s._shape = B, (B, h, w)
        if h % 8 or w % 8:
            raise ValueError("Batch_hand crops must have sides that are multiples of 8 (srcmx/Batch_model.py:377)")
        _lib.check(_lib.lib().opb_batch_hand_submit_u8(s.handle, ptr, where, B, h, w))

    def collect(self, session=None):
        s = session or self._session
        peaks = np.empty((s._n, 21, 3), dtype=np.float64)
        _lib.check(_lib.lib().opb_hand_wait(s.handle, peaks.ctypes.data))
        return peaks

    def __call__(self, batch_imgs):
        out = []
        for lo in range(0, len(batch_imgs), self.MAX_BATCH):
            self.submit(batch_imgs[lo:lo + self.MAX_BATCH])
            out.append(self.collect())
        return np.concatenate(out, 0)

    def last_maps(self, session=None):
        """blurred heat maps (B,h,w,22) float32 of the last submitted chunk (Batch_model.py:378-385)."""
        s = session or self._session
        B, h, w = s._shape
        blurred = np.empty((B, 22, h, w), dtype=np.float32)
        _lib.check(_lib.lib().opb_batch_maps(s.handle, blurred.ctypes.data))
        return np.ascontiguousarray(blurred.transpose(0, 2, 3, 1))
