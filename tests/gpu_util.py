"""Helpers for the -m gpu parity tests: thin wrappers that call the C ABI with torch CUDA tensors as device
memory (torch is plumbing here: allocation + data_ptr)."""
import ctypes

import numpy as np
import torch

from pytorch_openpose_b200 import _lib


def ctx():
    return _lib.context(0)


def conv2d(x_nhwc_bf16, weight, bias, relu, pool=False, out_fp32=False, impl=0):
    """x: (n,h,w,cin) bf16 CUDA tensor; weight (cout,cin,k,k) fp32 CPU tensor.  Returns (n,h',w',cout_store)."""
    n, h, w, cin = x_nhwc_bf16.shape
    cout, _, k, _ = weight.shape
    cs = (cout + 7) // 8 * 8
    ho, wo = (h // 2, w // 2) if pool else (h, w)
    out = torch.full((n, ho, wo, cs), float("nan"), device="cuda", dtype=torch.float32 if out_fp32 else torch.bfloat16)
    wn = np.ascontiguousarray(weight.numpy(), dtype=np.float32)
    bn = np.ascontiguousarray(bias.numpy(), dtype=np.float32)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().opb_conv2d(ctx(), x_nhwc_bf16.data_ptr(), n, h, w, cin, wn.ctypes.data, bn.ctypes.data, cout, k,
                                     int(relu), int(pool), int(out_fp32), out.data_ptr(), impl))
    return out


def preprocess(img, scale):
    H, W = img.shape[:2]
    h, w, hp, wp = (ctypes.c_int() for _ in range(4))
    torch.cuda.synchronize()
    _lib.check(_lib.lib().opb_scale_dims(H, W, scale, ctypes.byref(h), ctypes.byref(w), ctypes.byref(hp), ctypes.byref(wp)))
    d_img = torch.from_numpy(np.ascontiguousarray(img)).cuda()
    out = torch.zeros((hp.value, wp.value, 3), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    _lib.check(_lib.lib().opb_preprocess(ctx(), d_img.data_ptr(), H, W, scale, out.data_ptr()))
    return out.cpu().numpy(), (h.value, w.value, hp.value, wp.value)


def net_forward(session, in_u8):
    """in_u8: (n,hp,wp,3) uint8 numpy.  Returns (paf (n,ho,wo,38) | None, heat (n,ho,wo,19|22)) fp32 numpy."""
    n, hp, wp, _ = in_u8.shape
    d_in = torch.from_numpy(np.ascontiguousarray(in_u8)).cuda()
    body = session.net.kind == _lib.NET_BODY
    paf = torch.zeros((n, hp // 8, wp // 8, 40), device="cuda") if body else None
    heat = torch.zeros((n, hp // 8, wp // 8, 24), device="cuda")
    torch.cuda.synchronize()
    _lib.check(_lib.lib().opb_net_forward(session.handle, d_in.data_ptr(), n, hp, wp,
                                          paf.data_ptr() if body else None, heat.data_ptr()))
    if body:
        return paf.cpu().numpy()[..., :38], heat.cpu().numpy()[..., :19]
    return None, heat.cpu().numpy()[..., :22]


def upsample_avg(maps_nhwc, scales, H, W):
    """maps: list of (ho,wo,C) fp32 numpy (C<=cstride handled by padding to a multiple of 8)."""
    C = maps_nhwc[0].shape[2]
    cs = (C + 7) // 8 * 8
    dev = []
    for m in maps_nhwc:
        p = np.zeros(m.shape[:2] + (cs,), np.float32)
        p[..., :C] = m
        dev.append(torch.from_numpy(p).cuda())
    ptrs = (ctypes.c_void_p * len(dev))(*[d.data_ptr() for d in dev])
    sc, ns = _lib.scales_array(scales)
    out = torch.zeros((C, H, W), device="cuda")
    torch.cuda.synchronize()
    _lib.check(_lib.lib().opb_upsample_avg(ctx(), ptrs, sc, ns, C, cs, H, W, out.data_ptr()))
    return out.cpu().numpy()


def find_peaks(heat_chw, thre1=0.1, capacity=16384):
    C, H, W = heat_chw.shape
    d = torch.from_numpy(np.ascontiguousarray(heat_chw, dtype=np.float32)).cuda()
    cand = torch.zeros((capacity, 4), dtype=torch.float64, device="cuda")
    pb = (ctypes.c_int * 19)()
    n = ctypes.c_int()
    torch.cuda.synchronize()
    _lib.check(_lib.lib().opb_find_peaks(ctx(), d.data_ptr(), H, W, thre1, cand.data_ptr(), capacity, pb, ctypes.byref(n)))
    return cand[:n.value].cpu().numpy(), list(pb), cand


def find_peaks_blurred(blurred_chw, thre1=0.1, capacity=16384):
    C, H, W = blurred_chw.shape
    d = torch.from_numpy(np.ascontiguousarray(blurred_chw, dtype=np.float32)).cuda()
    cand = torch.zeros((capacity, 4), dtype=torch.float64, device="cuda")
    pb = (ctypes.c_int * 19)()
    n = ctypes.c_int()
    torch.cuda.synchronize()
    _lib.check(_lib.lib().opb_find_peaks_blurred(ctx(), d.data_ptr(), H, W, thre1, cand.data_ptr(), capacity, pb,
                                                 ctypes.byref(n)))
    return cand[:n.value].cpu().numpy(), list(pb), cand


def smooth(heat_chw):
    C, H, W = heat_chw.shape
    d = torch.from_numpy(np.ascontiguousarray(heat_chw, dtype=np.float32)).cuda()
    out = torch.zeros((C, H, W), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    _lib.check(_lib.lib().opb_smooth_debug(ctx(), d.data_ptr(), C, H, W, out.data_ptr()))
    return out.cpu().numpy()


def group_limbs(paf_chw, cand_dev, part_begin, thre2=0.05, subset_cap=1024, conn_cap=512):
    C, H, W = paf_chw.shape
    d = torch.from_numpy(np.ascontiguousarray(paf_chw, dtype=np.float32)).cuda()
    pb = (ctypes.c_int * 19)(*part_begin)
    subset = np.zeros((subset_cap, 20))
    conns = np.zeros((19, conn_cap, 5))
    cc = (ctypes.c_int * 19)()
    ns = ctypes.c_int()
    torch.cuda.synchronize()
    _lib.check(_lib.lib().opb_group_limbs(ctx(), d.data_ptr(), H, W, cand_dev.data_ptr(), pb, thre2, subset.ctypes.data,
                                          subset_cap, ctypes.byref(ns), conns.ctypes.data, conn_cap, cc))
    return subset[:ns.value].copy(), conns, list(cc)


def hand_peaks(heat_chw, thre=0.03):
    C, H, W = heat_chw.shape
    d = torch.from_numpy(np.ascontiguousarray(heat_chw[:21], dtype=np.float32)).cuda()
    out = np.zeros((21, 3))
    torch.cuda.synchronize()
    _lib.check(_lib.lib().opb_hand_peaks(ctx(), d.data_ptr(), H, W, thre, out.ctypes.data))
    return out
