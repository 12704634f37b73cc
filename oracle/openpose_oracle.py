"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy / torch-CPU) of the reference's OpenPose
inference path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference arm
may import this module; the product (`pytorch_openpose_b200`) never does.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md §4,
§8c).  This restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container by
`oracle/make_golden.py` (which imports /root/reference through `oracle/reference_loader.py`) and
committed as `tests/golden/*.npz`; `tests/test_oracle_golden.py` re-checks them on every CPU run, and
`tests/test_oracle_vs_reference.py` re-runs the live reference whenever /root/reference is present.

Third-party numerics the reference delegates to (un-vendored, unpinned in requirements.txt:1-6) and how
they are restated here:
  * cv2.resize(INTER_CUBIC) on uint8 (src/body.py:38, src/hand.py:38) -> `resize_cubic_u8`: OpenCV's
    open-source algorithm (A=-0.75, half-pixel centres, clamped taps, 11-bit fixed-point coefficients,
    float vertical pass S3..S0 for the SIMD body and integer pass for the <8-element row tail).
    Bit-exact with cv2 when cv2.setUseOptimized(False); the default build routes through closed-source
    Intel IPP, which differs by at most 1 LSB on ~5-10 % of pixels (checked in tests, documented in
    DESIGN.md).
  * cv2.resize(INTER_CUBIC) on float32 (src/body.py:55,57) -> `resize_cubic_f32` (<=1e-6 of cv2).
  * scipy.ndimage.gaussian_filter(sigma=3) (src/body.py:75, src/hand.py:62) -> `gaussian_sigma3`,
    bit-exact (same symmetric-kernel accumulation order as scipy's NI_Correlate1D).
  * skimage.measure.label(connectivity=2) (src/hand.py:68) -> scipy.ndimage.label with a full 3x3
    structure (both number components in raster order of their first pixel).
  * torch.conv2d / max_pool2d / relu (src/model.py) -> torch CPU fp32 functional calls.
"""
import math
import numpy as np

# --------------------------------------------------------------------------------------------
# constants (src/body.py:25-31,97-107 ; src/hand.py:26-31)
# --------------------------------------------------------------------------------------------
BOXSIZE = 368
STRIDE = 8
PAD_VALUE = 128
BODY_THRE1 = 0.1
BODY_THRE2 = 0.05
HAND_THRE = 0.03
MID_NUM = 10
# 1-based part ids per limb (src/body.py:97-99)
LIMB_SEQ = ((2, 3), (2, 6), (3, 4), (4, 5), (6, 7), (7, 8), (2, 9), (9, 10), (10, 11), (2, 12),
            (12, 13), (13, 14), (2, 1), (1, 15), (15, 17), (1, 16), (16, 18), (3, 17), (6, 18))
# PAF channel pair per limb = mapIdx[k]-19 (src/body.py:101-103,110)
PAF_CH = ((12, 13), (20, 21), (14, 15), (16, 17), (22, 23), (24, 25), (0, 1), (2, 3), (4, 5), (6, 7),
          (8, 9), (10, 11), (28, 29), (30, 31), (34, 35), (32, 33), (36, 37), (18, 19), (26, 27))

# --------------------------------------------------------------------------------------------
# network definition (src/model.py:25-214) as a flat layer table
# --------------------------------------------------------------------------------------------
_VGG_HEAD = [("conv1_1", 3, 64), ("conv1_2", 64, 64), "pool",
             ("conv2_1", 64, 128), ("conv2_2", 128, 128), "pool",
             ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv3_4", 256, 256), "pool",
             ("conv4_1", 256, 512), ("conv4_2", 512, 512)]


def body_layers():
    """[(block, [(name, cin, cout, k, relu) | 'pool'])] in the reference's module-creation order
    (src/model.py:35-104: model0, then block1_1, block1_2, block2_1, block2_2, ...)."""
    # src/model.py:30-33 lists 'Mconv7_stage6_L1' twice and omits 'Mconv7_stage6_L2' -> the final
    # heat-map conv IS followed by a ReLU.
    no_relu = {"conv5_5_CPM_L1", "conv5_5_CPM_L2"} | {"Mconv7_stage%d_L%d" % (s, b) for s in range(2, 7)
                                                      for b in (1, 2)} - {"Mconv7_stage6_L2"}
    head = [(n[0], n[1], n[2], 3) if n != "pool" else "pool" for n in _VGG_HEAD]
    head += [("conv4_3_CPM", 512, 256, 3), ("conv4_4_CPM", 256, 128, 3)]
    blocks = [("model0", head)]
    for b, cout in ((1, 38), (2, 19)):
        blocks.append(("model1_%d" % b, [("conv5_%d_CPM_L%d" % (i, b), 128, 128, 3) for i in (1, 2, 3)]
                       + [("conv5_4_CPM_L%d" % b, 128, 512, 1), ("conv5_5_CPM_L%d" % b, 512, cout, 1)]))
    for s in range(2, 7):
        for b, cout in ((1, 38), (2, 19)):
            blocks.append(("model%d_%d" % (s, b),
                           [("Mconv1_stage%d_L%d" % (s, b), 185, 128, 7)]
                           + [("Mconv%d_stage%d_L%d" % (i, s, b), 128, 128, 7) for i in (2, 3, 4, 5)]
                           + [("Mconv6_stage%d_L%d" % (s, b), 128, 128, 1),
                              ("Mconv7_stage%d_L%d" % (s, b), 128, cout, 1)]))
    return [(blk, [l if l == "pool" else l + (l[0] not in no_relu,) for l in layers]) for blk, layers in blocks]


def hand_layers():
    """src/model.py:144-195 (model1_0, model1_1, model2..model6)."""
    no_relu = {"conv6_2_CPM"} | {"Mconv7_stage%d" % s for s in range(2, 7)}
    head = [(n[0], n[1], n[2], 3) if n != "pool" else "pool" for n in _VGG_HEAD]
    head += [("conv4_3", 512, 512, 3), ("conv4_4", 512, 512, 3), ("conv5_1", 512, 512, 3),
             ("conv5_2", 512, 512, 3), ("conv5_3_CPM", 512, 128, 3)]
    blocks = [("model1_0", head), ("model1_1", [("conv6_1_CPM", 128, 512, 1), ("conv6_2_CPM", 512, 22, 1)])]
    for s in range(2, 7):
        blocks.append(("model%d" % s, [("Mconv1_stage%d" % s, 150, 128, 7)]
                       + [("Mconv%d_stage%d" % (i, s), 128, 128, 7) for i in (2, 3, 4, 5)]
                       + [("Mconv6_stage%d" % s, 128, 128, 1), ("Mconv7_stage%d" % s, 128, 22, 1)]))
    return [(blk, [l if l == "pool" else l + (l[0] not in no_relu,) for l in layers]) for blk, layers in blocks]


def make_weights(kind, seed=0, init="torch_default"):
    """Random-init weights in the caffe-keyed flat format the reference checkpoints use
    (`<layer>.weight` / `<layer>.bias`, src/util.py:36-40).

    init='torch_default': the very tensors `torch.manual_seed(seed); bodypose_model()` would hold --
        nn.Conv2d modules are created in the reference's order so the RNG stream matches
        (verified against the live reference in tests/test_oracle_vs_reference.py).
    init='kaiming': Kaiming-normal(ReLU) weights, zero bias -> maps with real spatial structure
        (SURVEY.md §7 'random-init weights give almost constant maps').
    """
    import torch
    layers = body_layers() if kind == "body" else hand_layers()
    sd = {}
    torch.manual_seed(seed)
    for _, block in layers:
        for l in block:
            if l == "pool":
                continue
            name, cin, cout, k, _ = l
            if init == "torch_default":
                conv = torch.nn.Conv2d(cin, cout, k, 1, k // 2)
                w, b = conv.weight.detach(), conv.bias.detach()
            elif init == "kaiming":
                w = torch.randn(cout, cin, k, k) * math.sqrt(2.0 / (cin * k * k))
                b = torch.zeros(cout)
            else:
                raise ValueError(init)
            sd[name + ".weight"] = w.contiguous()
            sd[name + ".bias"] = b.contiguous()
    return sd


def _round_bf16(t):
    import torch
    return t.to(torch.bfloat16).to(torch.float32)


def _run_block(x, block, sd, bf16):
    import torch
    import torch.nn.functional as F
    for l in block:
        if l == "pool":
            x = F.max_pool2d(x, 2, 2, 0)
            continue
        name, _, _, k, relu = l
        w, b = sd[name + ".weight"].float(), sd[name + ".bias"].float()
        if bf16:
            w = _round_bf16(w)
        x = F.conv2d(x, w, b, 1, k // 2)
        if relu:
            x = torch.relu(x)
        if bf16 and not name.startswith("Mconv7_stage6"):      # final maps stay fp32
            x = _round_bf16(x)
    return x


def body_net(x, sd, bf16=False):
    """bodypose_model.forward (src/model.py:106-133).  x: (N,3,Hp,Wp) float32 torch CPU tensor.
    Returns (paf (N,38,h,w), heat (N,19,h,w)).  bf16=True emulates the device numerics (weights and
    every stored activation rounded to bf16, fp32 accumulate, final stage-6 outputs kept fp32)."""
    import torch
    blocks = dict(body_layers())
    with torch.no_grad():
        if bf16:
            x = _round_bf16(x)
        feat = _run_block(x, blocks["model0"], sd, bf16)
        l1 = _run_block(feat, blocks["model1_1"], sd, bf16)
        l2 = _run_block(feat, blocks["model1_2"], sd, bf16)
        for s in range(2, 7):
            cat = torch.cat([l1, l2, feat], 1)
            l1 = _run_block(cat, blocks["model%d_1" % s], sd, bf16)
            l2 = _run_block(cat, blocks["model%d_2" % s], sd, bf16)
    return l1, l2


def hand_net(x, sd, bf16=False):
    """handpose_model.forward (src/model.py:197-214).  Returns (N,22,h,w)."""
    import torch
    blocks = dict(hand_layers())
    with torch.no_grad():
        if bf16:
            x = _round_bf16(x)
        feat = _run_block(x, blocks["model1_0"], sd, bf16)
        out = _run_block(feat, blocks["model1_1"], sd, bf16)
        for s in range(2, 7):
            out = _run_block(torch.cat([out, feat], 1), blocks["model%d" % s], sd, bf16)
    return out


# --------------------------------------------------------------------------------------------
# cv2.resize(INTER_CUBIC) restatements
# --------------------------------------------------------------------------------------------
def _cubic_coeffs_f32(frac):
    """OpenCV interpolateCubic, A=-0.75, evaluated in float32 (frac: float32 array)."""
    x = frac.astype(np.float32)
    A = np.float32(-0.75)
    one = np.float32(1)
    c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    c2 = ((A + np.float32(2)) * (one - x) - (A + np.float32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], -1).astype(np.float32)


def cubic_taps(src, dst, scale):
    """Per destination index: first tap index (may be <0 / >src-4: taps are clamped when used) and the
    4 float32 coefficients.  `scale` is the double src-per-dst sampling step cv2 uses."""
    d = np.arange(dst, dtype=np.float64)
    fx = ((d + 0.5) * scale - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int64)
    frac = (fx - sx.astype(np.float32)).astype(np.float32)
    return sx - 1, _cubic_coeffs_f32(frac)


def resize_dsize(n, f):
    """cv2.resize(fx=f): dsize = saturate_cast<int>(n*f) = round-half-even."""
    return int(np.rint(n * f))


def resize_cubic_u8(img, f):
    """cv2.resize(img, (0,0), fx=f, fy=f, INTER_CUBIC) for uint8 HWC (src/body.py:38)."""
    H, W, C = img.shape
    dw, dh = resize_dsize(W, f), resize_dsize(H, f)
    x0, xa = cubic_taps(W, dw, 1.0 / f)
    y0, ya = cubic_taps(H, dh, 1.0 / f)
    ixa = np.clip(np.rint(xa * np.float32(2048)), -32768, 32767).astype(np.int64)
    iya = np.clip(np.rint(ya * np.float32(2048)), -32768, 32767).astype(np.int64)
    xi = np.clip(x0[:, None] + np.arange(4)[None, :], 0, W - 1)
    yi = np.clip(y0[:, None] + np.arange(4)[None, :], 0, H - 1)
    hor = (img.astype(np.int64)[:, xi, :] * ixa[None, :, :, None]).sum(2)          # (H,dw,C) int32 range
    rows = hor[yi].reshape(dh, 4, dw * C)                                         # (dh,4,dw*C)
    # vertical pass, SIMD body: float32, S3*b3 first then S2,S1,S0 added (mul then add, no FMA)
    b = (iya.astype(np.float32) * np.float32(1.0 / (2048 * 2048)))[:, :, None]
    r = rows.astype(np.float32)
    acc = (r[:, 3] * b[:, 3]).astype(np.float32)
    for k in (2, 1, 0):
        acc = (acc + (r[:, k] * b[:, k]).astype(np.float32)).astype(np.float32)
    out = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    # row tail (< 8 elements): integer pass with 22-bit rounding shift
    n = dw * C
    nv = (n // 8) * 8
    if nv < n:
        t = ((rows[:, :, nv:] * iya[:, :, None]).sum(1) + (1 << 21)) >> 22
        out[:, nv:] = np.clip(t, 0, 255).astype(np.uint8)
    return out.reshape(dh, dw, C)


def resize_cubic_f32(img, dsize=None, f=None):
    """cv2.resize for float32 HWC: either (0,0),fx=fy=f (src/body.py:55) or explicit dsize=(w,h)
    (src/body.py:57), in which case cv2's step is 1/(dst/src) in double."""
    H, W, C = img.shape
    if dsize is None:
        dw, dh = resize_dsize(W, f), resize_dsize(H, f)
        sx = sy = 1.0 / f
    else:
        dw, dh = dsize
        sx, sy = 1.0 / (dw / W), 1.0 / (dh / H)
    x0, xa = cubic_taps(W, dw, sx)
    y0, ya = cubic_taps(H, dh, sy)
    xi = np.clip(x0[:, None] + np.arange(4)[None, :], 0, W - 1)
    yi = np.clip(y0[:, None] + np.arange(4)[None, :], 0, H - 1)
    s = img.astype(np.float32)[:, xi, :]
    hor = s[:, :, 0] * xa[None, :, 0, None]
    for k in (1, 2, 3):
        hor = hor + s[:, :, k] * xa[None, :, k, None]
    r = hor.astype(np.float32)[yi]
    out = r[:, 0] * ya[:, 0, None, None]
    for k in (1, 2, 3):
        out = out + r[:, k] * ya[:, k, None, None]
    return out.astype(np.float32)


def cubic_matrix(src, dst, scale):
    """Dense (dst,src) float64 matrix of one cubic resize pass (clamped taps accumulate)."""
    t0, co = cubic_taps(src, dst, scale)
    M = np.zeros((dst, src))
    for k in range(4):
        np.add.at(M, (np.arange(dst), np.clip(t0 + k, 0, src - 1)), co[:, k].astype(np.float64))
    return M


def composite_upsample_matrix(n_out_net, n_resized, n_orig):
    """1-D operator of src/body.py:55-57 along one axis: x8 cubic upsample of the net output
    (n_out_net -> 8*n_out_net), crop to the un-padded resized length n_resized, cubic resize to n_orig."""
    up = cubic_matrix(n_out_net, n_out_net * STRIDE, 1.0 / STRIDE)[:n_resized]
    down = cubic_matrix(n_resized, n_orig, 1.0 / (n_orig / n_resized))
    return down @ up


# --------------------------------------------------------------------------------------------
# preprocessing (src/body.py:32,38-41 ; src/util.py:12-32)
# --------------------------------------------------------------------------------------------
def scale_plan(H, W, scale_search):
    """Per scale: multiplier, resized (h,w), padded (hp,wp), net output (ho,wo)."""
    plan = []
    for s in scale_search:
        m = s * BOXSIZE / H
        h, w = resize_dsize(H, m), resize_dsize(W, m)
        hp, wp = -(-h // STRIDE) * STRIDE, -(-w // STRIDE) * STRIDE
        plan.append(dict(mult=m, h=h, w=w, hp=hp, wp=wp, ho=hp // STRIDE, wo=wp // STRIDE))
    return plan


def pad_right_down(img, stride=STRIDE, value=PAD_VALUE):
    """util.padRightDownCorner (src/util.py:12-32): pad bottom/right to a stride multiple."""
    h, w = img.shape[:2]
    ph, pw = (-h) % stride, (-w) % stride
    out = np.full((h + ph, w + pw, img.shape[2]), value, dtype=img.dtype)
    out[:h, :w] = img
    return out, [0, 0, ph, pw]


def preprocess(img, mult, use_cv2=False):
    """-> (uint8 padded HWC image, float32 NCHW tensor in [-0.5,0.5), pad)  (src/body.py:38-41)."""
    if use_cv2:
        import cv2
        resized = cv2.resize(img, (0, 0), fx=mult, fy=mult, interpolation=cv2.INTER_CUBIC)
    else:
        resized = resize_cubic_u8(img, mult)
    padded, pad = pad_right_down(resized)
    x = np.ascontiguousarray(np.transpose(np.float32(padded), (2, 0, 1))[None] / 256 - 0.5)
    return padded, x.astype(np.float32), pad


# --------------------------------------------------------------------------------------------
# upsample + cross-scale average (src/body.py:54-68 ; src/hand.py:52-57)
# --------------------------------------------------------------------------------------------
def upsample_avg(maps, plan, H, W, use_cv2=True):
    """maps: list (per scale) of (C,ho,wo) float32 net outputs.  Returns (H,W,C) float64 average, computed
    the reference's way: x8 cubic, crop padding, cubic resize to (W,H), += map/len(scales)."""
    C = maps[0].shape[0]
    avg = np.zeros((H, W, C))
    for m, p in zip(maps, plan):
        hwc = np.ascontiguousarray(np.transpose(m, (1, 2, 0)).astype(np.float32))
        if use_cv2:
            import cv2
            parts = []
            for c0 in range(0, C, 32):     # cv2.resize handles <=512 channels, keep chunks small
                chunk = np.ascontiguousarray(hwc[:, :, c0:c0 + 32])
                up = cv2.resize(chunk, (0, 0), fx=STRIDE, fy=STRIDE, interpolation=cv2.INTER_CUBIC)
                up = up.reshape(up.shape[0], up.shape[1], -1)[:p["h"], :p["w"], :]
                full = cv2.resize(up, (W, H), interpolation=cv2.INTER_CUBIC)
                parts.append(full.reshape(H, W, -1))
            full = np.concatenate(parts, 2)
        else:
            up = resize_cubic_f32(hwc, f=STRIDE)[:p["h"], :p["w"], :]
            full = resize_cubic_f32(up, dsize=(W, H))
        avg += full / len(plan)
    return avg


# --------------------------------------------------------------------------------------------
# Gaussian sigma=3 (scipy.ndimage.gaussian_filter, src/body.py:75, src/hand.py:62) -- bit-exact
# --------------------------------------------------------------------------------------------
GAUSS_RADIUS = 12     # int(4.0*3 + 0.5)


def gaussian_weights():
    x = np.arange(-GAUSS_RADIUS, GAUSS_RADIUS + 1)
    p = np.exp(-0.5 / 9.0 * x ** 2)
    return p / p.sum()


def _reflect_index(idx, n):
    # scipy 'reflect': d c b a | a b c d | d c b a (period 2n)
    idx = np.mod(idx, 2 * n)
    return np.where(idx >= n, 2 * n - 1 - idx, idx)


def _gauss1d(a, axis):
    w = gaussian_weights()
    r = GAUSS_RADIUS
    a = np.moveaxis(a, axis, -1)
    n = a.shape[-1]
    p = a[..., _reflect_index(np.arange(-r, n + r), n)]
    out = p[..., r:r + n] * w[r]
    for j in range(-r, 0):     # far taps first, symmetric pair summed before the multiply
        out = out + (p[..., r + j:r + j + n] + p[..., r - j:r - j + n]) * w[r + j]
    return np.moveaxis(out, -1, axis)


def gaussian_sigma3(m):
    """float64 (H,W) -> float64 (H,W); axis 0 first, then axis 1 (scipy order)."""
    return _gauss1d(_gauss1d(np.asarray(m, dtype=np.float64), 0), 1)


# --------------------------------------------------------------------------------------------
# body post-processing (src/body.py:70-212)
# --------------------------------------------------------------------------------------------
def find_peaks(heatmap_avg, thre1=BODY_THRE1, smooth=gaussian_sigma3):
    """src/body.py:70-94.  heatmap_avg (H,W,>=18) float64.  Returns list of 18 arrays (n,4)
    [x, y, raw score, id]; ids are cumulative over parts; order within a part is (y, x)."""
    peaks, counter = [], 0
    for part in range(18):
        raw = heatmap_avg[:, :, part]
        sm = smooth(raw)
        z = np.zeros_like(sm)
        nb = [z.copy() for _ in range(4)]
        nb[0][1:, :] = sm[:-1, :]
        nb[1][:-1, :] = sm[1:, :]
        nb[2][:, 1:] = sm[:, :-1]
        nb[3][:, :-1] = sm[:, 1:]
        keep = (sm >= nb[0]) & (sm >= nb[1]) & (sm >= nb[2]) & (sm >= nb[3]) & (sm > thre1)
        ys, xs = np.nonzero(keep)
        arr = np.zeros((len(xs), 4))
        arr[:, 0], arr[:, 1], arr[:, 2] = xs, ys, raw[ys, xs]
        arr[:, 3] = counter + np.arange(len(xs))
        counter += len(xs)
        peaks.append(arr)
    return peaks


def score_pairs(paf_avg, candA, candB, k, img_h, thre2=BODY_THRE2):
    """src/body.py:118-141 for limb k, vectorised over all (i,j) pairs but with the same float64
    operation order per pair.  Returns list of (i, j, score) that pass both criteria, in (i,j) order."""
    nA, nB = len(candA), len(candB)
    ax, ay = candA[:, 0][:, None], candA[:, 1][:, None]
    bx, by = candB[:, 0][None, :], candB[:, 1][None, :]
    vx, vy = (bx - ax) * np.ones((nA, nB)), (by - ay) * np.ones((nA, nB))
    norm = np.sqrt(vx * vx + vy * vy) + 1e-10
    ux, uy = vx / norm, vy / norm
    # np.linspace(a, b, 10): arange(10)*step + a with step=(b-a)/9, last sample forced to b
    t = np.arange(MID_NUM, dtype=np.float64)
    stepx, stepy = vx / (MID_NUM - 1), vy / (MID_NUM - 1)
    px = t[None, None, :] * stepx[:, :, None] + ax[:, :, None]
    py = t[None, None, :] * stepy[:, :, None] + ay[:, :, None]
    px[:, :, -1] = bx * np.ones((nA, nB))
    py[:, :, -1] = by * np.ones((nA, nB))
    xi, yi = np.rint(px).astype(np.int64), np.rint(py).astype(np.int64)
    cx, cy = PAF_CH[k]
    dots = paf_avg[yi, xi, cx] * ux[:, :, None] + paf_avg[yi, xi, cy] * uy[:, :, None]
    total = np.zeros((nA, nB))
    for s in range(MID_NUM):          # builtin sum(): sequential, starting from 0
        total = total + dots[:, :, s]
    prior = np.minimum(0.5 * img_h / norm - 1, 0)
    score = total / MID_NUM + prior
    ok = ((dots > thre2).sum(-1) > 0.8 * MID_NUM) & (score > 0)
    ii, jj = np.nonzero(ok)
    return [(int(i), int(j), float(score[i, j])) for i, j in zip(ii, jj)]


def match_limbs(peaks, paf_avg, img_h, thre2=BODY_THRE2):
    """src/body.py:105-155.  Returns (connection_all: list of 19 arrays (n,5) [idA,idB,score,i,j] or
    None for limbs with an empty side, special_k)."""
    connection_all, special = [], []
    for k, (pa, pb) in enumerate(LIMB_SEQ):
        candA, candB = peaks[pa - 1], peaks[pb - 1]
        if len(candA) == 0 or len(candB) == 0:
            special.append(k)
            connection_all.append(None)
            continue
        cands = score_pairs(paf_avg, candA, candB, k, img_h, thre2)
        cands.sort(key=lambda c: -c[2])        # stable, descending (== sorted(..., reverse=True))
        usedA, usedB, rows = set(), set(), []
        limit = min(len(candA), len(candB))
        for i, j, s in cands:
            if i in usedA or j in usedB:
                continue
            usedA.add(i)
            usedB.add(j)
            rows.append([candA[i, 3], candB[j, 3], s, i, j])
            if len(rows) >= limit:
                break
        connection_all.append(np.array(rows, dtype=np.float64).reshape(-1, 5))
    return connection_all, special


def assemble(peaks, connection_all, special):
    """src/body.py:157-212.  Returns (candidate (N,4) or shape (0,), subset (P,20))."""
    flat = [row for arr in peaks for row in arr]
    candidate = np.array(flat)
    rows = []          # list of 20-vectors (float64)
    for k, (pa, pb) in enumerate(LIMB_SEQ):
        if k in special:
            continue
        ia, ib = pa - 1, pb - 1
        for conn in connection_all[k]:
            a_id, b_id, limb_score = conn[0], conn[1], conn[2]
            hits = [j for j, r in enumerate(rows) if r[ia] == a_id or r[ib] == b_id]
            if len(hits) > 2:
                raise IndexError("list assignment index out of range")     # src/body.py:173
            if len(hits) == 1 or (len(hits) == 2 and any(
                    rows[hits[0]][c] >= 0 and rows[hits[1]][c] >= 0 for c in range(18))):
                r = rows[hits[0]]
                if len(hits) == 2 or r[ib] != b_id:
                    r[ib] = b_id
                    r[19] += 1
                    r[18] += candidate[int(b_id), 2] + limb_score
            elif len(hits) == 2:
                r1, r2 = rows[hits[0]], rows[hits[1]]
                r1[:18] += r2[:18] + 1
                r1[18:] += r2[18:]
                r1[18] += limb_score
                del rows[hits[1]]
            elif k < 17:
                r = -np.ones(20)
                r[ia], r[ib] = a_id, b_id
                r[19] = 2
                r[18] = (0 + candidate[int(a_id), 2] + candidate[int(b_id), 2]) + limb_score
                rows.append(r)
    kept = [r for r in rows if not (r[19] < 4 or r[18] / r[19] < 0.4)]
    subset = np.array(kept).reshape(-1, 20) if kept else -np.ones((0, 20))
    return candidate, subset


def body_postprocess(heatmap_avg, paf_avg, img_h, thre1=BODY_THRE1, thre2=BODY_THRE2):
    peaks = find_peaks(heatmap_avg, thre1)
    conns, special = match_limbs(peaks, paf_avg, img_h, thre2)
    return assemble(peaks, conns, special)


def body_call(img, sd, scale_search=(0.5,), use_cv2=True, bf16=False, return_maps=False):
    """Body.__call__ (src/body.py:24-212) on a uint8 BGR HWC image with weights `sd`."""
    import torch
    H, W = img.shape[:2]
    plan = scale_plan(H, W, scale_search)
    pafs, heats = [], []
    for p in plan:
        _, x, _ = preprocess(img, p["mult"], use_cv2)
        paf, heat = body_net(torch.from_numpy(x), sd, bf16)
        pafs.append(paf[0].numpy())
        heats.append(heat[0].numpy())
    heat_avg = upsample_avg(heats, plan, H, W, use_cv2)
    paf_avg = upsample_avg(pafs, plan, H, W, use_cv2)
    out = body_postprocess(heat_avg, paf_avg, H)
    return out + (heat_avg, paf_avg) if return_maps else out


# --------------------------------------------------------------------------------------------
# hand post-processing (src/hand.py:59-75 ; src/util.py:205-210)
# --------------------------------------------------------------------------------------------
def hand_postprocess(heatmap_avg, thre=HAND_THRE):
    """heatmap_avg (h,w,>=21) float64 -> (21,3) float64 [x, y, score] (zeros when nothing is above thre).
    Does not modify its input (the reference zeroes its local heatmap_avg in place)."""
    from scipy import ndimage as ndi
    eight = np.ones((3, 3), dtype=int)
    out = np.zeros((21, 3))
    for part in range(21):
        raw = heatmap_avg[:, :, part]
        binary = gaussian_sigma3(raw) > thre
        if not binary.any():
            continue
        lab, n = ndi.label(binary, structure=eight)
        sums = [np.sum(raw[lab == i]) for i in range(1, n + 1)]
        best = int(np.argmax(sums)) + 1
        kept = np.where(lab == best, raw, 0.0)
        flat = int(np.argmax(kept))                  # first maximum in row-major order == util.npmax
        y, x = divmod(flat, kept.shape[1])
        out[part] = (x, y, kept[y, x])
    return out


def hand_call(img, sd, scale_search=(0.5, 1.0, 1.5, 2.0), use_cv2=True, bf16=False, return_maps=False):
    """Hand.__call__ (src/hand.py:25-75)."""
    import torch
    H, W = img.shape[:2]
    plan = scale_plan(H, W, scale_search)
    maps = []
    for p in plan:
        _, x, _ = preprocess(img, p["mult"], use_cv2)
        maps.append(hand_net(torch.from_numpy(x), sd, bf16)[0].numpy())
    avg = upsample_avg(maps, plan, H, W, use_cv2)
    peaks = hand_postprocess(avg)
    return (peaks, avg) if return_maps else peaks


# --------------------------------------------------------------------------------------------
# batched estimators: Batch_body / Batch_hand, srcmx/Batch_model.py:107-406
# A second, numerically distinct contract: float frames in [0,1], torch bicubic resizes, one scale,
# 5x5 blur (utilmx.py:243-263), peaks found AND scored on the blurred map (utilmx.py:230-241).
# --------------------------------------------------------------------------------------------
BATCH_BODY_SCALE = 0.5        # Batch_model.py:118
BATCH_HAND_THRE = 0.035       # Batch_model.py:361
BLUR5 = np.array([[0.00078633, 0.00655965, 0.01330373, 0.00655965, 0.00078633],       # utilmx.py:248-252
                  [0.00655965, 0.05472157, 0.11098164, 0.05472157, 0.00655965],
                  [0.01330373, 0.11098164, 0.22508352, 0.11098164, 0.01330373],
                  [0.00655965, 0.05472157, 0.11098164, 0.05472157, 0.00655965],
                  [0.00078633, 0.00655965, 0.01330373, 0.00655965, 0.00078633]], dtype=np.float32)


def batch_size_pad(g_scale, height, width):
    """Batch_body.calculate_size_pad (Batch_model.py:340-345) -> (scale, h, w, pad_h, pad_w)."""
    scale = BOXSIZE * g_scale / height
    h, w = int(height * scale), int(width * scale)
    return scale, h, w, (STRIDE - (h % STRIDE)) % STRIDE, (STRIDE - (w % STRIDE)) % STRIDE


def blur5(x):
    """utilmx.GaussianBlurConv.__call__ (utilmx.py:261-263): depthwise 5x5 on a reflect-padded (no edge repeat)
    float32 NCHW tensor."""
    import torch
    import torch.nn.functional as F
    c = x.shape[1]
    w = torch.from_numpy(BLUR5)[None, None].expand(c, 1, 5, 5).contiguous()
    return F.conv2d(F.pad(x, (2, 2, 2, 2), mode="reflect"), w, groups=c)


def blur5_fixed_order(maps):
    """The same 5x5 blur in a FIXED float32 operation order (taps in row-major order, product and sum rounded
    separately): what the device kernel computes bit for bit.  torch's depthwise conv adds the same 25 products in an
    unspecified order, so it agrees with this to a few float32 ulps only.  maps (H,W,C) -> (H,W,C) float32."""
    m = np.asarray(maps, dtype=np.float32)
    p = np.pad(m, ((2, 2), (2, 2), (0, 0)), mode="reflect")          # reflect without edge repeat == torch 'reflect'
    H, W = m.shape[:2]
    acc = np.zeros_like(m)
    for dy in range(5):
        for dx in range(5):
            acc = acc + BLUR5[dy, dx] * p[dy:dy + H, dx:dx + W]
    return acc


def batch_find_peaks(blurred, thre1=BODY_THRE1):
    """utilmx.findpeaks_torch (utilmx.py:230-241) + the id/score bookkeeping of Batch_model.py:185-194 for ONE frame.
    blurred (H,W,>=18) float32.  The threshold comparison happens in float32.  Returns 18 arrays (n,4)."""
    peaks, counter = [], 0
    t = np.float32(thre1)
    for part in range(18):
        sm = np.asarray(blurred[:, :, part], dtype=np.float32)
        z = np.zeros_like(sm)
        nb = [z.copy() for _ in range(4)]
        nb[0][1:, :] = sm[:-1, :]
        nb[1][:-1, :] = sm[1:, :]
        nb[2][:, 1:] = sm[:, :-1]
        nb[3][:, :-1] = sm[:, 1:]
        keep = (sm > t) & (sm >= nb[2]) & (sm >= nb[3]) & (sm >= nb[0]) & (sm >= nb[1])
        ys, xs = np.nonzero(keep)
        arr = np.zeros((len(xs), 4))
        arr[:, 0], arr[:, 1], arr[:, 2] = xs, ys, sm[ys, xs]
        arr[:, 3] = counter + np.arange(len(xs))
        counter += len(xs)
        peaks.append(arr)
    return peaks


def batch_body_postprocess(blurred, paf, thre1=BODY_THRE1, thre2=BODY_THRE2):
    """Peaks + Batch_body.FindBody_frame (Batch_model.py:206-338, the grouping of src/body.py:96-212 with
    heatmap.shape[0] as the image height) on one frame's maps: blurred (H,W,19) float32, paf (H,W,38) float32."""
    peaks = batch_find_peaks(blurred, thre1)
    conns, special = match_limbs(peaks, np.asarray(paf, dtype=np.float64), blurred.shape[0], thre2)
    return assemble(peaks, conns, special)


def batch_body_maps(batch, sd, bf16=False):
    """Batch_body.__call__ up to the blurred maps (Batch_model.py:142-176).  batch: (B,3,h,w) float32 in [0,1].
    Returns (blurred heat (B,h,w,19), paf (B,h,w,38)) float32."""
    import torch
    import torch.nn.functional as F
    x = torch.as_tensor(batch, dtype=torch.float32)
    _, _, h, w = x.shape
    scale, n_h, n_w, pad_h, pad_w = batch_size_pad(BATCH_BODY_SCALE, h, w)
    with torch.no_grad():
        x = F.interpolate(x, scale_factor=scale, mode="bicubic")
        x = F.pad(x - 0.5, [0, pad_w, 0, pad_h], mode="constant", value=0)
        paf, heat = body_net(x, sd, bf16)
        outs = []
        for m in (heat, paf):
            m = F.interpolate(torch.as_tensor(m), scale_factor=STRIDE, mode="bicubic")[:, :, :n_h, :n_w]
            outs.append(F.interpolate(m, size=(h, w), mode="bicubic"))
        heat = blur5(outs[0])
    return (np.ascontiguousarray(heat.numpy().transpose(0, 2, 3, 1)),
            np.ascontiguousarray(outs[1].numpy().transpose(0, 2, 3, 1)))


def batch_body_call(batch, sd, bf16=False):
    """Batch_body.__call__ (Batch_model.py:142-204) -> list of (candidate, subset) per frame."""
    heat, paf = batch_body_maps(batch, sd, bf16)
    return [batch_body_postprocess(heat[b], paf[b]) for b in range(len(heat))]


def batch_hand_postprocess(blurred, thre=BATCH_HAND_THRE):
    """Batch_model.py:388-405 for one crop: blurred (h,w,>=21) float32 -> (21,3) float64.  Threshold, component
    sums (numpy float32 summation) and the maximum all use the blurred map."""
    from scipy import ndimage as ndi
    eight = np.ones((3, 3), dtype=int)
    out = np.zeros((21, 3))
    for part in range(21):
        m = np.asarray(blurred[:, :, part], dtype=np.float32)
        binary = m > np.float32(thre)
        if not binary.any():
            continue
        lab, n = ndi.label(binary, structure=eight)
        best = int(np.argmax([np.sum(m[lab == i]) for i in range(1, n + 1)])) + 1
        kept = np.where(lab == best, m, np.float32(0))
        y, x = divmod(int(np.argmax(kept)), kept.shape[1])
        out[part] = (x, y, kept[y, x])
    return out


def batch_hand_maps(batch, sd, bf16=False):
    """Batch_hand.__call__ up to the blurred maps (Batch_model.py:366-386): (B,3,S,S) float32 -> (B,S,S,22)."""
    import torch
    import torch.nn.functional as F
    with torch.no_grad():
        heat = hand_net(torch.as_tensor(batch, dtype=torch.float32) - 0.5, sd, bf16)
        heat = blur5(F.interpolate(torch.as_tensor(heat), scale_factor=STRIDE, mode="bicubic"))
    return np.ascontiguousarray(heat.numpy().transpose(0, 2, 3, 1))


def batch_hand_call(batch, sd, bf16=False):
    """Batch_hand.__call__ (Batch_model.py:366-406) -> (B,21,3) float64."""
    heat = batch_hand_maps(batch, sd, bf16)
    return np.array([batch_hand_postprocess(heat[b]) for b in range(len(heat))])


# --------------------------------------------------------------------------------------------
# util.handDetect (src/util.py:133-201)
# --------------------------------------------------------------------------------------------
def hand_detect(candidate, subset, img_h, img_w):
    out = []
    for person in subset.astype(int):
        for (sh, el, wr), is_left in (((5, 6, 7), True), ((2, 3, 4), False)):
            if person[sh] == -1 or person[el] == -1 or person[wr] == -1:
                continue
            x1, y1 = candidate[person[sh]][:2]
            x2, y2 = candidate[person[el]][:2]
            x3, y3 = candidate[person[wr]][:2]
            x = x3 + 0.33 * (x3 - x2)
            y = y3 + 0.33 * (y3 - y2)
            width = 1.5 * max(math.sqrt((x3 - x2) ** 2 + (y3 - y2) ** 2),
                              0.9 * math.sqrt((x2 - x1) ** 2 + (y2 - y1) ** 2))
            x, y = max(x - width / 2, 0), max(y - width / 2, 0)
            w1 = img_w - x if x + width > img_w else width
            w2 = img_h - y if y + width > img_h else width
            out.append([int(x), int(y), int(min(w1, w2)), is_left])
    return out


# --------------------------------------------------------------------------------------------
# synthetic scenes: maps injected at the heatmap_avg / paf_avg boundary
# --------------------------------------------------------------------------------------------
_TEMPLATE = np.array([  # 18 joints (x, y) in a 100x220 box: nose, neck, Rsho, Relb, Rwri, Lsho, Lelb, Lwri,
    [50, 20], [50, 50], [30, 52], [22, 85], [18, 115], [70, 52], [78, 85], [82, 115],   # Rhip, Rkne, Rank, Lhip,
    [38, 115], [36, 160], [35, 205], [62, 115], [64, 160], [65, 205],                   # Lkne, Lank, Reye, Leye,
    [45, 14], [55, 14], [40, 18], [60, 18]], dtype=np.float64)                          # Rear, Lear


def synthetic_scene(H=720, W=1280, grid=(10, 5), seed=0, jitter=1.0, dtype=np.float32):
    """grid=(cols,rows) people laid out on a grid; joint blobs exp(-d^2/(2*3^2)); PAF = limb unit vector
    within +-3 px of the segment.  Returns (heatmap_avg (H,W,19), paf_avg (H,W,38)) as float64 arrays
    holding `dtype`-representable values, and the joint positions (P,18,2)."""
    rng = np.random.default_rng(seed)
    cols, rows = grid
    heat = np.zeros((H, W, 19), dtype=np.float64)
    paf = np.zeros((H, W, 38), dtype=np.float64)
    cnt = np.zeros((H, W, 19), dtype=np.int32)
    cw, ch = W / cols, H / rows
    s = min(cw / 110.0, ch / 230.0)
    yy, xx = np.mgrid[0:H, 0:W]
    people = []
    for r in range(rows):
        for c in range(cols):
            org = np.array([c * cw + (cw - 100 * s) / 2, r * ch + (ch - 220 * s) / 2])
            joints = org + _TEMPLATE * s + rng.normal(0, jitter, _TEMPLATE.shape)
            people.append(joints)
            for j, (x, y) in enumerate(joints):
                x0, x1 = max(int(x) - 15, 0), min(int(x) + 16, W)
                y0, y1 = max(int(y) - 15, 0), min(int(y) + 16, H)
                d2 = (xx[y0:y1, x0:x1] - x) ** 2 + (yy[y0:y1, x0:x1] - y) ** 2
                heat[y0:y1, x0:x1, j] = np.maximum(heat[y0:y1, x0:x1, j], np.exp(-d2 / 18.0))
            for k, (pa, pb) in enumerate(LIMB_SEQ):
                a, b = joints[pa - 1], joints[pb - 1]
                v = b - a
                L = np.linalg.norm(v)
                if L < 1e-6:
                    continue
                u = v / L
                x0, x1 = max(int(min(a[0], b[0])) - 4, 0), min(int(max(a[0], b[0])) + 5, W)
                y0, y1 = max(int(min(a[1], b[1])) - 4, 0), min(int(max(a[1], b[1])) + 5, H)
                dx, dy = xx[y0:y1, x0:x1] - a[0], yy[y0:y1, x0:x1] - a[1]
                along, perp = dx * u[0] + dy * u[1], np.abs(dx * u[1] - dy * u[0])
                m = (along >= 0) & (along <= L) & (perp <= 3)
                cx, cy = PAF_CH[k]
                paf[y0:y1, x0:x1, cx][m] += u[0]
                paf[y0:y1, x0:x1, cy][m] += u[1]
                cnt[y0:y1, x0:x1, k][m] += 1
    for k in range(19):
        cx, cy = PAF_CH[k]
        n = np.maximum(cnt[:, :, k], 1)
        paf[:, :, cx] /= n
        paf[:, :, cy] /= n
    heat[:, :, 18] = 1 - heat[:, :, :18].max(-1)
    heat = heat.astype(dtype).astype(np.float64)
    paf = paf.astype(dtype).astype(np.float64)
    return heat, paf, np.array(people)
