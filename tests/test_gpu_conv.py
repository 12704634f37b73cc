"""tcgen05 implicit-GEMM convolution (csrc/conv_tc.cu) against a plain PyTorch fp32 reference of the same op
(torch.conv2d is what the reference network calls, src/model.py:15-17), on bf16-representable inputs so that
only the accumulation order differs.  Tolerance: 2e-3 of the output scale for fp32 outputs (fp32 accumulate over
K <= 9408 terms), plus one bf16 ulp (2^-8 relative) for bf16 outputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _reference(x_nhwc, w, b, relu, pool):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x = x_nhwc.float().permute(0, 3, 1, 2).contiguous()
    wq = w.to(torch.bfloat16).float().cuda()
    y = F.conv2d(x.double(), wq.double(), b.double().cuda(), 1, w.shape[2] // 2).float()   # fp64: an exact reference
    if relu:
        y = torch.relu(y)
    if pool:
        y = F.max_pool2d(y, 2, 2)
    return y.permute(0, 2, 3, 1).contiguous()


CASES = [
    # n, h, w, cin, cout, k, relu, pool, fp32
    (1, 8, 16, 64, 128, 1, True, False, False),      # exactly one tile, one k-block
    (1, 8, 16, 64, 64, 3, True, False, False),       # BLOCK_N=64 path, 9 k-blocks
    (1, 23, 41, 128, 128, 3, True, False, False),    # ragged tiles, s=0.5 feature map
    (1, 46, 82, 192, 128, 7, True, False, False),    # stage Mconv1: K = 49*192
    (1, 23, 41, 128, 38, 1, False, False, True),     # final PAF layer: fp32, cout padded to 64, no ReLU
    (1, 23, 41, 128, 19, 1, True, False, True),      # final heat layer (stage-6 L2 keeps its ReLU)
    (1, 24, 40, 64, 64, 3, True, True, False),       # fused 2x2 max-pool
    (2, 16, 24, 256, 256, 3, True, True, False),     # batch 2, two N tiles, pool
    (1, 12, 20, 512, 512, 3, True, False, False),    # conv4_2-like: K = 4608, 4 N tiles
    (3, 9, 7, 128, 512, 1, True, False, False),      # tiny maps smaller than the TMA box
    (1, 92, 164, 128, 128, 7, True, False, False),   # s=2.0 stage layer: 3 waves of tiles
    (2, 50, 70, 64, 64, 3, True, False, False),      # conv1_2 shape class (see test_conv_pair_resident_variant)
    (1, 184, 328, 64, 64, 3, True, True, False),     # conv1_2 at scale 0.5 with its fused pool
    (2, 23, 23, 128, 128, 3, True, False, False),    # hand map: last tile column AND last tile row are half tiles (mixed orientations)
    (1, 69, 69, 192, 128, 7, True, False, False),    # the same at 7x7
    (2, 72, 40, 128, 128, 3, True, True, False),     # 8 rows left in the last tile row, with the fused pool
    (3, 8, 8, 128, 128, 3, True, False, True),       # a single half tile per image (odd tile count: padding tile in the pair)
    (1, 40, 72, 128, 256, 3, True, False, False),    # 8 columns left in the last tile column, two N tiles
]


IMPLS = {"tap": 2, "patch0": 3, "patch1": 4, "pair": 5}       # per-tap tiles / patch-resident MODE 0 / MODE 1 (opb_conv2d impl)


@pytest.mark.parametrize("impl", sorted(IMPLS))
@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d_%dx%d_c%d_o%d_k%d_r%d_p%d_f%d" % tuple(int(v) for v in c))
def test_conv_tc_matches_torch(case, impl):
    from tests import gpu_util as G
    n, h, w, cin, cout, k, relu, pool, fp32 = case
    if k == 1 and impl != "tap":
        pytest.skip("1x1 layers always use the per-tap kernel")
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    x = (torch.randn(n, h, w, cin, generator=g) * 0.5).to(torch.bfloat16).cuda()
    wt = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    ref = _reference(x, wt, b, relu, pool)
    out = G.conv2d(x, wt, b, relu, pool, fp32, impl=IMPLS[impl]).float()
    assert not torch.isnan(out[..., :cout]).any(), "kernel left output elements unwritten"
    if out.shape[-1] > cout:
        assert (out[..., cout:] == 0).all(), "padded output channels must be written as zeros"
    scale = ref.abs().max().item()
    err = (out[..., :cout] - ref).abs().max().item()
    tol = 2e-3 * scale + (0 if fp32 else scale * 2 ** -8)
    assert err <= tol, "max err %.3e (scale %.3e, tol %.3e)" % (err, scale, tol)


@pytest.mark.parametrize("case", [c for c in CASES if c[3] == 64 and c[4] == 64 and c[5] == 3],
                         ids=lambda c: "n%d_%dx%d_p%d" % (c[0], c[1], c[2], int(c[7])))
def test_conv_pair_resident_variant(case, monkeypatch):
    """The opt-in conv1_2 variant of the CTA-pair kernel (weights resident in shared memory, one 24-column patch per
    tile for all three dx; measured slower than the default, kept for A/B -- DESIGN.md section 6) obeys the same contract."""
    from tests import gpu_util as G
    monkeypatch.setenv("OPB_CONV12_RESIDENT", "1")
    n, h, w, cin, cout, k, relu, pool, fp32 = case
    g = torch.Generator().manual_seed(77)
    x = (torch.randn(n, h, w, cin, generator=g) * 0.5).to(torch.bfloat16).cuda()
    wt = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    ref = _reference(x, wt, b, relu, pool)
    out = G.conv2d(x, wt, b, relu, pool, fp32, impl=5).float()
    scale = ref.abs().max().item()
    assert (out[..., :cout] - ref).abs().max().item() <= 2e-3 * scale + scale * 2 ** -8


@pytest.mark.parametrize("shape", [(1, 24, 40), (2, 50, 70), (1, 184, 328), (3, 16, 18), (1, 2, 2), (1, 38, 250)],
                         ids=lambda s: "n%d_%dx%d" % s)
def test_conv_pair_wide_pixel_form(shape):
    """conv1_2 as the networks run it: pairs of adjacent columns as one 128-channel pixel, 128 output columns, all-zero
    weight chunks skipped, pool over the column halves and row pairs (net.cu wide_pool_weights) -- same contract; against
    the plain N = 64 launch of the same layer only the fp32 accumulation order of the nine taps differs (one bf16 ulp)."""
    from tests import gpu_util as G
    n, h, w = shape
    g = torch.Generator().manual_seed(1000 + h * w)
    x = (torch.randn(n, h, w, 64, generator=g) * 0.5).to(torch.bfloat16).cuda()
    wt = torch.randn(64, 64, 3, 3, generator=g) * (2.0 / (64 * 9)) ** 0.5
    b = torch.randn(64, generator=g) * 0.1
    ref = _reference(x, wt, b, True, True)
    out = G.conv2d(x, wt, b, True, True, False, impl=6).float()
    plain = G.conv2d(x, wt, b, True, True, False, impl=5).float()
    assert out.shape == ref.shape and not torch.isnan(out).any()
    scale = ref.abs().max().item()
    assert (out - ref).abs().max().item() <= 2e-3 * scale + scale * 2 ** -8
    assert (out - plain).abs().max().item() <= scale * 2 ** -8


def test_conv_direct_crosscheck():
    """The scalar cross-check kernel obeys the same contract (used to bisect tensor-core issues)."""
    from tests import gpu_util as G
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(1, 10, 12, 64, generator=g)).to(torch.bfloat16).cuda()
    wt = torch.randn(24, 64, 3, 3, generator=g) * 0.05
    b = torch.randn(24, generator=g)
    ref = _reference(x, wt, b, True, False)
    out = G.conv2d(x, wt, b, True, False, True, impl=1)
    assert (out[..., :24] - ref).abs().max().item() < 1e-3 * ref.abs().max().item()
