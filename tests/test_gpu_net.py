"""The CNNs (csrc/conv_tc.cu strung together by csrc/net.cu) against the oracle's torch-CPU restatement of
src/model.py, with random-init weights of the same architecture.

Tolerances (north_star: "heatmaps and PAFs within 1e-2 relative (bf16 compute)"), relative = max|dev-ref| /
max|ref| per output tensor:
  * vs the oracle run with bf16-rounded weights/activations (same numerics, different summation order): 4e-3
  * vs the fp32 oracle, default PyTorch init (BASELINE config): 1e-2
"""
import numpy as np
import pytest
import torch

from oracle import openpose_oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def body_default():
    from pytorch_openpose_b200 import _lib
    sd = O.make_weights("body", 0)
    net = _lib.Net(_lib.NET_BODY, sd)
    return sd, net, net.session()


@pytest.mark.parametrize("hp,wp", [(48, 64), (184, 248)])
def test_body_net_default_init(body_default, hp, wp):
    from tests import gpu_util as G
    sd, net, sess = body_default
    img = np.random.default_rng(0).integers(0, 256, (1, hp, wp, 3), dtype=np.uint8)
    paf, heat = G.net_forward(sess, img)
    x = torch.from_numpy(np.transpose(img.astype(np.float32), (0, 3, 1, 2)) / 256 - 0.5)
    rp, rh = O.body_net(x, sd)
    bp, bh = O.body_net(x, sd, bf16=True)
    rp, rh, bp, bh = (t.numpy().transpose(0, 2, 3, 1) for t in (rp, rh, bp, bh))
    assert heat.min() >= 0.0                                       # stage-6 L2 ReLU quirk
    assert _rel(paf, bp) <= 4e-3 and _rel(heat, bh) <= 4e-3
    assert _rel(paf, rp) <= 1e-2 and _rel(heat, rh) <= 1e-2


def test_body_net_kaiming_structured():
    """Kaiming weights give maps with real spatial structure; the device must track the bf16-emulating oracle
    closely, and the error against fp32 is reported (bf16 rounding through ~50 layers)."""
    from pytorch_openpose_b200 import _lib
    from tests import gpu_util as G
    sd = O.make_weights("body", 2, "kaiming")
    sess = _lib.Net(_lib.NET_BODY, sd).session()
    rng = np.random.default_rng(1)
    import cv2
    img = cv2.GaussianBlur(rng.integers(0, 256, (96, 128, 3), dtype=np.uint8), (0, 0), 4)[None]
    paf, heat = G.net_forward(sess, img)
    x = torch.from_numpy(np.transpose(img.astype(np.float32), (0, 3, 1, 2)) / 256 - 0.5)
    bp, bh = (t.numpy().transpose(0, 2, 3, 1) for t in O.body_net(x, sd, bf16=True))
    rp, rh = (t.numpy().transpose(0, 2, 3, 1) for t in O.body_net(x, sd))
    print("kaiming: rel vs bf16-oracle paf %.2e heat %.2e ; vs fp32 paf %.2e heat %.2e ; spatial std %.3f"
          % (_rel(paf, bp), _rel(heat, bh), _rel(paf, rp), _rel(heat, rh), float(rh.std())))
    assert _rel(paf, bp) <= 2e-2 and _rel(heat, bh) <= 2e-2
    assert _rel(paf, rp) <= 5e-2 and _rel(heat, rh) <= 5e-2


def test_hand_net_default_init_batched():
    from pytorch_openpose_b200 import _lib
    from tests import gpu_util as G
    sd = O.make_weights("hand", 0)
    sess = _lib.Net(_lib.NET_HAND, sd).session()
    img = np.random.default_rng(2).integers(0, 256, (2, 64, 64, 3), dtype=np.uint8)
    _, heat = G.net_forward(sess, img)
    x = torch.from_numpy(np.transpose(img.astype(np.float32), (0, 3, 1, 2)) / 256 - 0.5)
    ref = O.hand_net(x, sd).numpy().transpose(0, 2, 3, 1)
    bref = O.hand_net(x, sd, bf16=True).numpy().transpose(0, 2, 3, 1)
    assert _rel(heat, bref) <= 4e-3
    assert _rel(heat, ref) <= 1e-2


def test_missing_layer_is_keyerror():
    from pytorch_openpose_b200 import Body
    sd = O.make_weights("body", 0)
    del sd["Mconv3_stage4_L2.weight"]
    with pytest.raises(KeyError):
        Body(sd)
