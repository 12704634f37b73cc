"""`bodypose_model` / `handpose_model` as torch modules with the reference's state-dict names
(src/model.py:25-214), built from the layer table the native library exports.  They exist for callers that
import the classes directly (srcmx/Batch_model.py:109,350) and for `util.transfer`; the product inference path
does not run them -- it runs the sm_100a kernels behind `Body` / `Hand`."""
from collections import OrderedDict

import torch
import torch.nn as nn


def _sequential(layers):
    mods = OrderedDict()
    for spec in layers:
        if spec == "pool":
            mods["pool%d_stage1" % (sum(k.startswith("pool") for k in mods) + 1)] = nn.MaxPool2d(2, 2, 0)
            continue
        name, cin, cout, k, relu = spec
        mods[name] = nn.Conv2d(cin, cout, k, 1, k // 2)
        if relu:
            mods["relu_" + name] = nn.ReLU(inplace=True)
    return nn.Sequential(mods)


def _vgg_head(tail):
    head = [("conv1_1", 3, 64, 3, True), ("conv1_2", 64, 64, 3, True), "pool",
            ("conv2_1", 64, 128, 3, True), ("conv2_2", 128, 128, 3, True), "pool",
            ("conv3_1", 128, 256, 3, True), ("conv3_2", 256, 256, 3, True), ("conv3_3", 256, 256, 3, True),
            ("conv3_4", 256, 256, 3, True), "pool",
            ("conv4_1", 256, 512, 3, True), ("conv4_2", 512, 512, 3, True)]
    return head + tail


class bodypose_model(nn.Module):
    def __init__(self):
        super().__init__()
        self.model0 = _sequential(_vgg_head([("conv4_3_CPM", 512, 256, 3, True), ("conv4_4_CPM", 256, 128, 3, True)]))
        blocks = OrderedDict()
        for b, cout in ((1, 38), (2, 19)):
            blocks["model1_%d" % b] = [("conv5_%d_CPM_L%d" % (i, b), 128, 128, 3, True) for i in (1, 2, 3)] + [
                ("conv5_4_CPM_L%d" % b, 128, 512, 1, True), ("conv5_5_CPM_L%d" % b, 512, cout, 1, False)]
        for s in range(2, 7):
            for b, cout in ((1, 38), (2, 19)):
                # the reference's no_relu list misses Mconv7_stage6_L2 (src/model.py:30-33): it keeps its ReLU
                last_relu = (s == 6 and b == 2)
                blocks["model%d_%d" % (s, b)] = (
                    [("Mconv1_stage%d_L%d" % (s, b), 185, 128, 7, True)]
                    + [("Mconv%d_stage%d_L%d" % (i, s, b), 128, 128, 7, True) for i in (2, 3, 4, 5)]
                    + [("Mconv6_stage%d_L%d" % (s, b), 128, 128, 1, True),
                       ("Mconv7_stage%d_L%d" % (s, b), 128, cout, 1, last_relu)])
        built = {k: _sequential(v) for k, v in blocks.items()}      # creation order = reference RNG order
        for b in (1, 2):                                            # attribute order = reference state-dict order
            for s in range(1, 7):
                setattr(self, "model%d_%d" % (s, b), built["model%d_%d" % (s, b)])

    def forward(self, x):
        feat = self.model0(x)
        paf, heat = self.model1_1(feat), self.model1_2(feat)
        for s in range(2, 7):
            cat = torch.cat([paf, heat, feat], 1)
            paf = getattr(self, "model%d_1" % s)(cat)
            heat = getattr(self, "model%d_2" % s)(cat)
        return paf, heat


class handpose_model(nn.Module):
    def __init__(self):
        super().__init__()
        self.model1_0 = _sequential(_vgg_head([("conv4_3", 512, 512, 3, True), ("conv4_4", 512, 512, 3, True),
                                               ("conv5_1", 512, 512, 3, True), ("conv5_2", 512, 512, 3, True),
                                               ("conv5_3_CPM", 512, 128, 3, True)]))
        self.model1_1 = _sequential([("conv6_1_CPM", 128, 512, 1, True), ("conv6_2_CPM", 512, 22, 1, False)])
        for s in range(2, 7):
            setattr(self, "model%d" % s, _sequential(
                [("Mconv1_stage%d" % s, 150, 128, 7, True)]
                + [("Mconv%d_stage%d" % (i, s), 128, 128, 7, True) for i in (2, 3, 4, 5)]
                + [("Mconv6_stage%d" % s, 128, 128, 1, True), ("Mconv7_stage%d" % s, 128, 22, 1, False)]))

    def forward(self, x):
        feat = self.model1_0(x)
        out = self.model1_1(feat)
        for s in range(2, 7):
            out = getattr(self, "model%d" % s)(torch.cat([out, feat], 1))
        return out


def random_checkpoint(kind, seed=0):
    """A random-init checkpoint in the reference's file format (caffe-keyed flat dict, src/util.py:36-40), i.e. what
    `torch.manual_seed(seed); bodypose_model()` holds -- for benchmarks and smoke runs: no trained weights exist
    offline."""
    torch.manual_seed(seed)
    net = bodypose_model() if kind == "body" else handpose_model()
    return {k.split(".", 1)[1]: v.detach().clone() for k, v in net.state_dict().items()}
