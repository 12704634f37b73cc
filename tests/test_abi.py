"""CPU checks of the boundary: the library loads, exports every symbol include/openpose_b200.h declares, reports
its layer tables, and refuses to run without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import openpose_oracle as O
from pytorch_openpose_b200 import _lib, model, util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported_and_bound():
    header = open(os.path.join(ROOT, "include", "openpose_b200.h")).read()
    declared = set(re.findall(r"\b(opb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), "library does not export " + name
    assert declared == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert lib.opb_abi_version() == 1


def test_layer_tables_match_reference_architecture():
    for kind, name in ((_lib.NET_BODY, "body"), (_lib.NET_HAND, "hand")):
        table = _lib.layer_table(kind)
        ref = [l for _, block in (O.body_layers() if name == "body" else O.hand_layers()) for l in block if l != "pool"]
        assert sorted((n, co, ci, k, r) for n, co, ci, k, r in table) == sorted((n, co, ci, k, r) for n, ci, co, k, r in ref)


def test_torch_modules_have_reference_state_dict():
    sd = O.make_weights("body", 0)
    m = model.bodypose_model()
    m.load_state_dict(util.transfer(m, sd))
    x = torch.rand(1, 3, 32, 40) - 0.5
    with torch.no_grad():
        paf, heat = m(x)
    rp, rh = O.body_net(x, sd)
    assert torch.equal(paf, rp) and torch.equal(heat, rh)
    mh = model.handpose_model()
    sdh = O.make_weights("hand", 0)
    mh.load_state_dict(util.transfer(mh, sdh))
    with torch.no_grad():
        assert torch.equal(mh(x), O.hand_net(x, sdh))
    with pytest.raises(KeyError):
        util.transfer(m, {})


def test_host_helpers_match_oracle(golden):
    img = np.random.default_rng(0).integers(0, 256, (45, 67, 3), dtype=np.uint8)
    a, pad = util.padRightDownCorner(img, 8, 128)
    b, pad2 = O.pad_right_down(img)
    assert np.array_equal(a, b) and pad == pad2 == [0, 0, 3, 5]
    g = golden("body_postproc")
    hands = util.handDetect(g["cand_p8"], g["subset_p8"], np.zeros((360, 640, 3), np.uint8))
    assert np.array_equal(np.array([[x, y, w, int(l)] for x, y, w, l in hands]), g["hands_p8"])
    m = np.array([[1., 5., 5.], [5., 2., 0.]])
    assert util.npmax(m) == (0, 1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback():
    with pytest.raises(_lib.OpbError) as e:
        _lib.context(0)
    assert e.value.code == _lib.OPB_ERR_NO_DEVICE


def test_header_is_plain_c_and_links(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++ / torch types) and a C program must link against
    the shared library and call into it (no GPU needed for these calls)."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "openpose_b200.h"\n'
                   'int main(void) {\n'
                   '    opb_context* ctx = 0;\n'
                   '    int h, w, hp, wp;\n'
                   '    if (opb_abi_version() != OPB_ABI_VERSION) return 1;\n'
                   '    if (opb_scale_dims(720, 1280, 0.5, &h, &w, &hp, &wp) != OPB_OK) return 2;\n'
                   '    printf("%d %d %d %d %d\\n", h, w, hp, wp, opb_context_create(0, &ctx));\n'
                   '    return 0;\n}\n')
    libdir = os.path.join(ROOT, "pytorch_openpose_b200")
    exe = str(tmp_path / "abi")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                    "-L", libdir, "-lopenpose_b200", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    assert out[:4] == ["184", "327", "184", "328"]                    # C2 scale 0.5 -> 327x184 padded 328x184
    if not torch.cuda.is_available():
        assert int(out[4]) == _lib.OPB_ERR_NO_DEVICE
