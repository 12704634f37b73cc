"""PAF grouping on the crowded synthetic scene (BASELINE config 5) -- for ncu launch lists."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import openpose_oracle as O            # noqa: E402
from pytorch_openpose_b200 import _lib             # noqa: E402

H, W = 720, 1280
heat, paf, _ = O.synthetic_scene(H, W, (10, 5), seed=0)
d_heat = torch.from_numpy(np.ascontiguousarray(heat.transpose(2, 0, 1), dtype=np.float32)).cuda()
d_paf = torch.from_numpy(np.ascontiguousarray(paf.transpose(2, 0, 1), dtype=np.float32)).cuda()
torch.cuda.synchronize()
ms, nc, ns = ctypes.c_float(), ctypes.c_int(), ctypes.c_int()
_lib.check(_lib.lib().opb_bench_grouping(_lib.context(0), d_heat.data_ptr(), d_paf.data_ptr(), H, W, int(sys.argv[1]) if len(sys.argv) > 1 else 5,
                                         ctypes.byref(ms), ctypes.byref(nc), ctypes.byref(ns)))
print("grouping %.3f ms/frame, %d candidates, %d persons" % (ms.value, nc.value, ns.value))
