import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
def barrier(): torch.cuda.synchronize()
print(json.dumps(bench.extra_bodyhand_c4(0, 0, barrier, lambda x: x, 1)))
