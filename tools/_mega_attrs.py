import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msrflute_b200.ops import _ext as e
C = e.load(required=True)
torch.zeros(1, device="cuda")
for smem in (0, 40000, 60000, 100000):
    print(smem, C.slotnet_mega_attrs(smem))
