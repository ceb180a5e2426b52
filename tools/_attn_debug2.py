import os, sys, faulthandler, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) == 1:
    for mode in ("single", "multi_contig", "multi_hf", "multi_hf_nobias_sync"):
        r = subprocess.run([sys.executable, __file__, mode], capture_output=True, text=True)
        print("==", mode, "rc", r.returncode, (r.stdout + r.stderr)[-600:].replace("\n", " | "), flush=True)
    sys.exit(0)
faulthandler.enable()
import torch
from msrflute_b200.ops import attention_ops as A
mode = sys.argv[1]
if mode == "single":
    torch.autograd.set_multithreading_enabled(False)
B, H, S = 2, 3, 128
mk = lambda: (torch.randn(B, S, H, 64, device="cuda") * 0.8).to(torch.bfloat16)
if mode == "multi_contig":
    q, k, v = [mk().transpose(1, 2).contiguous().requires_grad_(True) for _ in range(3)]
else:
    q, k, v = [mk().transpose(1, 2).requires_grad_(True) for _ in range(3)]
out = A.attention(q, k, v)
torch.cuda.synchronize()
print("fwd ok", flush=True)
go = (torch.randn(out.shape, device="cuda") * 0.5).to(torch.bfloat16)
out.backward(go)
torch.cuda.synchronize()
print("bwd ok", float(q.grad.float().norm()), flush=True)
