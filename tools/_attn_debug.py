import os, sys, faulthandler
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.enable()
import torch
from msrflute_b200.ops import _ext, attention_ops as A
C = _ext.load(required=True)
B, H, S = 2, 3, 128
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda: (torch.randn(B, S, H, 64, device="cuda", generator=g) * 0.8).to(torch.bfloat16).transpose(1, 2)
q, k, v = mk(), mk(), mk()
seed = torch.zeros(1, dtype=torch.int64, device="cuda")
print("fwd...", flush=True)
out, lse = C.attention_fwd(q, k, v, None, 0.125, 0.0, seed)
torch.cuda.synchronize()
ref = A.attention_reference(q, k, v)
print("fwd rel err", float((out.float() - ref).norm() / ref.norm()), "lse", lse[0, 0, :4].tolist(), flush=True)
s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * 0.125
print("lse ref", torch.logsumexp(s, -1)[0, 0, :4].tolist(), flush=True)
go = (torch.randn(out.shape, device="cuda") * 0.5).to(torch.bfloat16)
print("bwd...", flush=True)
dq, dk, dv = C.attention_bwd(q, k, v, out, lse, go, None, 0.125, 0.0, seed)
torch.cuda.synchronize()
print("bwd done", flush=True)
qq, kk, vv = [t.detach().clone().requires_grad_(True) for t in (q, k, v)]
r = A.attention_reference(qq, kk, vv); r.backward(go.float())
for n, a, b in (("dq", dq, qq.grad), ("dk", dk, kk.grad), ("dv", dv, vv.grad)):
    a = a.float().transpose(1, 2)
    print(n, "rel err", float((a - b.float()).norm() / b.float().norm()), flush=True)
