"""Driver for `ncu`: one warm-up step of the flagship SlotNet program, then selected ops once each.
    ncu --set full --clock-control none --import-source on -k regex:sn_gemm_kernel -s <gemm launches per step> -c N ...
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ops = [int(a) for a in sys.argv[1:]] or [4, 15, 25, 40]
    from msrflute_b200.models.resnet_gn import RESNET
    from msrflute_b200.models.slotnet_resnet import SlotNetResNet
    from msrflute_b200.parallel.arena import ArenaLayout
    S = 10
    torch.manual_seed(0)
    model = RESNET({"group_norm": 2, "num_classes": 1000}).cuda()
    layout = ArenaLayout.from_module(model)
    plan = SlotNetResNet.plan(model, layout)
    W = torch.randn(S, plan["numel"], device="cuda") * 0.05
    G = torch.zeros(S, plan["numel"], device="cuda")
    net = SlotNetResNet(model, W, G, plan, batch=20)
    net.prog.set_side_stream(False)
    x = torch.rand(S * 20, 3, 32, 32, device="cuda") * 255
    y = torch.randint(0, 1000, (S * 20,), device="cuda")
    net.step(x, y)
    torch.cuda.synchronize()
    print("gemm launches per step:", sum(1 for n in net.op_names if n.split()[0] in ("fprop", "dgrad", "wgrad")))
    for i in ops:
        print("op", i, net.op_names[i])
        net.prog.run(i, i + 1)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
