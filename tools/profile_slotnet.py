"""Per-launch device time of the SlotNet program (flagship shapes: S slots x batch 20), measured with CUDA events
around each op replayed in isolation, plus the whole step.  Writes gpurun_out/slotnet_ops.txt.

    python tools/profile_slotnet.py [S]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def variants(S):
    """Whole-step graph time under measurement hooks: accumulators per tile, MMAs off, TMA loads off."""
    from msrflute_b200.models.resnet_gn import RESNET
    from msrflute_b200.models.slotnet_resnet import SlotNetResNet, SlotProgramBuilder
    from msrflute_b200.parallel.arena import ArenaLayout
    out = []
    for name, nacc, dbg in (("nacc=1", 1, 0), ("nacc=2", 2, 0), ("nacc=default(4)", None, 0), ("no MMA", None, 1),
                            ("no TMA", None, 2), ("no MMA, no TMA", None, 3)):
        SlotProgramBuilder.NACC, SlotProgramBuilder.DBG = nacc, dbg
        torch.manual_seed(0)
        model = RESNET({"group_norm": 2, "num_classes": 1000}).cuda()
        layout = ArenaLayout.from_module(model)
        plan = SlotNetResNet.plan(model, layout)
        W = torch.randn(S, plan["numel"], device="cuda") * 0.05
        G = torch.zeros(S, plan["numel"], device="cuda")
        net = SlotNetResNet(model, W, G, plan, batch=20)
        x = torch.rand(S * 20, 3, 32, 32, device="cuda") * 255
        y = torch.randint(0, 1000, (S * 20,), device="cuda")
        net.step(x, y)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            net.prog.run(0, -1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        e1.synchronize()
        out.append("variant {:18s}: whole step {:.1f} us".format(name, e0.elapsed_time(e1) / 20 * 1e3))
        del net, g
    SlotProgramBuilder.NACC, SlotProgramBuilder.DBG = None, 0
    return out


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    from msrflute_b200.models.resnet_gn import RESNET
    from msrflute_b200.models.slotnet_resnet import SlotNetResNet
    from msrflute_b200.parallel.arena import ArenaLayout
    torch.manual_seed(0)
    model = RESNET({"group_norm": 2, "num_classes": 1000}).cuda()
    layout = ArenaLayout.from_module(model)
    plan = SlotNetResNet.plan(model, layout)
    W = torch.zeros(S, plan["numel"], device="cuda")
    G = torch.zeros(S, plan["numel"], device="cuda")
    flat = torch.zeros(layout.padded_numel, device="cuda")
    for p, o, k in zip(model.parameters(), layout.offsets, layout.sizes):
        flat[o:o + k] = p.detach().reshape(-1)
    im = plan["index_map"].cuda().long()
    W[:, im >= 0] = flat[im[im >= 0]]
    net = SlotNetResNet(model, W, G, plan, batch=20)
    x = torch.rand(S * 20, 3, 32, 32, device="cuda") * 255
    y = torch.randint(0, 1000, (S * 20,), device="cuda")
    for _ in range(3):
        G.zero_()
        net.step(x, y)
    torch.cuda.synchronize()
    lines = []
    names = net.op_names + ["op"] * (net.n_ops - len(net.op_names))
    total = 0.0
    reps = 10
    flush = torch.empty(64 * 1024 * 1024, device="cuda")          # 256 MB > L2
    for i in range(net.n_ops):
        ms = 0.0
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            net.prog.run(i, i + 1)
            e1.record()
            e1.synchronize()
            ms += e0.elapsed_time(e1)
        ms /= reps
        total += ms
        lines.append("{:3d} {:8.1f} us  {}{}".format(i, ms * 1e3, names[i], "   [fwd]" if i < net.n_fwd else ""))
    # whole step, serial and with the side stream, captured in a graph
    for side in (False, True):
        net.prog.set_side_stream(side)
        g = torch.cuda.CUDAGraph()
        G.zero_()
        net.step(x, y)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            net.prog.run(0, -1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        e1.synchronize()
        lines.append("graph replay of the whole step, side stream {}: {:.1f} us".format(side, e0.elapsed_time(e1) / 20 * 1e3))
    lines.append("sum of isolated ops (cold L2): {:.1f} us over {} ops".format(total * 1e3, net.n_ops))
    del net
    lines.extend(variants(S))
    out = "\n".join(lines)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "slotnet_ops.txt"), "w") as f:
        f.write(out + "\n")
    print(out)


if __name__ == "__main__":
    main()
