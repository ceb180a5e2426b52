"""Whole local step of the flagship (S slots x batch 20): per-launch program vs the persistent step kernels.
Writes gpurun_out/slotnet_fused.txt.

    python tools/profile_fused.py [S]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def build(S, env):
    from msrflute_b200.models.resnet_gn import RESNET
    from msrflute_b200.models.slotnet_resnet import SlotNetResNet
    from msrflute_b200.parallel.arena import ArenaLayout
    os.environ.update(env)
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
    return net, W, G


def timed_graph(fn, reps=30):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    lines = []
    for tag, env in (("per-launch (side stream + PDL)", {"FLUTE_SLOTNET_FUSED": "0"}),
                     ("fused, 2 CTAs/SM", {"FLUTE_SLOTNET_FUSED": "1", "FLUTE_SLOTNET_FUSED_CTAS": "2"}),
                     ("fused, 1 CTA/SM", {"FLUTE_SLOTNET_FUSED": "1", "FLUTE_SLOTNET_FUSED_CTAS": "1"})):
        try:
            net, W, G = build(S, env)
            x0 = torch.rand(S * 20, 3, 32, 32, device="cuda") * 255
            y0 = torch.randint(0, 1000, (S * 20,), device="cuda")
            net.step(x0, y0)
            torch.cuda.synchronize()
        except Exception as exc:  # noqa: BLE001
            lines.append("{:34s}: FAILED {}: {}".format(tag, type(exc).__name__, str(exc)[:300]))
            print(lines[-1])
            continue
        x = torch.rand(S * 20, 3, 32, 32, device="cuda") * 255
        y = torch.randint(0, 1000, (S * 20,), device="cuda")
        net.step(x, y)
        torch.cuda.synchronize()
        whole = timed_graph(lambda: net.prog.run(0, -1))
        lines.append("{:34s}: whole step {:7.1f} us, {} launches".format(tag, whole, int(net.prog.run(0, -1))))
        if net.fused:
            fwd = timed_graph(lambda: net.prog.run(net.fwd_fuse_begin, net.fwd_fuse_end))
            bwd = timed_graph(lambda: net.prog.run(net.bwd_fuse_begin, net.bwd_fuse_end))
            info = net.prog.mega_info()
            lines.append("    forward range  {:7.1f} us  ({} ops, {} phases, {} CTAs)".format(
                fwd, int(info[0][1] - info[0][0]), int(info[0][2]), int(info[0][3])))
            lines.append("    backward range {:7.1f} us  ({} ops, {} phases, {} CTAs)".format(
                bwd, int(info[1][1] - info[1][0]), int(info[1][2]), int(info[1][3])))
            lines.append("    mega_info (begin, end, phases, ctas, est, max occupancy, smem): {}".format(info))
            rest = [(i, timed_graph(lambda i=i: net.prog.run(i, i + 1))) for i in
                    list(range(0, net.fwd_fuse_begin)) + list(range(net.fwd_fuse_end, net.bwd_fuse_begin)) +
                    list(range(net.bwd_fuse_end, net.n_ops))]
            lines.append("    other launches: " + ", ".join("op{} {:.1f}".format(i, t) for i, t in rest))
        torch.cuda.synchronize()
        del net
    out = "\n".join(lines)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "slotnet_fused.txt"), "w") as f:
        f.write(out + "\n")
    print(out)


if __name__ == "__main__":
    main()
