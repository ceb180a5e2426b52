"""Sweep MN-major shared-memory descriptor variants of the SlotNet GEMM kernel on a GPU box and report which one
reproduces PyTorch's weight / data gradients (one gpurun call instead of one rebuild per hypothesis).

    python tools/debug_slotnet.py            # writes gpurun_out/slotnet_sweep.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

VARIANTS = {
    "lbo4096_sbo512_k64_l1": [4096, 512, 64, 1],
    "lbo512_sbo4096_k64_l1": [512, 4096, 64, 1],
    "lbo4096_sbo1024_k64_l1": [4096, 1024, 64, 1],
    "lbo4096_sbo512_k32_l1": [4096, 512, 32, 1],
    "lbo128_sbo512_k64_l1": [128, 512, 64, 1],
}


def main():
    from msrflute_b200.models.slotnet_resnet import SlotNetResNet
    import test_slotnet_gpu as T
    out = []
    for name, desc in VARIANTS.items():
        SlotNetResNet.MN_DESC = desc
        try:
            worst, lines = T.run_compare(S=2, B=20)
            keep = [ln for ln in lines if "fc.weight" in ln or "layer4.1.bn2" in ln or "layer4.1.conv2.weight" in ln
                    or "layer1.0.conv1.weight" in ln or "net.conv1.weight" in ln or "layer2.0.conv1.weight" in ln]
            out.append("### {} {} -> worst {}".format(name, desc, worst))
            out.extend(keep[:8])
        except Exception as e:  # keep sweeping
            out.append("### {} {} -> EXCEPTION {}".format(name, desc, str(e)[:300]))
        torch.cuda.synchronize()
    SlotNetResNet.MN_DESC = None
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "slotnet_sweep.txt"), "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
