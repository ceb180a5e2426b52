"""GPU-idle analysis of flagship rounds with torch.profiler (CUPTI): where does the wall time of a round go?
Writes gpurun_out/round_timeline.txt (kernel-busy union, largest idle gaps and what surrounded them, top kernels)."""
import json
import os
import sys
import collections

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity

from bench import build_flagship, build_job


def main():
    rounds = int(os.environ.get("ROUNDS", "3"))
    task = os.environ.get("TASK", "cv_resnet_fedcifar100")
    if task == "cv_resnet_fedcifar100":
        job = build_flagship(n_clients_per_round=10, users=500, norm="gn")
    else:
        job = build_job(task)
    for _ in range(int(os.environ.get("SETUP", "8"))):
        job.run_round()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(rounds):
            job.run_round()
        torch.cuda.synchronize()
    os.makedirs("gpurun_out", exist_ok=True)
    path = "gpurun_out/round_trace.json"
    tag = "" if task == "cv_resnet_fedcifar100" else "_" + task
    prof.export_chrome_trace(path)
    tr = json.load(open(path))["traceEvents"]
    gpu = [e for e in tr if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    cpu = [e for e in tr if e.get("ph") == "X" and e.get("cat") in ("cpu_op", "user_annotation", "cuda_runtime", "cuda_driver")]
    gpu.sort(key=lambda e: e["ts"])
    t0, t1 = gpu[0]["ts"], max(e["ts"] + e["dur"] for e in gpu)
    out = []
    busy, cur_end, gaps = 0.0, t0, []
    prev = None
    for e in gpu:
        s, d = e["ts"], e["dur"]
        if s > cur_end:
            gaps.append((s - cur_end, cur_end - t0, prev["name"][:60] if prev else "", e["name"][:60]))
            busy += d
            cur_end = s + d
        else:
            if s + d > cur_end:
                busy += s + d - cur_end
                cur_end = s + d
        prev = e
    span = t1 - t0
    out.append("rounds {}  span {:.2f} ms  gpu busy {:.2f} ms ({:.1f}%)  kernels {}  per round: span {:.2f} ms busy {:.2f} ms".format(
        rounds, span / 1e3, busy / 1e3, 100 * busy / span, len(gpu), span / 1e3 / rounds, busy / 1e3 / rounds))
    hist = collections.Counter()
    for g in gaps:
        b = "<2us" if g[0] < 2 else "<5us" if g[0] < 5 else "<20us" if g[0] < 20 else "<100us" if g[0] < 100 else ">=100us"
        hist[b] += g[0]
    out.append("idle by gap size (ms): " + ", ".join("{} {:.2f}".format(k, v / 1e3) for k, v in sorted(hist.items())))
    out.append("largest gaps: (gap us, at ms, after kernel -> before kernel)")
    for g in sorted(gaps, reverse=True)[:40]:
        out.append("  {:9.1f} us at {:8.2f} ms   {}  ->  {}".format(g[0], g[1] / 1e3, g[2], g[3]))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in gpu:
        agg[e["name"][:80]][0] += 1
        agg[e["name"][:80]][1] += e["dur"]
    out.append("top kernels by total time (per round):")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        out.append("  {:9.1f} us  {:5d}  {:7.2f} us/launch  {}".format(t / rounds, n // rounds, t / n, k))
    # CPU-side: biggest ops while GPU idle are the suspects
    cagg = collections.defaultdict(lambda: [0, 0.0])
    for e in cpu:
        if e.get("cat") in ("user_annotation", "cpu_op"):
            cagg[e["name"][:60]][0] += 1
            cagg[e["name"][:60]][1] += e["dur"]
    out.append("top CPU ops (inclusive, per round):")
    for k, (n, t) in sorted(cagg.items(), key=lambda kv: -kv[1][1])[:25]:
        out.append("  {:9.1f} us  {:5d}  {}".format(t / rounds, n // rounds, k))
    txt = "\n".join(out)
    print(txt)
    open("gpurun_out/round_timeline{}.txt".format(tag), "w").write(txt + "\n")
    os.remove(path)
    from msrflute_b200.utils.async_ckpt import get_checkpointer
    get_checkpointer().close()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
