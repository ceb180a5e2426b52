"""Micro-benchmarks of the hand-written kernels against their rooflines (B200_PROFILING.md timing hygiene:
warm-up, CUDA events, L2 flush between iterations, max/medians reported).  Writes gpurun_out/kernels.json."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from msrflute_b200.ops import _ext, arena_ops  # noqa: E402

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "fallback": True}
FLUSH = None


def timeit(fn, iters=20, warmup=5, flush=True):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def bench_gemm(out):
    ext = _ext.load(required=True)
    rows = []
    for (M, N, K) in [(4096, 768, 768), (4096, 3072, 768), (4096, 768, 3072), (8192, 8192, 1024), (8192, 8192, 8192),
                      (1280, 64, 576), (20, 1000, 512)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        b = torch.randn(N, K, device="cuda").bfloat16()
        per_impl = {}
        for name, impl in (("simple", 1), ("persistent128", 2), ("persistent256", 3), ("auto", 0)):
            ext.gemm_set_impl(impl)
            per_impl[name], _ = timeit(lambda: ext.gemm_bf16_tn(a, b, None, False, False))
        ext.gemm_set_impl(0)
        ours = per_impl["auto"]
        lib, _ = timeit(lambda: torch.matmul(a, b.t()))
        fl = 2.0 * M * N * K
        rows.append({"M": M, "N": N, "K": K, "ours_ms": ours, "cublas_ms": lib, "ours_tflops": fl / ours / 1e9,
                     "cublas_tflops": fl / lib / 1e9, "frac_of_measured_peak": fl / ours / 1e9 / PEAKS["bf16_tflops"],
                     "tflops_by_impl": {k: fl / v / 1e9 for k, v in per_impl.items()}})
        print(rows[-1])
    out["gemm_tcgen05"] = rows


def bench_arena(out):
    rows = []
    for S, P in [(10, 11_699_136), (1, 11_699_136), (16, 1_206_592)]:
        w, g = torch.randn(S, P, device="cuda"), torch.randn(S, P, device="cuda")
        hyper = arena_ops.make_hyper(S, "cuda", lr=0.1, max_norm=5.0)
        stats = torch.zeros(S, 4, device="cuda")
        ms, _ = timeit(lambda: arena_ops.fused_client_step(w, g, hyper, stats, n_logical=P, zero_grad=True))
        bytes_ = S * P * 4 * 5        # read g (reduce) + read g, read w, write w, write g (update)
        rows.append({"kernel": "fused_client_step", "S": S, "P": P, "ms": ms, "GBps": bytes_ / ms / 1e6,
                     "frac_of_measured_hbm": bytes_ / ms / 1e6 / PEAKS["hbm_gbs"]})
        acc, wg, wts = torch.zeros(P, device="cuda"), torch.randn(P, device="cuda"), torch.rand(S, device="cuda")
        ms, _ = timeit(lambda: arena_ops.accumulate_pseudo_grad(acc, wg, w, wts))
        bytes_ = (S + 3) * P * 4
        rows.append({"kernel": "accumulate_pseudo_grad", "S": S, "P": P, "ms": ms, "GBps": bytes_ / ms / 1e6,
                     "frac_of_measured_hbm": bytes_ / ms / 1e6 / PEAKS["hbm_gbs"]})
        print(rows[-2], rows[-1])
    P = 11_699_136
    for kind in ("sgd", "adam"):
        st = arena_ops.ServerOptState(kind, P, "cuda", lr=1.0)
        w = torch.randn(P, device="cuda")
        accs = [torch.randn(P, device="cuda") for _ in range(1)]
        ws = torch.tensor(10.0, device="cuda")
        ms, _ = timeit(lambda: arena_ops.server_update(w, accs, ws, st, zero_accs=False))
        nbuf = 3 + (4 if kind == "adam" else 0)          # acc read, w read, w write (+ m,v read+write)
        rows.append({"kernel": "server_update_" + kind, "P": P, "ms": ms, "GBps": nbuf * P * 4 / ms / 1e6,
                     "frac_of_measured_hbm": nbuf * P * 4 / ms / 1e6 / PEAKS["hbm_gbs"]})
        print(rows[-1])
    out["arena"] = rows


def bench_conv(out):
    ext = _ext.load(required=True)
    rows = []
    S, B = 10, 20
    for name, (Cin, H, Cout, k, stride, pad) in {"stem7x7": (3, 32, 64, 7, 2, 3), "layer1": (64, 8, 64, 3, 1, 1),
                                                "layer2": (128, 4, 128, 3, 1, 1), "layer3": (256, 2, 256, 3, 1, 1),
                                                "layer4": (512, 1, 512, 3, 1, 1)}.items():
        n = Cout * Cin * k * k
        P = (n + 31) // 32 * 32
        W = torch.randn(S, P, device="cuda") * 0.05
        G = torch.zeros(S, P, device="cuda")
        x = torch.randn(S, B, Cin, H, H, device="cuda")
        y = ext.slot_conv_fprop(x, W, 0, Cout, k, k, stride, pad, False)
        dy = torch.randn_like(y)
        ext.slot_conv_set_impl(1)
        f, _ = timeit(lambda: ext.slot_conv_fprop(x, W, 0, Cout, k, k, stride, pad, False))
        d, _ = timeit(lambda: ext.slot_conv_dgrad(dy, W, 0, Cin, H, H, k, k, stride, pad, False))
        wg, _ = timeit(lambda: ext.slot_conv_wgrad(x, dy, G, 0, k, k, stride, pad, False))
        ext.slot_conv_set_impl(2)
        tf, _ = timeit(lambda: ext.slot_conv_fprop(x, W, 0, Cout, k, k, stride, pad, False))
        td, _ = timeit(lambda: ext.slot_conv_dgrad(dy, W, 0, Cin, H, H, k, k, stride, pad, False))
        tw, _ = timeit(lambda: ext.slot_conv_wgrad(x, dy, G, 0, k, k, stride, pad, False))
        ext.slot_conv_set_impl(0)
        fl = 2.0 * S * B * y.shape[3] * y.shape[4] * Cout * Cin * k * k
        torch.backends.cudnn.allow_tf32 = True
        xs, ws = x[0], W[0, :n].view(Cout, Cin, k, k)
        lib, _ = timeit(lambda: [torch.nn.functional.conv2d(x[s], W[s, :n].view(Cout, Cin, k, k), None, stride, pad) for s in range(S)])
        rows.append({"layer": name, "fma_fprop_ms": f, "fma_dgrad_ms": d, "fma_wgrad_ms": wg,
                     "tcgen05_fprop_ms": tf, "tcgen05_dgrad_ms": td, "tcgen05_wgrad_ms": tw, "dense_gflop": fl / 1e9,
                     "fma_fprop_dense_tflops": fl / f / 1e9, "tcgen05_fprop_dense_tflops": fl / tf / 1e9, "cudnn_10_launches_fprop_ms": lib})
        print(rows[-1])
    out["slot_conv"] = rows


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="all")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    res = {"peaks": PEAKS}
    if a.only in ("all", "gemm"):
        bench_gemm(res)
    if a.only in ("all", "arena"):
        bench_arena(res)
    if a.only in ("all", "conv"):
        bench_conv(res)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "kernels_{}.json".format(a.only)), "w") as f:
        json.dump(res, f, indent=1)
