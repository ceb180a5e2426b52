"""Host-side cProfile of flagship rounds (where does the Python time between GPU syncs go?)."""
import cProfile
import io
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import build_flagship


def main():
    job = build_flagship(n_clients_per_round=10, users=500, norm="gn")
    for _ in range(8):
        job.run_round()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        job.run_round()
    torch.cuda.synchronize()
    pr.disable()
    out = io.StringIO()
    st = pstats.Stats(pr, stream=out)
    st.sort_stats("cumulative").print_stats(70)
    st.sort_stats("tottime").print_stats(35)
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/host_profile.txt", "w").write(out.getvalue())
    print(out.getvalue()[:3000])
    from msrflute_b200.utils.async_ckpt import get_checkpointer
    get_checkpointer().close()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
