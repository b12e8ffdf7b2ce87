"""Per-kernel timing of the DB-AIAT prior at the bench shape (eager launches, CUDA events).
   python tests/gpu_perf_dbaiat.py [--B 64] [--T 301] [--reps 3]"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prior_diffuse_b200 import weights as W  # noqa: E402
from prior_diffuse_b200.dbaiat import DBAIATEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--T", type=int, default=301)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    sd = W.randomize_norm_stats(W.init_state_dict("aia_complex_trans_ri", 1234), 4321)
    eng = DBAIATEngine(sd, dev)
    x = torch.randn(a.B, 2, a.T, 161, device=dev)
    eng.forward(x)
    torch.cuda.synchronize()
    tot = collections.OrderedDict()
    for _ in range(a.reps):
        eng.timing = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.forward(x)
        e1.record()
        torch.cuda.synchronize()
        for name, s, e in eng.timing:
            tot[name] = tot.get(name, 0.0) + s.elapsed_time(e) / a.reps
        whole = e0.elapsed_time(e1)
    eng.timing = None
    for k, v in tot.items():
        print(f"{k:16s} {v:8.3f} ms")
    print(f"sum {sum(tot.values()):.3f} ms   whole pass {whole:.3f} ms   frames {a.B * a.T}")


if __name__ == "__main__":
    main()
