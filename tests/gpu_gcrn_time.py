"""Per-launch times of one GCRN evaluation at the bench shape (64 x 301).   python tests/gpu_gcrn_time.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prior_diffuse_b200 import weights as W  # noqa: E402
from prior_diffuse_b200.gcrn import GCRNEngine  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    eng = GCRNEngine(W.init_state_dict("GCRN", 1234), dev)
    y = torch.randn(64, 2, 301, 161, device=dev)
    for _ in range(2):
        eng.forward(y)
    torch.cuda.synchronize()
    eng.timing = []
    eng.forward(y)
    torch.cuda.synchronize()
    tot = 0.0
    for name, a, b in eng.timing:
        ms = a.elapsed_time(b)
        tot += ms
        print(f"{name:16s} {ms:.3f} ms")
    print(f"total {tot:.3f} ms")


if __name__ == "__main__":
    main()
