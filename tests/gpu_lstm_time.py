"""Time one LSTM layer recurrence (both groups) at the bench shape; PDSE_LSTM_MODE caps the kernel variant.
   python tests/gpu_lstm_time.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prior_diffuse_b200 import lib as plib, weights as W  # noqa: E402
from prior_diffuse_b200.gcrn import GCRNEngine  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    eng = GCRNEngine(W.init_state_dict("GCRN", 1234), dev)
    B, T = 64, 301
    y = torch.randn(B, 2, T, 161, device=dev)
    eng.forward(y)
    torch.cuda.synchronize()
    eng.timing = []
    eng.forward(y)
    torch.cuda.synchronize()
    for name, a, b in eng.timing:
        if "lstm" in name:
            print(name, f"{a.elapsed_time(b):.3f} ms")
    print("mode cap", os.environ.get("PDSE_LSTM_MODE"))


if __name__ == "__main__":
    main()
