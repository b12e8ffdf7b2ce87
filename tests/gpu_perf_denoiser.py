"""Per-kernel timing of one denoiser evaluation at a bench-sized shape (CUDA events)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prior_diffuse_b200 import lib as plib, pack as P, weights as W  # noqa: E402
from prior_diffuse_b200.denoiser import DenoiserEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--T", type=int, default=301)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    sd = W.init_state_dict("DiffUNet1", 1234)
    eng = DenoiserEngine(sd, dev)
    x = torch.randn(a.B, 2, a.T, 161, device=dev)
    x0 = 0.1 * torch.randn(a.B, 2, a.T, 161, device=dev)
    rows = eng.time_bias(torch.tensor([4.0866]))
    for upto, name in (("enc", "encoder"), ("tcm", "enc+tcm"), (None, "full")):
        for _ in range(3):
            eng.forward(x, x0, rows, 0, upto=upto)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            eng.forward(x, x0, rows, 0, upto=upto)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        print(f"{name}: {ms:.3f} ms", flush=True)
    flop = a.B * (25635600 * a.T + 6580480)
    print(f"B={a.B} T={a.T}: {flop / 1e12:.3f} TFLOP/step -> {flop / (ms * 1e-3) / 1e12:.1f} TFLOP/s algorithmic")


if __name__ == "__main__":
    main()
