"""One eager pass of the bench workload (no CUDA graph) -- the command profiled with ncu.
   python tests/gpu_perf_pass.py [--B 64] [--passes 2]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from prior_diffuse_b200 import Enhancer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--passes", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g, d = bench.seeded_weights()
    enh = Enhancer(g, d, dev, use_graph=False)
    wav = bench.synthetic_wav(a.B).to(dev)
    for _ in range(a.passes):
        out = enh.enhance(wav)
    torch.cuda.synchronize()
    print("ok", float(out.abs().mean()))


if __name__ == "__main__":
    main()
