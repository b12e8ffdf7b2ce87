"""per-kernel timing of one diff2.DiffWave evaluation (64 x 3 s)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prior_diffuse_b200 import weights as W
from prior_diffuse_b200.diffwave import DiffWaveEngine

B, L = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 48000
dev = torch.device("cuda:0")
sd = W.init_state_dict("DiffWave", 1)
sd["output_projection.weight"] = torch.randn_like(sd["output_projection.weight"]) * 0.2
eng = DiffWaveEngine(sd, dev)
a, c, t = torch.randn(B, L, device=dev), torch.randn(B, L, device=dev) * 0.5, torch.full((B,), 7.0, device=dev)
for _ in range(3):
    eng.forward(a, c, t)
torch.cuda.synchronize()
eng.timing = []
eng.forward(a, c, t)
torch.cuda.synchronize()
tot = {}
per = []
for name, e0, e1 in eng.timing:
    ms = e0.elapsed_time(e1)
    tot[name] = tot.get(name, 0.0) + ms
    if name == "dw_layer":
        per.append(round(ms, 3))
eng.timing = None
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    eng.forward(a, c, t)
e1.record()
torch.cuda.synchronize()
rows = B * L
print({k: round(v, 3) for k, v in tot.items()}, "layers:", per)
ms = e0.elapsed_time(e1) / 5
print(f"forward {ms:.3f} ms; per layer {tot['dw_layer'] / len(per):.3f} ms; algorithmic bytes/row/layer = 1280 (x rw 512 + skip rw 512 + y in/out 256)"
      f" -> {rows * 1280 / (tot['dw_layer'] / len(per) * 1e-3) / 1e9:.0f} GB/s; {rows * 2 * 128 * (384 + 64) / (tot['dw_layer'] / len(per) * 1e-3) / 1e12:.1f} TFLOP/s")
