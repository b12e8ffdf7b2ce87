"""Fused vs split decoder path: equality of eps and per-block timings (CUDA events, eager launches)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prior_diffuse_b200 import weights as W
from prior_diffuse_b200.denoiser import DenoiserEngine

dev = torch.device("cuda:0")
sd = W.randomize_norm_stats(W.init_state_dict("DiffUNet1", 1234), 4321)
eng = DenoiserEngine(sd, dev)
for B, T in (((2, 37),) if os.environ.get("SMALL") else ((2, 37), (64, 301))):
    x = torch.randn(B, 2, T, 161, device=dev)
    x0 = 0.1 * torch.randn(B, 2, T, 161, device=dev)
    rows = eng.time_bias(torch.tensor([4.0866]))
    res = {}
    for split in (False, True):
        eng.dec_split = split
        for _ in range(2):
            eng.forward(x, x0, rows, 0)
        torch.cuda.synchronize()
        eng.timing = []
        for _ in range(3):
            res[split] = eng.forward(x, x0, rows, 0).clone()
        torch.cuda.synchronize()
        tm = {}
        for n, e0, e1 in eng.timing:
            tm[n] = tm.get(n, 0.0) + e0.elapsed_time(e1) / 3
        eng.timing = None
        print(f"B={B} T={T} split={split}: " + "  ".join(f"{k} {v:.3f}" for k, v in tm.items() if k.startswith("dec")) +
              f"  | dec total {sum(v for k, v in tm.items() if k.startswith('dec')):.3f} ms  all {sum(tm.values()):.3f} ms")
    d = (res[True] - res[False]).abs().max().item()
    print(f"   max |split - fused| = {d:.3e}  equal={torch.equal(res[True], res[False])}  finite={torch.isfinite(res[True]).all().item()}")
