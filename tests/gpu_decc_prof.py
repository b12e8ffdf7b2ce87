"""Per-phase cycle breakdown of the split decoder's conv kernel (debug hook): consumer 0 and its MMA issuer, CTA (0,0)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prior_diffuse_b200 import lib as plib, weights as W
from prior_diffuse_b200.denoiser import DenoiserEngine

dev = torch.device("cuda:0")
eng = DenoiserEngine(W.init_state_dict("DiffUNet1", 1234), dev)
eng.dec_split = True
B, T = 64, 301
x = torch.randn(B, 2, T, 161, device=dev)
x0 = 0.1 * torch.randn(B, 2, T, 161, device=dev)
rows = eng.time_bias(torch.tensor([4.0866]))
for _ in range(2):
    eng.forward(x, x0, rows, 0)
torch.cuda.synchronize()
L = plib.load()
prof = torch.zeros(12 * 5, dtype=torch.int64, device=dev)
orig = L.pdse_dec_fwd
k = [0]
def hook(*a):
    L.pdse_debug_dec_prof(plib.C.c_void_p(prof.data_ptr() + 96 * k[0]))
    k[0] += 1
    return orig(*a)
L.pdse_dec_fwd = hook
eng.forward(x, x0, rows, 0)
torch.cuda.synchronize()
L.pdse_dec_fwd = orig
L.pdse_debug_dec_prof(None)
for blk, p in zip(["dec5", "dec4", "dec3", "dec2", "dec1"], prof.view(5, 12).tolist()):
    n = max(p[6], 1)
    print(f"{blk}: {n} items in CTA(0,0); cycles per item: consumer: wait acc {p[0] / n:.0f}  gate {p[1] / n:.0f}  wait out-GEMM {p[2] / n:.0f}  "
          f"store {p[3] / n:.0f} | issuer: wait go {p[4] / n:.0f}  issue {p[5] / n:.0f}")
