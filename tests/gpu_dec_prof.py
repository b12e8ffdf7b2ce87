"""Per-phase cycle breakdown of the decoder kernel (debug hook): producer and first consumer warpgroup of CTA (0,0)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prior_diffuse_b200 import lib as plib, weights as W
from prior_diffuse_b200.denoiser import DenoiserEngine

dev = torch.device("cuda:0")
eng = DenoiserEngine(W.init_state_dict("DiffUNet1", 1234), dev)
B, T = 64, 301
x = torch.randn(B, 2, T, 161, device=dev)
x0 = 0.1 * torch.randn(B, 2, T, 161, device=dev)
rows = eng.time_bias(torch.tensor([4.0866]))
for _ in range(2):
    eng.forward(x, x0, rows, 0)
torch.cuda.synchronize()
L = plib.load()
prof = torch.zeros(12 * 5, dtype=torch.int64, device=dev)
# one buffer slice per decoder block: run the blocks one at a time through the timing hook
eng.timing = []
names_p = ["wait x", "wait skip", "GEMM1 x", "GEMM1 skip", "wait h_empty", "scatter"]
names_c = ["wait h_full", "MMA issue", "MMA done", "tail+store"]
orig = L.pdse_dec_fwd
k = [0]
class Hook:
    def __call__(self, *a):
        L.pdse_debug_dec_prof(plib.C.c_void_p(prof.data_ptr() + 96 * k[0]))
        k[0] += 1
        return orig(*a)
L.pdse_dec_fwd = Hook()
eng.forward(x, x0, rows, 0)
torch.cuda.synchronize()
L.pdse_dec_fwd = orig
L.pdse_debug_dec_prof(None)
pr = prof.view(5, 12).tolist()
tm = {n: e0.elapsed_time(e1) for n, e0, e1 in eng.timing}
for i, blk in enumerate(["dec5", "dec4", "dec3", "dec2", "dec1"]):
    p = pr[i]
    tiles = max(p[6], 1)
    print(f"{blk}: {tm.get(blk, 0):.3f} ms, {tiles} tiles in CTA(0,0); cycles per tile:")
    print("   producer: " + "  ".join(f"{n} {v / tiles:.0f}" for n, v in zip(names_p, p[:6])) + f"  | sum {sum(p[:6]) / tiles:.0f}")
    print("   consumer: " + "  ".join(f"{n} {v / tiles:.0f}" for n, v in zip(names_c, p[8:12])) + f"  | sum {sum(p[8:12]) / tiles:.0f}")

# ---- persistent TCM kernel
tp = torch.zeros(12, dtype=torch.int64, device=dev)
L.pdse_debug_tcm_prof(plib.ptr(tp))
eng.timing = []
eng.forward(x, x0, rows, 0, upto="tcm")
torch.cuda.synchronize()
L.pdse_debug_tcm_prof(None)
v = tp.tolist()
n = max(v[11], 1)
names_t = ["xold issue+zero", "wait loads", "conv GEMM", "gate epi", "wait w3/w1", "64->256 GEMM", "residual epi", "256->64 GEMM", "out epi", "dep wait", "hand-over"]
tmt = {nm: e0.elapsed_time(e1) for nm, e0, e1 in eng.timing}
print(f"tcm_flow: {tmt.get('tcm_flow', 0):.3f} ms, {n} tasks in CTA 0; cycles per task:")
for nm, c in zip(names_t, v[:11]):
    print(f"   {nm:24s} {c / n:8.0f}")
print(f"   sum {sum(v[:11]) / n:.0f}")
