"""Stage-by-stage GPU diagnostics of the GCRN prior against the CPU oracle."""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pdse_oracle as O  # noqa: E402
from prior_diffuse_b200 import pack as P, weights as W  # noqa: E402
from prior_diffuse_b200.gcrn import GCRNEngine  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(torch.linalg.norm(a - b) / (torch.linalg.norm(b) + 1e-30))


def from_ug(a, Fq):
    B, CC, R, _ = a.shape
    Pp = Fq + 1
    T = (R - 1) // Pp
    v = a[:, :, :T * Pp].float().view(B, CC, T, Pp, 8)[:, :, :, 1:, :]      # rows t*P+1+f
    # note: the guard of frame t is row t*P; values of frame t are rows t*P+1 .. t*P+F
    v = a.float()[:, :, 1:T * Pp + 1].view(B, CC, T, Pp, 8)[:, :, :, :Fq, :]
    return v.permute(0, 1, 4, 2, 3).reshape(B, CC * 8, T, Fq).cpu()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--T", type=int, default=21)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    sd = W.randomize_norm_stats(W.init_state_dict("GCRN", 1234), 4321)
    eng = GCRNEngine(sd, dev)
    torch.manual_seed(9)
    y = torch.randn(a.B, 2, a.T, 161)
    taps = {}
    t0 = time.time()
    ref = O.gcrn_forward(sd, y, taps) / 11.0
    print(f"oracle gcrn: {time.time() - t0:.2f}s", flush=True)
    got = eng.forward(y.to(dev).contiguous())
    torch.cuda.synchronize()
    ws = eng.workspace(a.B, a.T)
    for i in range(1, 5):
        print(f"enc{i} (ELU'd skip) rel={rel(from_ug(ws[f'e{i}_ug'], P.GCRN_F[i]), F.elu(taps['enc'][i - 1])):.3e}", flush=True)
    print(f"enc5 rel={rel(from_ug(ws['e5_ug'], 4), taps['enc'][4]):.3e}")
    print(f"glstm rel={rel(from_ug(ws['lstm_ug'], 4), taps['glstm']):.3e}")
    # decoder intermediates
    e = taps["enc"]
    out = torch.cat((taps["glstm"], e[4]), 1)
    for br in (1, 2):
        d = out
        for i in range(5, 1, -1):
            op = (0, 1) if i == 2 else (0, 0)
            d = O._bn(O._glu_convT(sd, f"conv{i}_t_{br}", d, op), sd, f"bn{i}_t_{br}")
            print(f"  dec{br}_{i} rel={rel(from_ug(ws[f'd{br}_{i}'], d.shape[-1]), F.elu(d)):.3e}", flush=True)
            d = F.elu(torch.cat((d, e[i - 2]), 1))
    print(f"x_init rel={rel(got, ref):.3e}")
    print("DIAG DONE")


if __name__ == "__main__":
    main()
