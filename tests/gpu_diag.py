"""Stage-by-stage GPU diagnostics (run on the B200 box): prints rel-L2 of every kernel stage
against the CPU oracle.  Not a pytest file; used to localise bugs with one gpurun call.

    python tests/gpu_diag.py [--B 2] [--T 40]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pdse_oracle as O  # noqa: E402
from prior_diffuse_b200 import lib as plib, pack as P, signal as S, weights as W  # noqa: E402
from prior_diffuse_b200.denoiser import DenoiserEngine  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(torch.linalg.norm(a - b) / (torch.linalg.norm(b) + 1e-30))


def from_cp8_split(a, F):
    """[B][C/8][T*2Q][8] -> [B,C,T,F]"""
    B, CC, NP, _ = a.shape
    Q = (F + 1) // 2
    T = NP // (2 * Q)
    a = a.float().view(B, CC, T, 2, Q, 8)
    out = torch.zeros(B, CC * 8, T, F)
    for f in range(F):
        out[:, :, :, f] = a[:, :, :, f & 1, f >> 1, :].permute(0, 1, 3, 2).reshape(B, CC * 8, T).cpu()
    return out


def probe(L, dev):
    torch.manual_seed(1)
    for (N, K, shift) in [(64, 64, 0), (64, 64, 1), (64, 64, 5), (32, 32, 3), (256, 64, 8), (64, 320, 37), (128, 128, 2)]:
        rows = 128 + 48
        A = torch.randn(rows, K)
        Bm = torch.randn(N, K)
        A_cp = A.view(rows, K // 8, 8).permute(1, 0, 2).contiguous().to(dev).to(torch.bfloat16)
        B_cp = Bm.view(N, K // 8, 8).permute(1, 0, 2).contiguous().to(dev).to(torch.bfloat16)
        ref = A.to(torch.bfloat16).float()[shift:shift + 128] @ Bm.to(torch.bfloat16).float().T
        for swap in (0,):
            D = torch.zeros(128, N, device=dev)
            plib.check(L.pdse_probe_gemm(plib.ptr(A_cp), plib.ptr(B_cp), plib.ptr(D), rows, N, K, shift, swap,
                                         plib.stream_ptr()))
            torch.cuda.synchronize()
            print(f"probe N={N} K={K} shift={shift} swap={swap}: rel={rel(D, ref):.3e}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--T", type=int, default=40)
    ap.add_argument("--skip-probe", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    L = plib.load(require_device=True)
    print("device", torch.cuda.get_device_name(0), "sms", L.pdse_sm_count(), flush=True)
    if not args.skip_probe:
        probe(L, dev)

    # ---------------- signal
    torch.manual_seed(3)
    for (B, n) in [(1, 1600), (3, 4321), (2, 48000)]:
        wav = 0.1 * torch.randn(B, n)
        ref = O.stft(wav)
        got = S.stft(wav.to(dev))
        print(f"stft B={B} L={n}: rel={rel(got, ref):.3e}")
        w, c = O.rms_normalize(wav)
        r = S.rms(wav.to(dev))
        print(f"  rms rel={rel(r, c.view(-1)):.3e}")
        refc = O.stft_compress(w)
        gotc = S.stft_compress(wav.to(dev), r)
        print(f"  stft+compress rel={rel(gotc, refc):.3e}")
        spec = torch.randn(B, 2, 1 + n // 160, 161)
        refw = O.decompress_istft(spec, n) * c
        gotw = S.decompress_istft(spec.to(dev), n, r)
        print(f"  decompress+istft rel={rel(gotw, refw):.3e}")
        print(f"  istft rel={rel(S.istft(spec.to(dev), n), O.istft(spec, n)):.3e}", flush=True)

    # ---------------- update kernels
    B, T = args.B, args.T
    n = B * 2 * T * 161
    x = torch.randn(B, 2, T, 161)
    e = torch.randn(B, 2, T, 161)
    x0 = 0.1 * torch.randn(B, 2, T, 161)
    npad = (n + 3) // 4 * 4

    def padded(t):
        buf = torch.zeros(npad, device=dev)
        buf[:n] = t.reshape(-1).to(dev)
        return buf

    xd, ed, x0d = padded(x), padded(e), padded(x0)
    plib.check(L.pdse_ddpm_update_f32(plib.ptr(xd), plib.ptr(ed), None, None, None, n, 0, 1.118, 0.4, 0.0, 0, 0, 1.0,
                                      0, 0, plib.stream_ptr()))
    print(f"update rel={rel(xd[:n].view_as(x), 1.118 * (x - 0.4 * e)):.3e}")
    od = torch.zeros(npad, device=dev)
    xd = padded(x)
    plib.check(L.pdse_ddpm_update_f32(plib.ptr(xd), plib.ptr(ed), plib.ptr(x0d), None, plib.ptr(od), n, 0, 1.118, 0.4,
                                      0.0, 0, 1, 11.0, 0, 0, plib.stream_ptr()))
    print(f"update+finalize rel={rel(od[:n].view_as(x), (1.118 * (x - 0.4 * e) + x0) * 11):.3e}")
    am = torch.zeros(B * 2, device=dev)
    plib.check(L.pdse_absmax_f32(plib.ptr(x0d), B * 2, T * 161, plib.ptr(am), plib.stream_ptr()))
    xd = padded(x)
    plib.check(L.pdse_init_state_f32(plib.ptr(xd), plib.ptr(x0d), plib.ptr(am), n, T * 161, 0, 0, 0, plib.stream_ptr()))
    print(f"sigma mask rel={rel(xd[:n].view_as(x), x * O.sigma_mask(x0) ** 0.5):.3e}")
    z = torch.zeros(1 << 22, device=dev)
    plib.check(L.pdse_init_state_f32(plib.ptr(z), None, None, z.numel(), 0, 1, 7, 0, plib.stream_ptr()))
    print(f"philox normal: mean={z.mean().item():.4f} std={z.std().item():.4f} kurt={(z**4).mean().item():.3f}", flush=True)

    # ---------------- denoiser
    sd = W.randomize_norm_stats(W.init_state_dict("DiffUNet1", 1234), 4321)
    eng = DenoiserEngine(sd, dev)
    torch.manual_seed(5)
    x = torch.randn(B, 2, T, 161)
    x0 = 0.3 * torch.randn(B, 2, T, 161)
    t = torch.tensor([4.086654, 22.992493, 0.0, 42.918644][:B] + [7.0] * max(0, B - 4))
    taps = {}
    t0 = time.time()
    ref = O.diffunet1_forward(sd, x, x0, t, taps)
    print(f"oracle denoiser B={B} T={T}: {time.time() - t0:.2f}s")
    rows = eng.time_bias(t)
    torch.cuda.synchronize()
    from tests import emu
    rows_ref = torch.from_numpy(emu.emu_time(P.pack_diffunet1(sd)["time"], t.numpy()))
    print(f"time bias rows rel={rel(rows, rows_ref):.3e}", flush=True)
    xd, x0d = x.to(dev).contiguous(), x0.to(dev).contiguous()
    for upto in ("enc", "tcm", None):
        got = eng.forward(xd, x0d, rows, P.N_BIAS_ROW, upto=upto)
        torch.cuda.synchronize()
        ws = eng.workspace(B, T)
        if upto == "enc":
            for i in range(1, 6):
                print(f"enc{i} rel={rel(from_cp8_split(ws[f'e{i}'], P.ENC_F[i]), taps['skips'][i - 1]):.3e}", flush=True)
        elif upto == "tcm":
            print(f"tcm rel={rel(from_cp8_split(ws['dec_in'], 4), taps['tcm']):.3e}", flush=True)
        else:
            print(f"eps rel={rel(got, ref):.3e}  (re {rel(got[:, 0], ref[:, 0]):.3e} im {rel(got[:, 1], ref[:, 1]):.3e})")
            # decoder intermediates
            for br, name in enumerate(("de_real", "de_imag")):
                h = taps["tcm"]
                temb = taps["temb"]
                for i in range(5, 1, -1):
                    h = O._biconvtransglu(sd, f"{name}.de{i}.0", torch.cat((h, taps["skips"][i - 1]), 1), temb)[:, :, :-1]
                    h = O._prelu(O._bn(h, sd, f"{name}.de{i}.2"), sd[f"{name}.de{i}.3.weight"])
                    print(f"  {name}.de{i} rel={rel(from_cp8_split(ws[f'd{br}_{i}'], h.shape[-1]), h):.3e}", flush=True)
    print("DIAG DONE")


if __name__ == "__main__":
    main()
