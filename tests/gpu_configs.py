"""SURVEY.md 8(d) configurations other than the bench line, measured on one B200 (per-GPU share of the
multi-GPU configs) and written to gpurun_out/configs.json:
   cfg3  aia_complex_trans_ri prior + DiffUNet1, full 50-step schedule, 32 x 3 s per GPU
   cfg4  GCRN prior + DiffUNet1, 6 steps, sigma mask, 64 x 10 s per GPU (the 2-GPU share of B = 128)
   cfg5  microbench sweep: STFT+compress, decompress+ISTFT, one denoiser step + update
   eager the oracle (the reference's algorithm in PyTorch eager) on the SAME B200, fp32 (TF32 convs), cfg2 shape
python tests/gpu_configs.py [--skip-eager]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prior_diffuse_b200 import Enhancer, signal as S, weights as W  # noqa: E402
from prior_diffuse_b200 import lib as plib  # noqa: E402
from prior_diffuse_b200.denoiser import DenoiserEngine  # noqa: E402


def weights(name):
    return W.randomize_norm_stats(W.init_state_dict(name, 1234), 4321)


def timed(fn, reps, flush):
    fn()
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return sum(ms) / len(ms)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-eager", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    plib.load(require_device=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    g = torch.Generator().manual_seed(1234)
    want = lambda k: not a.only or k in a.only.split(",")

    if want("cfg3"):
        B, L = 32, 48000
        enh = Enhancer(weights("aia_complex_trans_ri"), weights("DiffUNet1"), dev, fast_sampling=False,
                       prior="aia_complex_trans_ri")
        wav = (0.1 * torch.randn(B, L, generator=g)).to(dev)
        ms = timed(lambda: enh.enhance(wav), 5, flush)
        out["cfg3"] = {"workload": "aia_complex_trans_ri + DiffUNet1, 50 reverse steps, 32 x 3 s per GPU (B = 256 on 8 GPUs)",
                       "ms_per_pass": ms, "audio_s_per_s": B * L / 16000 / (ms * 1e-3), "kernels_per_pass": enh.kernels_per_call}
        print("cfg3", out["cfg3"], flush=True)
        del enh
        torch.cuda.empty_cache()
    if want("cfg4"):
        B, L = 64, 160000
        enh = Enhancer(weights("GCRN"), weights("DiffUNet1"), dev, fast_sampling=True, sigma_mask=True)
        wav = (0.1 * torch.randn(B, L, generator=g)).to(dev)
        ms = timed(lambda: enh.enhance(wav), 5, flush)
        out["cfg4"] = {"workload": "GCRN + DiffUNet1, 6 steps, sigma mask, 64 x 10 s per GPU (B = 128 on 2 GPUs)",
                       "ms_per_pass": ms, "audio_s_per_s": B * L / 16000 / (ms * 1e-3)}
        print("cfg4", out["cfg4"], flush=True)
        del enh
        torch.cuda.empty_cache()
    if want("cfg5"):
        rows = []
        eng = DenoiserEngine(weights("DiffUNet1"), dev)
        Lh = plib.load()
        for B in (1, 4, 16, 64, 256, 1024):
            for sec in (1, 2, 3, 5, 10, 30):
                if B * sec > 3072:
                    continue
                L = 16000 * sec
                T = 1 + L // 160
                wav = (0.1 * torch.randn(B, L, generator=g)).to(dev)
                spec = S.stft_compress(wav)
                t_stft = timed(lambda: S.stft_compress(wav, out=spec), 5, flush)
                wout = torch.empty_like(wav)
                t_istft = timed(lambda: S.decompress_istft(spec, L, out=wout), 5, flush)
                x = torch.randn(B, 2, T, 161, device=dev)
                x0 = 0.3 * torch.randn(B, 2, T, 161, device=dev)
                rowsb = eng.time_bias(torch.tensor([4.086654]))
                nel = x.numel()

                def step():
                    eps = eng.forward(x, x0, rowsb, 0)
                    plib.check(Lh.pdse_ddpm_update_f32(plib.ptr(x), plib.ptr(eps), plib.ptr(x0), None, None, nel, T * 161,
                                                       1.0, 0.0, 0.0, 0, 0, 11.0, 0, 0, plib.stream_ptr()))
                t_step = timed(step, 5, flush)
                frames = B * T
                rows.append({"B": B, "seconds": sec, "frames": frames,
                             "stft_ms": t_stft, "stft_GBs": frames * 1928 / (t_stft * 1e-3) / 1e9,
                             "istft_ms": t_istft, "istft_GBs": frames * 1928 / (t_istft * 1e-3) / 1e9,
                             "denoiser_step_ms": t_step,
                             "denoiser_TFLOPs": B * (25635600 * T + 6580480) / (t_step * 1e-3) / 1e12})
                print(rows[-1], flush=True)
                eng._ws.clear()
                torch.cuda.empty_cache()
        out["cfg5"] = rows
    if want("eager") and not a.skip_eager:
        from oracle import pdse_oracle as O
        B, L = 8, 48000
        gd = {k: v.to(dev) for k, v in weights("GCRN").items()}
        dd = {k: v.to(dev) for k, v in weights("DiffUNet1").items()}
        wav = (0.1 * torch.randn(B, L, generator=g)).to(dev)
        x_T = torch.randn(B, 2, 301, 161, device=dev)
        with torch.no_grad():
            for _ in range(2):
                O.enhance(gd, dd, wav, x_T, True, False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                O.enhance(gd, dd, wav, x_T, True, False)
            torch.cuda.synchronize()
            sec = (time.perf_counter() - t0) / 3
        out["eager_b200"] = {"workload": "oracle (PyTorch eager, fp32 / TF32 convs) on the B200, GCRN + DiffUNet1 6 steps, 8 x 3 s",
                             "ms_per_pass": sec * 1e3, "audio_s_per_s": B * 3.0 / sec}
        print("eager", out["eager_b200"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
