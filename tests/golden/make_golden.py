"""Generate the committed golden fixtures from the REAL reference modules.

Runs only in the build container (needs ``/root/reference``).  It
  1. imports the reference ``nn.Module``s (model/gcrn.py, model/diff3.py) with the
     shims SURVEY.md 8c lists (ptflops stub, params by path, CUDA_VISIBLE_DEVICES restored),
  2. loads OUR seeded weight tables into them with ``strict=True`` (pins key names/shapes),
  3. runs them on small seeded inputs and writes inputs-by-seed + outputs to
     ``tests/golden/*.npz`` / ``*.json``,
  4. checks the oracle restatement (oracle/pdse_oracle.py) against the same outputs.

    python tests/golden/make_golden.py
"""
import contextlib
import importlib.util
import io
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import pdse_oracle as O  # noqa: E402
from prior_diffuse_b200 import weights as W  # noqa: E402


def import_reference():
    cvd = os.environ.get("CUDA_VISIBLE_DEVICES")
    sys.path.insert(0, REF)
    stub = types.ModuleType("ptflops")
    stub.get_model_complexity_info = lambda *a, **k: (0, 0)
    sys.modules["ptflops"] = stub
    import model.gcrn as gcrn
    import model.diff3 as diff3
    import model.diff as diffm
    import model.dbaiat as dbaiat
    import model.piror_grad as nocon
    spec = importlib.util.spec_from_file_location("ref_params", os.path.join(REF, "utils/params.py"))
    pm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pm)
    if cvd is None:
        os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    else:
        os.environ["CUDA_VISIBLE_DEVICES"] = cvd
    # utils/metrics.py by path; its pesq / pystoi / librosa imports are absent here and unused by SNRseg
    for name in ("pesq", "pystoi", "pystoi.stoi", "librosa", "soundfile"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pesq"].pesq = None
    sys.modules["pesq"].PesqError = Exception
    sys.modules["pystoi.stoi"].stoi = None
    spec = importlib.util.spec_from_file_location("ref_metrics", os.path.join(REF, "utils/metrics.py"))
    mt = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mt)
    except Exception as e:      # pragma: no cover
        print("utils/metrics.py not importable:", e)
        mt = None
    return gcrn, diff3, pm.params, diffm, dbaiat, mt, nocon


def seeded_weights(name):
    return W.randomize_norm_stats(W.init_state_dict(name, seed=1234), seed=4321)


def seeded(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


def rel(a, b):
    return float(torch.linalg.norm((a - b).double()) / (torch.linalg.norm(b.double()) + 1e-30))


def main():
    torch.set_grad_enabled(False)
    torch.set_num_threads(8)
    gcrn_mod, diff3_mod, params, diff_mod, dbaiat_mod, metrics_mod, nocon_mod = import_reference()
    sd_g = seeded_weights("GCRN")
    sd_d = seeded_weights("DiffUNet1")
    with contextlib.redirect_stdout(io.StringIO()):
        ref_g = gcrn_mod.GCRN().eval()
        ref_d = diff3_mod.DiffUNet1(params).eval()
    with contextlib.redirect_stdout(io.StringIO()):
        ref_u = diff_mod.DiffUNet().eval()
        ref_a = dbaiat_mod.aia_complex_trans_ri().eval()
        ref_n = nocon_mod.Nocon(params).eval()
    sd_u = seeded_weights("DiffUNet")
    sd_a = seeded_weights("aia_complex_trans_ri")
    sd_n = seeded_weights("Nocon")
    keys = {
        "DiffUNet": [[k, list(v.shape), str(v.dtype)] for k, v in ref_u.state_dict().items()],
        "GCRN": [[k, list(v.shape), str(v.dtype)] for k, v in ref_g.state_dict().items()],
        "aia_complex_trans_ri": [[k, list(v.shape), str(v.dtype)] for k, v in ref_a.state_dict().items()],
        "Nocon": [[k, list(v.shape), str(v.dtype)] for k, v in ref_n.state_dict().items()],
        "DiffUNet1": [[k, list(v.shape), str(v.dtype)] for k, v in ref_d.state_dict().items()],
    }
    json.dump(keys, open(os.path.join(HERE, "state_dict_keys.json"), "w"))
    ref_u.load_state_dict(sd_u, strict=True)
    ref_a.load_state_dict(sd_a, strict=True)
    ref_n.load_state_dict(sd_n, strict=True)
    ref_g.load_state_dict(sd_g, strict=True)
    ref_d.load_state_dict(sd_d, strict=True)
    n_param = {"DiffUNet": sum(p.numel() for p in ref_u.parameters()),
               "aia_complex_trans_ri": sum(p.numel() for p in ref_a.parameters()),
               "Nocon": sum(p.numel() for p in ref_n.parameters()),
               "GCRN": sum(p.numel() for p in ref_g.parameters()),
               "DiffUNet1": sum(p.numel() for p in ref_d.parameters())}
    print("params", n_param)   # 9 771 340 / 2 780 273

    report = {}
    # ---- schedule (reference constants; trainer :105-156 restated, pinned on the :459 comment)
    sched = {}
    for fast in (True, False):
        a, b, ac, s, T = O.inference_schedule(fast, params.noise_schedule, params.inference_noise_schedule)
        sched["fast" if fast else "full"] = dict(alpha=list(map(float, a)), beta=list(map(float, b)),
                                                 alpha_cum=list(map(float, ac)), sigmas=list(map(float, s)),
                                                 T=list(map(float, T)))
    _, _, _, s_cd, _ = O.inference_schedule(True, np.linspace(1e-4, 0.035, 50).tolist(),
                                            [0.0001, 0.001, 0.01, 0.05, 0.2, 0.35])
    sched["cdiffuse_sigmas"] = list(map(float, s_cd))   # comment at trainer :459
    json.dump(sched, open(os.path.join(HERE, "schedule.json"), "w"), indent=1)

    # ---- GCRN
    out = {}
    for tag, (B, T, seed) in {"a": (2, 24, 11), "b": (1, 67, 12)}.items():
        x = seeded((B, 2, T, 161), seed)
        y_ref = ref_g(x.clone())
        y_or = O.gcrn_forward(sd_g, x)
        report[f"gcrn_{tag}"] = rel(y_or, y_ref)
        out[f"gcrn_{tag}_meta"] = np.array([B, T, seed])
        out[f"gcrn_{tag}_y"] = y_ref.numpy()
    # ---- DiffUNet prior (model/diff.py; 1 662 565 parameters, trainer :673)
    for tag, (B, T, seed) in {"a": (2, 19, 51)}.items():
        x = seeded((B, 2, T, 161), seed)
        y_ref = ref_u(x)
        report[f"diffunet_{tag}"] = rel(O.diffunet_forward(sd_u, x), y_ref)
        out[f"diffunet_{tag}_meta"] = np.array([B, T, seed])
        out[f"diffunet_{tag}_y"] = y_ref.numpy()
    # ---- aia_complex_trans_ri prior (model/dbaiat.py; 1 179 030 parameters)
    for tag, (B, T, seed) in {"a": (2, 23, 61), "b": (1, 40, 62)}.items():
        x = seeded((B, 2, T, 161), seed)
        y_ref = ref_a(x)
        report[f"dbaiat_{tag}"] = rel(O.dbaiat_forward(sd_a, x), y_ref)
        out[f"dbaiat_{tag}_meta"] = np.array([B, T, seed])
        out[f"dbaiat_{tag}_y"] = y_ref.numpy()
    # ---- Nocon (model/piror_grad.py), the deltamu denoiser
    for tag, (B, T, seed, tval) in {"a": (2, 21, 81, 10.451817)}.items():
        x = seeded((B, 2, T, 161), seed)
        t = torch.full((B,), tval, dtype=torch.float32)
        y_ref = ref_n(x, t)
        report[f"nocon_{tag}"] = rel(O.diffunet1_forward(sd_n, x, None, t), y_ref)
        out[f"nocon_{tag}_meta"] = np.array([B, T, seed, float(tval)])
        out[f"nocon_{tag}_y"] = y_ref.numpy()
    # ---- DiffUNet1 (float t = fast schedule, int t = full schedule)
    for tag, (B, T, seed, tval) in {"a": (2, 24, 21, 4.086654), "b": (1, 100, 22, 42.918644),
                                    "c": (2, 17, 23, 7)}.items():
        x = seeded((B, 2, T, 161), seed)
        x0 = seeded((B, 2, T, 161), seed + 100, 0.3)
        t = (torch.full((B,), tval, dtype=torch.float32) if isinstance(tval, float)
             else torch.full((B,), tval, dtype=torch.int64))
        y_ref = ref_d(x, x0, t)
        y_or = O.diffunet1_forward(sd_d, x, x0, t)
        report[f"ddpm_{tag}"] = rel(y_or, y_ref)
        out[f"ddpm_{tag}_meta"] = np.array([B, T, seed, float(tval)])
        out[f"ddpm_{tag}_y"] = y_ref.numpy()
    # ---- whole path with the reference modules driving the restated loop
    B, L = 2, 3200
    wav = seeded((B, L), 31, 0.1)
    T = 1 + L // 160
    x_T = seeded((B, 2, T, 161), 7)
    for tag, mask in (("plain", False), ("sigma", True)):
        w, c = O.rms_normalize(wav)
        feat = O.stft_compress(w)
        x_init = ref_g(feat.clone()) / 11.0
        alpha, beta, alpha_cum, sigmas, Tn = O.inference_schedule(True)
        audio = x_T.clone()
        if mask:
            audio = audio * (O.sigma_mask(x_init) ** 0.5)
        for n in range(len(alpha) - 1, -1, -1):
            c1 = 1 / alpha[n] ** 0.5
            c2 = beta[n] / (1 - alpha_cum[n]) ** 0.5
            eps = ref_d(audio, x_init, torch.tensor([Tn[n]]).repeat(B))
            audio = c1 * (audio - c2 * eps)
        spec = (audio + x_init) * 11.0
        y = O.decompress_istft(spec, L) * c
        st = {}
        y_or = O.enhance(sd_g, sd_d, wav, x_T, True, mask, stages=st)
        report[f"e2e_{tag}_wav"] = rel(y_or, y)
        report[f"e2e_{tag}_spec"] = rel(st["spec"], spec)
        out[f"e2e_{tag}_spec"] = spec.numpy()
        out[f"e2e_{tag}_wav"] = y.numpy()
        out[f"e2e_{tag}_xinit"] = x_init.numpy()
    out["e2e_meta"] = np.array([B, L, 31, 7])
    # ---- segmental SNR (utils/metrics.py:36-55) on seeded clean / degraded pairs
    if metrics_mod is not None:
        vals = []
        cases = [(16000, 71, 0.3), (48000, 72, 0.05), (5000, 73, 1.5), (1333, 74, 0.0)]
        for n, seed, noise in cases:
            c = seeded((n,), seed, 0.1).double().numpy()
            c[n // 3:n // 2] *= 0.01                       # a quiet stretch: exercises the -10 dB clamp
            p_ = c + noise * 0.1 * seeded((n,), seed + 50).double().numpy()
            ref_v = float(metrics_mod.SNRseg(c, p_, 16000))
            vals.append(ref_v)
            report[f"ssnr_{n}"] = abs(O.snr_seg(c, p_) - ref_v)
        out["ssnr_cases"] = np.array(cases, dtype=np.float64)
        out["ssnr_vals"] = np.array(vals)
    # ---- STFT: torch.stft (what the reference calls) vs the written-out DFT
    w1 = seeded((1, 1600), 41, 0.1)
    z = O.stft(w1)[0].numpy()
    zd = O.stft_direct_f64(w1[0].numpy())
    report["stft_direct_vs_torch"] = float(np.linalg.norm(z - zd) / np.linalg.norm(zd))
    out["stft_meta"] = np.array([1, 1600, 41])
    out["stft_z"] = z
    out["stft_compressed"] = O.stft_compress(w1).numpy()
    out["istft_roundtrip"] = O.decompress_istft(O.stft_compress(w1), 1600).numpy()

    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    report["n_param"] = n_param
    json.dump(report, open(os.path.join(HERE, "oracle_vs_reference.json"), "w"), indent=1)
    print(json.dumps(report, indent=1))
    bad = {k: v for k, v in report.items() if isinstance(v, float) and v > 2e-5}   # networks: rel-L2; ssnr_*: |dB difference|
    assert not bad, bad


if __name__ == "__main__":
    main()
