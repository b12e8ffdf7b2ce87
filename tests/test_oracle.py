"""The oracle restatement against the committed golden vectors (made from the reference's own
modules by tests/golden/make_golden.py) and, when /root/reference is present, against a live run."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import pdse_oracle as O
from prior_diffuse_b200 import pipeline as PL
from prior_diffuse_b200 import weights as W

HERE = os.path.dirname(os.path.abspath(__file__))


def seeded(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def weights(name):
    return W.randomize_norm_stats(W.init_state_dict(name, seed=1234), seed=4321)


def test_schedule_matches_reference_constants():
    ref = json.load(open(os.path.join(HERE, "golden", "schedule.json")))
    for fast, tag in ((True, "fast"), (False, "full")):
        for impl in (O.inference_schedule, PL.inference_schedule):
            a, b, ac, s, T = impl(fast)
            assert np.allclose(a, ref[tag]["alpha"], rtol=0, atol=1e-15)
            assert np.allclose(ac, ref[tag]["alpha_cum"], rtol=1e-14)
            assert np.allclose(s, ref[tag]["sigmas"], rtol=1e-12)
            assert np.allclose(T, ref[tag]["T"], rtol=0, atol=1e-6)
    # the reference's own comment at trainer/complex_ddpm_trainer.py:459
    assert np.allclose(np.round(ref["cdiffuse_sigmas"], 3), [0.715, 0.0095, 0.031, 0.096, 0.221, 0.412], atol=1.1e-3)
    assert np.allclose(PL.inference_schedule(False)[4], np.arange(50), atol=1e-5)


def test_param_counts():
    # SURVEY.md 8c / trainer :673-style counts of the reference modules
    n = {k: sum(v.numel() for kk, v in W.init_state_dict(k).items()
                if not kk.endswith(("running_mean", "running_var", "num_batches_tracked")))
         for k in ("GCRN", "DiffUNet1", "DiffUNet", "aia_complex_trans_ri", "Nocon")}
    # 1 662 565 is the reference's own number: comment at trainer/complex_ddpm_trainer.py:673
    assert n == {"GCRN": 9771340, "DiffUNet1": 2780273, "DiffUNet": 1662565, "aia_complex_trans_ri": 1179030,
                 "Nocon": 2780263}


@pytest.mark.parametrize("tag", ["a", "b"])
def test_gcrn_golden(golden, tag):
    B, T, seed = (int(v) for v in golden[f"gcrn_{tag}_meta"])
    y = O.gcrn_forward(weights("GCRN"), seeded((B, 2, T, 161), seed))
    assert rel(y.numpy(), golden[f"gcrn_{tag}_y"]) < 2e-6


@pytest.mark.parametrize("tag", ["a", "b"])
def test_dbaiat_prior_golden(golden, tag):
    B, T, seed = (int(v) for v in golden[f"dbaiat_{tag}_meta"])
    y = O.dbaiat_forward(weights("aia_complex_trans_ri"), seeded((B, 2, T, 161), seed))
    assert rel(y.numpy(), golden[f"dbaiat_{tag}_y"]) < 5e-6


def test_dbaiat_utterances_are_independent():
    # eval-mode network: GroupNorm / attention / GRU never mix utterances (what makes sharding exact)
    sd = weights("aia_complex_trans_ri")
    x = seeded((2, 2, 11, 161), 9)
    y = O.dbaiat_forward(sd, x)
    y1 = O.dbaiat_forward(sd, x[1:])
    assert rel(y[1:].numpy(), y1.numpy()) < 2e-6


def test_nocon_golden(golden):
    B, T, seed, tval = golden["nocon_a_meta"]
    x = seeded((int(B), 2, int(T), 161), int(seed))
    y = O.diffunet1_forward(weights("Nocon"), x, None, torch.full((int(B),), float(tval)))
    assert rel(y.numpy(), golden["nocon_a_y"]) < 2e-6


def test_diffunet_prior_golden(golden):
    B, T, seed = (int(v) for v in golden["diffunet_a_meta"])
    y = O.diffunet_forward(weights("DiffUNet"), seeded((B, 2, T, 161), seed))
    assert rel(y.numpy(), golden["diffunet_a_y"]) < 2e-6


def test_diffunet_is_diffunet1_with_dead_time_paths():
    # the adapter the kernels use (pack.diffunet_as_diffunet1) is exact in fp32
    from prior_diffuse_b200.pack import diffunet_as_diffunet1
    sd = weights("DiffUNet")
    x = seeded((1, 2, 9, 161), 3)
    a = O.diffunet_forward(sd, x)
    b = O.diffunet1_forward(diffunet_as_diffunet1(sd), x, seeded((1, 2, 9, 161), 4), torch.tensor([17.3]))
    assert rel(b.numpy(), a.numpy()) < 1e-6


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_diffunet1_golden(golden, tag):
    B, T, seed, tval = golden[f"ddpm_{tag}_meta"]
    B, T, seed = int(B), int(T), int(seed)
    x, x0 = seeded((B, 2, T, 161), seed), seeded((B, 2, T, 161), seed + 100, 0.3)
    t = torch.full((B,), int(tval), dtype=torch.int64) if tag == "c" else torch.full((B,), float(tval))
    y = O.diffunet1_forward(weights("DiffUNet1"), x, x0, t)
    assert rel(y.numpy(), golden[f"ddpm_{tag}_y"]) < 2e-6


@pytest.mark.parametrize("tag,mask", [("plain", False), ("sigma", True)])
def test_end_to_end_golden(golden, tag, mask):
    B, L, ws, xs = (int(v) for v in golden["e2e_meta"])
    wav = seeded((B, L), ws, 0.1)
    x_T = seeded((B, 2, 1 + L // 160, 161), xs)
    st = {}
    y = O.enhance(weights("GCRN"), weights("DiffUNet1"), wav, x_T, True, mask, stages=st)
    assert rel(st["x_init"].numpy(), golden[f"e2e_{tag}_xinit"]) < 2e-6
    assert rel(st["spec"].numpy(), golden[f"e2e_{tag}_spec"]) < 5e-6
    assert rel(y.numpy(), golden[f"e2e_{tag}_wav"]) < 5e-6


def test_stft_golden_and_definition(golden):
    B, L, seed = (int(v) for v in golden["stft_meta"])
    w = seeded((B, L), seed, 0.1)
    assert rel(O.stft(w)[0].numpy(), golden["stft_z"]) < 1e-6
    assert rel(O.stft_direct_f64(w[0].numpy()), golden["stft_z"]) < 1e-6
    assert rel(O.stft_compress(w).numpy(), golden["stft_compressed"]) < 1e-6
    assert rel(O.decompress_istft(O.stft_compress(w), L).numpy(), golden["istft_roundtrip"]) < 1e-5
    # compress / decompress are inverses; ISTFT(STFT(x)) = x
    assert rel(O.decompress_sqrt(O.compress_sqrt(O.stft(w))).numpy(), O.stft(w).numpy()) < 1e-5
    assert rel(O.istft(O.stft(w), L).numpy(), w.numpy()) < 1e-5


def test_segmental_snr_golden(golden):
    # utils/metrics.py:36-55 run by make_golden.py on the same seeded signals
    for (n, seed, noise), ref in zip(golden["ssnr_cases"], golden["ssnr_vals"]):
        n, seed = int(n), int(seed)
        c = seeded((n,), seed, 0.1).double().numpy()
        c[n // 3:n // 2] *= 0.01
        p = c + noise * 0.1 * seeded((n,), seed + 50).double().numpy()
        assert abs(O.snr_seg(c, p) - ref) < 1e-9


def test_reverse_loop_is_deterministic_given_xT():
    # newsigma == 0 (trainer :986-992): two runs with the same x_T agree bit for bit
    sd = weights("DiffUNet1")
    x0, xT = seeded((1, 2, 5, 161), 1, 0.1), seeded((1, 2, 5, 161), 2)
    a = O.reverse_loop(sd, x0, xT, True)
    b = O.reverse_loop(sd, x0, xT, True)
    assert torch.equal(a, b)


@pytest.mark.skipif(not os.path.isdir("/root/reference/model"), reason="reference checkout not present")
def test_oracle_against_live_reference():
    report = json.load(open(os.path.join(HERE, "golden", "oracle_vs_reference.json")))
    assert all(v < 2e-5 for v in report.values() if isinstance(v, float))


# ------------------------------------------------------------------ diff2.DiffWave
@pytest.fixture(scope="module")
def golden_dw():
    return np.load(os.path.join(HERE, "golden", "golden_diffwave.npz"))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_diffwave_golden(golden_dw, tag):
    from tests.golden.make_golden_diffwave import diffwave_weights
    B, L, seed = (int(v) for v in golden_dw[f"diffwave_{tag}_meta"])
    t = torch.from_numpy(golden_dw[f"diffwave_{tag}_t"])
    y = O.diffwave_forward(diffwave_weights(), seeded((B, L), seed), seeded((B, L), seed + 100, 0.5), t, W.DIFFWAVE_CYCLE)
    assert y.shape == (B, 1, L) and rel(y.numpy(), golden_dw[f"diffwave_{tag}_y"]) < 2e-6


def test_diffwave_oracle_pinned_against_live_reference():
    report = json.load(open(os.path.join(HERE, "golden", "oracle_vs_reference_diffwave.json")))
    assert set(report) == {"a", "b", "c"} and max(report.values()) < 2e-6
