"""Packed weights + kernel index math (NumPy emulation) vs the oracle, stage by stage."""
import numpy as np
import pytest
import torch

from oracle import pdse_oracle as O
from prior_diffuse_b200 import pack as P
from prior_diffuse_b200 import weights as W
from tests import emu


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


@pytest.fixture(scope="module")
def setup():
    torch.manual_seed(0)
    sd = W.randomize_norm_stats(W.init_state_dict("DiffUNet1", 1234), 4321)
    packed = P.pack_diffunet1(sd)
    B, T = 2, 7
    x = torch.randn(B, 2, T, 161)
    x0 = 0.3 * torch.randn(B, 2, T, 161)
    t = torch.tensor([4.086654, 22.992493])
    taps = {}
    y = O.diffunet1_forward(sd, x, x0, t, taps)
    return sd, packed, x, x0, t, taps, y


def test_time_rows(setup):
    sd, packed, x, x0, t, taps, y = setup
    rows = emu.emu_time(packed["time"], t.numpy())
    temb = taps["temb"]
    tb1 = torch.nn.functional.linear(temb, sd["en.tp1.weight"], sd["en.tp1.bias"]).numpy()
    assert rel(rows[:, 0:2], tb1) < 1e-5
    tb3 = torch.nn.functional.linear(temb, sd["en.tp3.weight"], sd["en.tp3.bias"])
    hb3 = torch.nn.functional.linear(tb3, sd["en.conv3.conv1.weight"][:, :, 0, 0], sd["en.conv3.conv1.bias"]).numpy()
    o = P.bias_off_enc(3)
    assert rel(rows[:, o:o + 32], hb3) < 1e-5


def test_denoiser_emulation(setup):
    sd, packed, x, x0, t, taps, y = setup
    rows = emu.emu_time(packed["time"], t.numpy())
    B, _, T, _ = x.shape
    e = emu.emu_enc1(packed["enc1"], x.numpy().astype(np.float64), x0.numpy().astype(np.float64), rows)
    skips = [e]
    assert rel(emu.from_cp8_split(e, 79), taps["skips"][0].numpy()) < 1e-5
    for i in range(2, 6):
        o = P.bias_off_enc(i)
        nt = {2: 3, 3: 6, 4: 12, 5: 25}[i]
        e = emu.emu_enc(packed[f"enc{i}"], e, P.ENC_F[i - 1], rows[:, o:o + 32], nt)
        skips.append(e)
        assert rel(emu.from_cp8_split(e, P.ENC_F[i]), taps["skips"][i - 1].numpy()) < 1e-5, i
    h = emu.emu_tcm([packed[f"tcm{k}"] for k in range(18)], e, W.TCM_DILATIONS)
    assert rel(emu.from_cp8_split(h, 4), taps["tcm"].numpy()) < 1e-5
    outs = []
    for bi in range(2):
        d = h
        for i in range(5, 0, -1):
            o = P.bias_off_dec(bi, i)
            Fin = P.ENC_F[i]
            kw = 5 if i == 1 else 3
            nt = max(1, 128 // (Fin + (kw - 1) // 2)) if i != 1 else 3
            d = emu.emu_dec(packed[f"dec{bi}_{i}"], d, skips[i - 1], Fin, kw, rows[:, o:o + 32], nt)
        outs.append(d)
    eps = np.stack(outs, axis=1)
    assert rel(eps, y.numpy()) < 1e-5


def test_gcrn_emulation():
    torch.manual_seed(0)
    sd = W.randomize_norm_stats(W.init_state_dict("GCRN", 1234), 4321)
    pk = P.pack_gcrn(sd)
    B, T = 2, 5
    y = torch.randn(B, 2, T, 161)
    taps = {}
    ref = O.gcrn_forward(sd, y, taps) / 11.0
    got = emu.emu_gcrn(pk, y.numpy().astype(np.float64))
    assert rel(got, ref.numpy()) < 1e-5


def test_warp_fft_lane_arithmetic_matches_numpy_fft_and_the_oracle():
    """csrc/signal.cu's one-frame-per-warp real FFT (5 x 32 decomposition, shuffle butterflies, untangling) re-run in
    float32 NumPy: against numpy.fft in float64 and against the oracle's torch.stft framing, inside the 1e-5 bar"""
    import numpy as np
    import torch
    from oracle import pdse_oracle as O
    from tests import emu
    rng = np.random.default_rng(3)
    hann, _ = emu.fft_tables()
    for _ in range(4):
        s = rng.standard_normal(320).astype(np.float32)
        X = emu.stft_frame_lanes(s)
        ref = np.fft.rfft(s.astype(np.float64) * hann.astype(np.float64))
        assert np.linalg.norm(X - ref) / np.linalg.norm(ref) < 2e-6
        Xin = (rng.standard_normal(161) + 1j * rng.standard_normal(161)).astype(np.complex64)
        x = emu.istft_frame_lanes(Xin)
        refx = np.fft.irfft(Xin.astype(np.complex128), 320) * hann.astype(np.float64)
        assert np.linalg.norm(x - refx) / np.linalg.norm(refx) < 2e-6
    # one utterance through the oracle's STFT (center / reflect framing): frame t of the emulation == column t
    wav = torch.from_numpy(rng.standard_normal(1600).astype(np.float32))[None]
    spec = O.stft(wav)[0].numpy()                      # [2, T, 161]
    padded = np.pad(wav[0].numpy(), 160, mode="reflect")
    for t in (0, 3, 10):
        X = emu.stft_frame_lanes(padded[t * 160:t * 160 + 320])
        ref = spec[0, t] + 1j * spec[1, t]
        assert np.linalg.norm(X - ref) / np.linalg.norm(ref) < 1e-5


def test_diffwave_packing_emulation():
    """the packed DiffWave operands, pushed through a NumPy statement of the kernels' arithmetic
    (tests/emu.emu_diffwave), reproduce the oracle -- pins the K order (phase, tap, channel), the bias blocks and the
    guard-row convention on the CPU"""
    import math
    from oracle import pdse_oracle as O
    from prior_diffuse_b200.diffwave import pack_diffwave
    sd = W.init_state_dict("DiffWave", 31)
    sd = {k: v for k, v in sd.items() if not k.startswith("residual_layers.") or int(k.split(".")[1]) < 12}
    sd["output_projection.weight"] = 0.2 * torch.randn(sd["output_projection.weight"].shape, generator=torch.Generator().manual_seed(3))
    L = 1500                                                  # below 2 x the largest dilation (512): the guards matter
    g = torch.Generator().manual_seed(8)
    audio, init = torch.randn(1, L, generator=g), 0.5 * torch.randn(1, L, generator=g)
    t = torch.tensor([17])
    ref = O.diffwave_forward(sd, audio, init, t, 10)[0, 0].double().numpy()
    pk = pack_diffwave(sd, 10)
    table = O.time_embedding_table(50)
    e = table[t]
    for k in ("projection1", "projection2"):
        e = torch.nn.functional.linear(e, sd[f"diffusion_embedding.{k}.weight"], sd[f"diffusion_embedding.{k}.bias"])
        e = e * torch.sigmoid(e)
    e_rows = (e.double().numpy() @ pk["rows"].T + pk["rbias"]).reshape(pk["layers"], 64)
    out = emu.emu_diffwave(pk, pk["win"], pk["wout"], audio[0].double().numpy(), init[0].double().numpy(), e_rows)
    assert pk["layers"] == 12 and pk["dilations"][9:12] == [512, 1, 2]
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    assert err < 2e-5, err                                    # fp64 emulation vs the fp32 oracle (bias hi / lo split: 2^-17)
