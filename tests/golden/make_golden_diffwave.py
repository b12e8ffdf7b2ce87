"""Golden fixtures for SURVEY 8(f) item 4, ``diff2.DiffWave`` (model/diff2.py:12-158), from the REAL reference module.

The reference cannot construct the module from its own ``utils/params.py`` (no residual_channels / residual_layers /
dilation_cycle_length / n_mels entries), so the params object here carries the DiffWave base configuration
(64 channels, 30 layers, dilation cycle 10) next to the reference's 50-step noise schedule.  OUR seeded weight table is
loaded with ``strict=True`` (pins key names / shapes / order), the module runs on small seeded inputs, and the outputs go
to ``tests/golden/golden_diffwave.npz`` together with the oracle-vs-reference error.

    python tests/golden/make_golden_diffwave.py
"""
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


class AttrDict(dict):
    __getattr__ = dict.__getitem__


def seeded(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


def diffwave_weights(seed=1234):
    """default init, except the two tensors the reference zero-initialises or that would make the test blind:
    output_projection.weight is nn.init.zeros_ in the reference (:26) -- a trained checkpoint has it non-zero"""
    sd = W.init_state_dict("DiffWave", seed)
    g = torch.Generator().manual_seed(seed + 1)
    sd["output_projection.weight"] = 0.2 * torch.randn(sd["output_projection.weight"].shape, generator=g)
    return sd


def main():
    cvd = os.environ.get("CUDA_VISIBLE_DEVICES")
    sys.path.insert(0, REF)
    stub = types.ModuleType("ptflops")
    stub.get_model_complexity_info = lambda *a, **k: (0, 0)
    sys.modules["ptflops"] = stub
    import model.diff2 as diff2
    if cvd is None:
        os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    else:
        os.environ["CUDA_VISIBLE_DEVICES"] = cvd
    params = AttrDict(residual_channels=W.DIFFWAVE_CHANNELS, residual_layers=W.DIFFWAVE_LAYERS,
                      dilation_cycle_length=W.DIFFWAVE_CYCLE, n_mels=80, noise_schedule=np.linspace(1e-4, 0.05, 50).tolist())
    torch.manual_seed(0)
    m = diff2.DiffWave(None, params).eval()
    sd = diffwave_weights()
    m.load_state_dict(sd, strict=True)
    keys = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    assert [k for k, _ in keys] == list(sd.keys()), "registration order differs"
    json.dump(keys, open(os.path.join(HERE, "state_dict_keys_diffwave.json"), "w"))
    out, errs = {}, {}
    cases = {"a": (2, 3000, 11, torch.tensor([3, 41])), "b": (1, 5000, 12, torch.tensor([22.992493])),
             "c": (3, 777, 13, torch.tensor([0, 49, 7]))}
    # (fractional steps: the reference's _lerp_embedding (:83-88) multiplies a [B, 128] tensor by a [B] one, which only
    #  broadcasts for B = 1 -- case "b"; this library lerps per utterance for any B)
    with torch.no_grad():
        for tag, (B, L, seed, t) in cases.items():
            audio, init = seeded((B, L), seed), seeded((B, L), seed + 100, 0.5)
            y = m(audio, init, t)
            yo = O.diffwave_forward(sd, audio, init, t, W.DIFFWAVE_CYCLE)
            errs[tag] = float(torch.linalg.norm(yo - y) / torch.linalg.norm(y))
            out[f"diffwave_{tag}_meta"] = np.array([B, L, seed], dtype=np.int64)
            out[f"diffwave_{tag}_t"] = t.numpy()
            out[f"diffwave_{tag}_y"] = y.numpy()
    print("oracle vs reference rel-L2:", errs)
    assert max(errs.values()) < 1e-5
    out["params"] = np.array([W.DIFFWAVE_CHANNELS, W.DIFFWAVE_LAYERS, W.DIFFWAVE_CYCLE], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "golden_diffwave.npz"), **out)
    json.dump(errs, open(os.path.join(HERE, "oracle_vs_reference_diffwave.json"), "w"))
    print("parameters:", sum(p.numel() for p in m.parameters()))


if __name__ == "__main__":
    main()
