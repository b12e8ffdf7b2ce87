"""Drop-in module surface: constructor, state_dict layout, checkpoint round trip (CPU only)."""
import json
import os

import pytest
import torch

from prior_diffuse_b200 import GCRN, DiffUNet, DiffUNet1, DiffWave, Nocon, aia_complex_trans_ri
from prior_diffuse_b200 import weights as W

HERE = os.path.dirname(os.path.abspath(__file__))


class _Params(dict):
    __getattr__ = dict.__getitem__


@pytest.mark.parametrize("cls", [GCRN, DiffUNet1, DiffUNet, aia_complex_trans_ri, Nocon])
def test_state_dict_layout_matches_reference(cls):
    ref = json.load(open(os.path.join(HERE, "golden", "state_dict_keys.json")))[cls.__name__]
    m = cls(_Params(noise_schedule=[0.0] * 50)) if cls in (DiffUNet1, Nocon) else cls()
    sd = m.state_dict()
    assert list(sd) == [k for k, _, _ in ref]
    for k, shape, dtype in ref:
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == dtype, k
    assert "time_embedding.embedding" not in sd     # persistent=False in the reference (diff3.py:65)


def test_diffwave_state_dict_layout_matches_reference():
    # tests/golden/make_golden_diffwave.py loaded this table into the reference's diff2.DiffWave with strict=True
    ref = json.load(open(os.path.join(HERE, "golden", "state_dict_keys_diffwave.json")))
    for m in (DiffWave(), DiffWave(None, _Params(residual_channels=64, residual_layers=30, dilation_cycle_length=10))):
        sd = m.state_dict()
        assert list(sd) == [k for k, _ in ref]
        assert all(list(sd[k].shape) == shape for k, shape in ref)
        assert "diffusion_embedding.embedding" not in sd      # persistent=False (diff2.py:68)
    assert len(DiffWave(None, _Params(residual_layers=4)).state_dict()) < len(ref)
    with pytest.raises(ValueError):
        DiffWave(None, _Params(residual_channels=128))
    with pytest.raises(RuntimeError):
        DiffWave().eval()(torch.zeros(1, 100), torch.zeros(1, 100), torch.zeros(1))      # on the CPU: no silent fallback


def test_checkpoint_round_trip(tmp_path):
    # trainer/complex_ddpm_trainer.py:616-622 saves [prior_sd, opt_sd, ddpm_sd, opt_ddpm_sd]
    g, d = GCRN(), DiffUNet1()
    torch.save([W.init_state_dict("GCRN", 3), {}, W.init_state_dict("DiffUNet1", 4), {}], tmp_path / "best_checkpoint.pth")
    ck = torch.load(tmp_path / "best_checkpoint.pth")
    g.load_state_dict(ck[0])
    d.load_state_dict(ck[2])
    assert torch.equal(g.state_dict()["fc1.weight"], ck[0]["fc1.weight"])
    assert torch.equal(d.state_dict()["TCMs.2.residual6.conv2.2.bias"], ck[2]["TCMs.2.residual6.conv2.2.bias"])
    with pytest.raises(RuntimeError):
        g.load_state_dict({"nope": torch.zeros(1)})


def test_default_init_is_seeded_by_the_global_rng():
    torch.manual_seed(1234)
    a = GCRN().state_dict()["conv3.conv1.weight"]
    torch.manual_seed(1234)
    b = GCRN().state_dict()["conv3.conv1.weight"]
    assert torch.equal(a, b)
    assert float(a.abs().max()) <= 1.0 / (32 * 3) ** 0.5 + 1e-7


def test_training_mode_is_rejected():
    m = GCRN()
    assert m.training
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 2, 4, 161))
