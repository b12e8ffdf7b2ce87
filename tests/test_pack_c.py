"""The C packers of the ABI (csrc/pack.cpp: pdse_pack_diffunet1 / pdse_pack_gcrn) against the NumPy statement of the
operand layouts (prior_diffuse_b200/pack.py, itself pinned against the oracle by test_pack_emulation.py): every section
of the blob, element for element.  Both sides compute in float64 and round once to fp32 (then to the 16-bit operand format, fp16 by default); the only
freedom is the summation order of the composed matrices, so a handful of last-bit differences are tolerated and counted."""
import numpy as np
import pytest
import torch

from prior_diffuse_b200 import lib, pack as P, weights as W


def _weights(name):
    return W.randomize_norm_stats(W.init_state_dict(name, seed=1234), seed=4321)


def _bf16_bits(a64):     # bit patterns in the library's operand format (csrc/opfmt.h)
    return torch.from_numpy(a64.astype(np.float32)).to(lib.op_dtype()).view(torch.int16).numpy()


def _section(blob, directory, name):
    dtype, off, n = directory[name]
    if dtype == 0:
        return blob[off:off + 2 * n].view(np.int16)
    return blob[off:off + 4 * n].view(np.float32)


def _compare(blob, directory, name, ref64, stats):
    got = _section(blob, directory, name)
    assert got.size == ref64.size, (name, got.size, ref64.size)
    if directory[name][0] == 0:
        ref = _bf16_bits(ref64.reshape(-1))
        bad = got != ref
        if bad.any():     # a differing last bit of a 16-bit value: adjacent bit patterns only
            assert np.abs(got[bad].astype(np.int32) - ref[bad].astype(np.int32)).max() <= 1, name
    else:
        ref = ref64.reshape(-1).astype(np.float32)
        bad = got != ref
        if bad.any():
            assert np.allclose(got[bad], ref[bad], rtol=3e-7, atol=1e-30), name
    stats[0] += int(bad.sum())
    stats[1] += got.size


def test_c_packer_matches_numpy_packer_diffunet1():
    sd = _weights("DiffUNet1")
    blob, directory = lib.pack_state_dict(lib.NET_DIFFUNET1, sd)
    ref = P.pack_diffunet1(sd)
    stats = [0, 0]
    for name, b in ref.items():
        if name == "time":
            for k, v in b.items():
                _compare(blob, directory, "time." + k, np.asarray(v, dtype=np.float64), stats)
            continue
        _compare(blob, directory, name + ".wb", b.flat("h").astype(np.float64), stats)
        _compare(blob, directory, name + ".wf", b.flat("f").astype(np.float64), stats)
    assert stats[1] > 2_000_000 and stats[0] <= stats[1] * 1e-5, stats
    # the sinusoid table is shipped as exact bit patterns (csrc/time_table.inc): identical to torch's, not merely close
    from tests.golden.make_time_table import table_bits
    assert np.array_equal(_section(blob, directory, "time.table").view(np.uint32), table_bits())


def test_c_packer_matches_numpy_packer_gcrn():
    sd = _weights("GCRN")
    blob, directory = lib.pack_state_dict(lib.NET_GCRN, sd)
    ref = P.pack_gcrn(sd)
    stats = [0, 0]
    for name, b in ref.items():
        if name.startswith("lstm"):
            _compare(blob, directory, name + ".w_ih", b.h["w_ih"].astype(np.float64), stats)
            _compare(blob, directory, name + ".w_hh", b.h["w_hh"].astype(np.float64), stats)
            _compare(blob, directory, name + ".wf", b.flat("f").astype(np.float64), stats)
        elif name.startswith("dec"):
            _compare(blob, directory, name + ".w_even", b.h["w_even"].astype(np.float64), stats)
            _compare(blob, directory, name + ".w_odd", b.h["w_odd"].astype(np.float64), stats)
            _compare(blob, directory, name + ".wf", b.flat("f").astype(np.float64), stats)
        else:
            if b.h:
                _compare(blob, directory, name + ".wb", b.flat("h").astype(np.float64), stats)
            _compare(blob, directory, name + ".wf", b.flat("f").astype(np.float64), stats)
    assert stats[1] > 9_000_000 and stats[0] <= stats[1] * 1e-5, stats


def test_c_packer_reports_missing_entries_and_workspace_sizes():
    L = lib.load()
    sd = _weights("DiffUNet1")
    sd.pop("en.conv3.l.weight")
    with pytest.raises(RuntimeError, match="en.conv3.l.weight"):
        lib.pack_state_dict(lib.NET_DIFFUNET1, sd)
    small, big = L.pdse_workspace_bytes(lib.NET_DIFFUNET1, 2, 40), L.pdse_workspace_bytes(lib.NET_DIFFUNET1, 64, 301)
    assert 0 < small < big < 8 << 30
    assert L.pdse_workspace_bytes(lib.NET_GCRN, 64, 301) > 0
    assert L.pdse_workspace_bytes(lib.NET_GCRN, 65, 301) < 0 and L.pdse_workspace_bytes(7, 1, 1) < 0


def test_fp16_conversion_of_the_c_packer_matches_torch():
    """subnormals, ties, the largest finite value and the overflow guard of the C packer's fp32 -> fp16 conversion"""
    if lib.op_dtype() != torch.float16:
        pytest.skip("library built with bf16 operands")
    sd = _weights("GCRN")
    vals = torch.tensor([0.0, -0.0, 5.9604645e-8, 2.9802322e-8, 2.98023224e-8 * 1.0001, 8.9406967e-8, 6.0975552e-5, 6.1035156e-5, 6.1e-5,
                         1.0009766, 1.00048828125, 1.00146484375, 65504.0, 65519.9, -3.3e-5, 1e-7, 0.333333, 1234.567, -2047.5, 4.1e-6])
    # w_hh is copied into its section without arithmetic: plant the probe values there
    whh = sd["glstm.lstm_list1.0.weight_hh_l0"].clone()
    whh.view(-1)[:vals.numel()] = vals
    sd["glstm.lstm_list1.0.weight_hh_l0"] = whh
    blob, directory = lib.pack_state_dict(lib.NET_GCRN, sd)
    got = _section(blob, directory, "lstm1_0.w_hh")
    # w_hh section = 16 slices [64][128][8] of rows n = cta*128 + lane; reference row 0 (gate 0, unit 0) is n = 0: its first
    # 20 K-elements sit at plane kc = k // 8, row 0, element k % 8
    pos = [(k // 8) * 128 * 8 + (k % 8) for k in range(vals.numel())]
    want = vals.to(torch.float16).view(torch.int16).numpy()
    assert np.array_equal(got[pos], want), (got[pos], want)
    whh.view(-1)[0] = 70000.0
    with pytest.raises(RuntimeError, match="fp16 operand range"):
        lib.pack_state_dict(lib.NET_GCRN, sd)
