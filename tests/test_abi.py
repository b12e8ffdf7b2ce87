"""C-ABI library: builds for sm_100a, loads, exports every symbol include/pdse.h declares.
No compute call is made here (no GPU in this container)."""
import os
import re

import pytest
import torch

from prior_diffuse_b200 import build, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pdse.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pdse_\w+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    L = lib.load()
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/pdse.h but not exported"
    assert sorted(lib.exported_symbols()) == names, "ctypes signature table and header disagree"
    assert L.pdse_abi_version() == 2
    assert L.pdse_bias_row_floats() == 452


def test_sass_contains_blackwell_instructions():
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass      # tcgen05.mma
    assert "LDTM" in sass         # tcgen05.ld
    assert "UBLKCP" in sass       # cp.async.bulk
    # legacy mma.sync only where it is the right tool: the head_dim = 8 TF32 attention of the DB-AIAT prior
    # (short and K/V-streaming variants) and the 161 x 161 fp32-accurate (3xTF32) fc of the GCRN output stage
    legacy = [blk.split("\n", 1)[0].strip() for blk in sass.split("Function :")[1:] if "HMMA." in blk]
    allowed = ("aia_attn_kernel", "aia_attn_long_kernel", "gout_kernel")
    assert legacy and all(any(k in name for k in allowed) for name in legacy), legacy


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu():
    from prior_diffuse_b200 import GCRN, DiffUNet1, signal
    with pytest.raises(RuntimeError):
        lib.load(require_device=True)
    with pytest.raises(RuntimeError):
        GCRN().eval()(torch.zeros(1, 2, 4, 161))
    with pytest.raises(RuntimeError):
        DiffUNet1().eval()(torch.zeros(1, 2, 4, 161), torch.zeros(1, 2, 4, 161), torch.zeros(1))
    with pytest.raises((RuntimeError, AssertionError)):
        signal.stft(torch.zeros(1, 1600))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "prior_diffuse_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+.*oracle", src, flags=re.M), f"{f} imports the oracle"
