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
    assert L.pdse_abi_version() == 3 and L.pdse_operand_format() in (0, 1)
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


# ------------------------------------------------------------------ C-only host (tests/abi_host.c)
def build_abi_host(tmpdir):
    """gcc-compile the C host against include/pdse.h and link it to the in-tree library + the CUDA runtime"""
    import shutil
    import subprocess
    lib_dir = os.path.dirname(build.build())
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = os.path.join(str(tmpdir), "abi_host")
    cmd = [shutil.which("gcc") or "gcc", "-std=c11", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(cuda, "include"), os.path.join(ROOT, "tests", "abi_host.c"), "-o", exe,
           "-L", lib_dir, "-lpdse", "-L", os.path.join(cuda, "lib64"), "-lcudart",
           f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def write_abi_input(path, tensors):
    import struct
    import numpy as np
    with open(path, "wb") as f:
        f.write(struct.pack("<i", len(tensors)))
        for name, t in tensors.items():
            a = np.ascontiguousarray(t.detach().cpu().float().numpy()).reshape(-1)
            f.write(struct.pack("<i", len(name)) + name.encode() + struct.pack("<q", a.size))
            f.write(a.tobytes())


def abi_state(B, T, seed=1):
    from prior_diffuse_b200 import weights as W
    g = W.randomize_norm_stats(W.init_state_dict("GCRN", 1234), 4321)
    d = W.randomize_norm_stats(W.init_state_dict("DiffUNet1", 1234), 4321)
    gen = torch.Generator().manual_seed(seed)
    tensors = {"@shape": torch.tensor([float(B), float(T)]),
               "@x": torch.randn(B, 2, T, 161, generator=gen), "@x0": 0.3 * torch.randn(B, 2, T, 161, generator=gen),
               "@t": torch.tensor([22.992493]), "@y": torch.randn(B, 2, T, 161, generator=gen)}
    for k, v in d.items():
        if v.is_floating_point():
            tensors["ddpm/" + k] = v
    for k, v in g.items():
        if v.is_floating_point():
            tensors["gcrn/" + k] = v
    return g, d, tensors


def test_c_host_compiles_links_and_packs_without_python(tmp_path):
    """a host written in C sees everything it needs in include/pdse.h: the program builds with -Wall -Werror, and its
    packing leg (no GPU needed) yields exactly the blobs the Python binding gets from the same entry points"""
    import subprocess
    import numpy as np
    exe = build_abi_host(tmp_path)
    g, d, tensors = abi_state(1, 4)
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    write_abi_input(inp, tensors)
    r = subprocess.run([exe, inp, out, "--pack-only"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    blob = np.fromfile(out, dtype=np.uint8)
    bd, _ = lib.pack_state_dict(lib.NET_DIFFUNET1, d)
    bg, _ = lib.pack_state_dict(lib.NET_GCRN, g)
    assert blob.size == bd.size + bg.size
    assert np.array_equal(blob[:bd.size], bd) and np.array_equal(blob[bd.size:], bg)
