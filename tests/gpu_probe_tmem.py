"""Microbenchmarks behind the kernel design notes in DESIGN.md (run on the GPU box):
TMEM drain rate (tcgen05.ld) alone / under a concurrent MMA stream, MMA issue rate, bulk-copy rate per SM."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prior_diffuse_b200 import lib as plib  # noqa: E402


def run(mode, iters=200, mma_n=256, ld_cols=16, copy_bytes=0, ctas=148, cta_stride=64 * 32768, nblk=64):
    L = plib.load(require_device=True)
    out = torch.zeros(4 * ctas, dtype=torch.int64, device="cuda")
    src = torch.zeros(ctas * 64 * 32768 + 32768, dtype=torch.uint8, device="cuda")
    plib.check(L.pdse_probe_tmem(plib.ptr(out), plib.ptr(src), mode, iters, mma_n, ld_cols, copy_bytes, ctas, cta_stride, nblk,
                                 plib.stream_ptr()))
    torch.cuda.synchronize()
    o = out.view(ctas, 4).double().mean(0)
    return o


def main():
    for ctas in (1, 148):
        print(f"--- {ctas} CTA(s)")
        for ld in (16, 32):
            o = run(2, ld_cols=ld, ctas=ctas)
            print(f"drain alone            x{ld}: {o[0]:9.0f} cycles per 128x256 fp32 ({131072 / o[0]:.1f} B/cycle)")
        for n in (16, 32, 64, 128, 256):
            o = run(1, mma_n=n, ctas=ctas, iters=2000)
            print(f"MMA alone  N={n:3d}: {o[2] / o[1]:7.1f} cycles per 128xNx16 MMA  (floor {n / 2})")
        for n in (64, 128):
            for shift, lbo in ((0, 0), (1, 0), (3, 0), (4, 0), (8, 0), (0, 468), (3, 468), (0, 472), (3, 472)):
                o = run(1 | (shift << 8) | (lbo << 16), mma_n=n, ctas=ctas, iters=2000)
                print(f"MMA alone  N={n:3d}, A window shifted by {shift} rows, plane stride {lbo or 128} rows: {o[2] / o[1]:7.1f} cycles")
        for n in (64, 256):
            for ld in (16, 32):
                o = run(3, mma_n=n, ld_cols=ld, ctas=ctas)
                print(f"drain + MMA N={n:3d} x{ld}: {o[0]:9.0f} cycles per drain ({131072 / o[0]:.1f} B/cycle); "
                      f"{o[2] / o[1]:7.1f} cycles per MMA")
        for cb in (2048, 8192, 32768):
            o = run(4, copy_bytes=cb, ctas=ctas)
            print(f"bulk copy {cb:6d} B x4 in flight: {o[3] / 1000:6.1f} B/cycle per SM")
        for cb in (2048, 16384, 32768):
            o = run(4, copy_bytes=cb, ctas=ctas, cta_stride=0, nblk=64, iters=400)
            print(f"bulk copy {cb:6d} B, ALL CTAs read the SAME 2 MB (L2 hits): {o[3] / 1000:6.1f} B/cycle per SM")
            o = run(4, copy_bytes=cb, ctas=ctas, cta_stride=8 * 32768, nblk=8, iters=400)
            print(f"bulk copy {cb:6d} B, private 256 KB per CTA (L2 hits):       {o[3] / 1000:6.1f} B/cycle per SM")
        o = run(7, mma_n=256, ld_cols=32, copy_bytes=32768, ctas=ctas)
        print(f"all three: drain {131072 / o[0]:.1f} B/cycle, {o[2] / o[1]:.1f} cycles per MMA, copy {o[3] / 1000:.1f} B/cycle")


if __name__ == "__main__":
    main()
