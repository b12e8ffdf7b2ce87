"""Is the ~60-cycle cost of a small-N tcgen05.mma a dependency on the accumulator, the shared-memory A read, or the issue path?"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.gpu_probe_tmem import run  # noqa: E402

for n in (16, 32, 64):
    for mode, what in ((1, "one accumulator, A in shared memory"), (1 | 8, "four accumulators in turn, A in shared memory"),
                       (1 | 16, "one accumulator, A in tensor memory"), (1 | 32, "four accumulators in turn, A in tensor memory")):
        o = run(mode, mma_n=n, ctas=1, iters=2000)
        print(f"N={n:3d} {what:48s}: {o[2] / o[1]:7.1f} cycles per MMA")
for n in (16, 64, 128):
    o = run(1 | 64, mma_n=n, ctas=1, iters=2000)
    print(f"N={n:3d} two issuer warps, an accumulator each: {o[2] / o[1]:7.1f} and {o[3] / 1000:7.1f} cycles per MMA per issuer")
