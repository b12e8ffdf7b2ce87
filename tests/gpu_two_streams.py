"""Experiment: one 64-utterance graph per pass vs two 32-utterance graphs on two streams (the LSTM recurrence of one half
occupies 32 SMs for 1.3 ms; the other half's denoiser kernels can use the rest).   python tests/gpu_two_streams.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from prior_diffuse_b200 import Enhancer  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    g, d = bench.seeded_weights()
    wav = bench.synthetic_wav(64).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    one = Enhancer(g, d, dev)
    for _ in range(3):
        one.enhance(wav)
    torch.cuda.synchronize()
    K = 10
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.zero_()
    a.record()
    for _ in range(K):
        one.enhance(wav)
    b.record()
    torch.cuda.synchronize()
    print(f"one graph of 64:        {a.elapsed_time(b) / K:.3f} ms per 64 utterances (back to back, no flush)")

    halves = [Enhancer(g, d, dev) for _ in range(2)]
    streams = [torch.cuda.Stream(dev) for _ in range(2)]
    w = [wav[:32].contiguous(), wav[32:].contiguous()]
    for h, s, x in zip(halves, streams, w):
        with torch.cuda.stream(s):
            for _ in range(3):
                h.enhance(x)
    torch.cuda.synchronize()
    for stagger in (False, True):
        flush.zero_()
        torch.cuda.synchronize()
        a.record()
        for s in streams:
            s.wait_event(a)
        for k in range(K):
            for i, (h, s, x) in enumerate(zip(halves, streams, w)):
                with torch.cuda.stream(s):
                    if stagger and k == 0 and i == 1:
                        torch.cuda._sleep(int(5e6))      # ~2.5 ms: start the second stream half a pass later
                    h.enhance(x)
        for s in streams:
            ev = torch.cuda.Event()
            ev.record(s)
            torch.cuda.current_stream().wait_event(ev)
        b.record()
        torch.cuda.synchronize()
        print(f"two graphs of 32, two streams (stagger={stagger}): {a.elapsed_time(b) / K:.3f} ms per 64 utterances")
    # single stream, two halves back to back (what splitting alone costs)
    a.record()
    for k in range(K):
        for h, x in zip(halves, w):
            h.enhance(x)
    b.record()
    torch.cuda.synchronize()
    print(f"two graphs of 32, one stream: {a.elapsed_time(b) / K:.3f} ms per 64 utterances")


if __name__ == "__main__":
    main()
