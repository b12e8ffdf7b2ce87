"""Per-phase cycle breakdown of the LSTM recurrence (debug hook)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from prior_diffuse_b200 import lib as plib
from prior_diffuse_b200.gcrn import GCRNEngine

dev = torch.device("cuda:0")
g, _ = bench.seeded_weights()
eng = GCRNEngine(g, dev)
y = torch.randn(64, 2, 301, 161, device=dev)
eng.forward(y)
torch.cuda.synchronize()
prof = torch.zeros(6, dtype=torch.int64, device=dev)
L = plib.load()
L.pdse_debug_lstm_prof(plib.ptr(prof))
eng.forward(y)
torch.cuda.synchronize()
L.pdse_debug_lstm_prof(None)
names = (["barrier wait", "h load+sts", "sync+MMA", "gates+sync", "cell+h stores", "tail"] if os.environ.get("PDSE_LSTM_MODE") in ("0", "1")
         else ["wait peers' h", "CTA barrier", "MMA issue+done", "gates+barrier", "cell+stage+push", "h stores"])
tot = prof.sum().item()
for n, v in zip(names, prof.tolist()):
    print(f"{n:16s} {v / 301:9.0f} cycles/step  {100 * v / tot:5.1f}%")
print("total cycles/step", tot / 301)
