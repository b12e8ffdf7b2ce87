"""How many 16-CTA clusters of the DSMEM recurrence kernel fit on the device at once (debug hook)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from prior_diffuse_b200 import lib as plib
L = plib.load(require_device=True)
torch.zeros(1, device="cuda:0")
print("co-resident 16-CTA clusters: bp=32:", L.pdse_debug_lstm_clusters(32), " bp=16:", L.pdse_debug_lstm_clusters(16),
      " SMs:", L.pdse_sm_count())
