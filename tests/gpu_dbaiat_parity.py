import sys, torch
sys.path.insert(0, '/root/repo')
from oracle import pdse_oracle as O
from prior_diffuse_b200 import weights as W
from prior_diffuse_b200.dbaiat import DBAIATEngine
dev = torch.device("cuda:0")
sd = W.randomize_norm_stats(W.init_state_dict("aia_complex_trans_ri", 1234), 4321)
eng = DBAIATEngine(sd, dev)
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(torch.linalg.norm(a - b) / torch.linalg.norm(b))
for B, T, seed in ((3, 37, 77), (2, 301, 5), (1, 1001, 6)):
    x = torch.randn(B, 2, T, 161, generator=torch.Generator().manual_seed(seed))
    taps = {}
    ref = O.dbaiat_forward(sd, x, taps) / 11.0
    y = eng.forward(x.to(dev))
    ws = eng.workspace(B, T)
    print(f"B={B} T={T}: X_init rel-L2 {rel(y, ref):.3e}; aia state {rel(ws['S'].view(B, T, 80, 32).permute(0, 3, 1, 2), taps['aia_state3']):.3e}", flush=True)
