import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_model import _model_kw, oracle_step
from raincast_gnn_b200.models import GNN
from raincast_gnn_b200.pyg_compat import DataLoader
from raincast_gnn_b200.utils import synthetic as syn
from raincast_gnn_b200.utils.dataset import SyntheticEUPPBench
dev = torch.device("cuda:0")
h = int(sys.argv[1]); mode = sys.argv[2]
ds = SyntheticEUPPBench(n_dates=2, members=11)
batch = next(iter(DataLoader(ds, batch_size=2)))
kw = _model_kw(dict(f=35, h=h, layers=4, loss="MixedLoss", grad_u="True"))
ours = GNN(**kw); sd = syn.seeded_state_dict(ours.state_dict(), seed=5); ours.load_state_dict(sd); ours.to(dev).train()
ours.deepset.compute_dtype = mode
import copy
if len(sys.argv) > 3:      # bf16-consistent oracle: ensemble and phi[0].weight rounded to bf16 (what the tensor cores see)
    sd = dict(sd); sd["deepset.phi.0.weight"] = sd["deepset.phi.0.weight"].bfloat16().float()
    ob = copy.copy(batch); ob.ensemble = batch.ensemble.bfloat16().float()
    p64, l64, g64, _ = oracle_step(kw, sd, ob, torch.float64)
else:
    p64, l64, g64, _ = oracle_step(kw, sd, batch, torch.float64)
b = batch.to(dev); p = ours(b); l = ours.loss_fn.crps(p, b.y); l.backward()
print("preds", float((p.detach().cpu() - p64).abs().max() / p64.abs().max()), "loss", abs(l.item() - l64.item()) / abs(l64.item()))
for k, v in ours.named_parameters():
    d = v.grad.cpu().double() - g64[k]
    print(f"{k:40s} l2 {float(d.norm()/g64[k].norm()):9.2e}  max {float(d.abs().max()/g64[k].abs().max()):9.2e}")
