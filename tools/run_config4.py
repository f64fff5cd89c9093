"""BASELINE.json config 4: one 100k-node station graph (2 978 560 edges), 51 members, H=128, L=4, mixed_u.
Times the blocks of one training step with CUDA events and (optionally) checks preds / loss against the CPU oracle."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from raincast_gnn_b200 import graph as G, kernels as K
from raincast_gnn_b200.engine import TrainEngine
from raincast_gnn_b200.models import GNN
from raincast_gnn_b200.utils import synthetic as syn

dev = torch.device("cuda:0")
n, em = 100_000, int(sys.argv[1]) if len(sys.argv) > 1 else 51
check = len(sys.argv) > 2 and sys.argv[2] == "check"
hidden = int(sys.argv[3]) if len(sys.argv) > 3 else 128          # 512 + "bf16" as 5th argument: config 5 at this shape
bf16 = len(sys.argv) > 4 and sys.argv[4] == "bf16"
ei, ea = G.radius_graph_from_coords(syn.station_coords(n, 1000.0, 0), syn.scaled_graph_radius(n, 1000.0))
sg = G.build_station_graph(ei, ea, n).to(dev)
x, ens = syn.node_features(n, em, B.FEATS, seed=3)
y = syn.log_precip_targets(n, seed=3)
kw = dict(B.MODEL_KW, hidden_channels_gnn=hidden, out_channels_gnn=hidden)
model = GNN(**kw)
model.load_state_dict(syn.seeded_state_dict(model.state_dict(), seed=2024))
model = model.to(dev).train()
if bf16:
    model.deepset.compute_dtype = "bf16"
eng = TrainEngine(model, sg, n, em, B.FEATS, use_cuda_graph=False)
eng.load_batch(x.to(dev), ens.to(dev), y.to(dev))
blk = eng._blocks

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

for it in range(3):
    t = [("start", ev())]
    emb, s_ds = K.deepsets_fwd(blk["ds"][0], eng.ens, bf16=bf16); t.append(("deepsets fwd", ev()))
    node, s_dr = K.dimred_fwd(blk["dr"][0], eng.x, emb); t.append(("dim_red fwd", ev()))
    saved, h = [], node
    for i, (Pl, _) in enumerate(blk["layers"]):
        h, s = K.gine_layer_fwd(Pl, h, eng.graph, first=(i == 0), training=True); saved.append(s)
    t.append(("4 GINE layers fwd", ev()))
    raw, s_h = K.head_fwd(blk["head"][0], h)
    loss, d_raw, _ = K.crps_fwd_bwd(raw, eng.y, eng.kind, raw_input=True, xi=0.5, t=5.0); t.append(("head + CRPS", ev()))
    d = K.head_bwd(blk["head"][0], s_h, d_raw, blk["head"][1])
    for i in reversed(range(4)):
        d = K.gine_layer_bwd(blk["layers"][i][0], saved[i], eng.graph, d, blk["layers"][i][1], first=(i == 0), training=True)
    t.append(("head + 4 GINE layers bwd", ev()))
    d_emb = K.dimred_bwd(blk["dr"][0], s_dr, d, blk["dr"][1]); t.append(("dim_red bwd", ev()))
    K.deepsets_bwd(blk["ds"][0], s_ds, d_emb, blk["ds"][1]); t.append(("deepsets bwd", ev()))
    eng._optimizer(); t.append(("AdamW", ev()))
    torch.cuda.synchronize()
    if it == 2:
        tot = t[0][1].elapsed_time(t[-1][1])
        print(f"config 4 step (Em={em}): {tot:.2f} ms   loss {loss.item():.6f}   peak mem {torch.cuda.max_memory_allocated()/2**30:.2f} GiB")
        for (a, ea_), (b_, eb) in zip(t[:-1], t[1:]):
            print(f"   {b_:28s} {ea_.elapsed_time(eb):8.3f} ms")
if check:
    from oracle import model as om, pyg as opyg
    ref = B.seeded_model(om.GNN); ref.train()
    torch.set_num_threads(os.cpu_count())
    t0 = time.time()
    ob = opyg.Data(x=x, ensemble=ens, edge_index=ei, edge_attr=ea, y=y)
    p = ref(ob); l = ref.loss_fn.crps(p, y); l.backward()
    print(f"oracle (CPU fp32, {os.cpu_count()} threads): {time.time()-t0:.1f} s fwd+bwd, loss {l.item():.6f}")
    model2 = B.seeded_model(GNN).to(dev).train()
    from raincast_gnn_b200.pyg_compat import Data
    dd = Data(x=x.to(dev), ensemble=ens.to(dev), edge_index=ei.to(dev), edge_attr=ea.to(dev), y=y.to(dev)); dd.station_graph = sg
    p2 = model2(dd); l2 = model2.loss_fn.crps(p2, dd.y); l2.backward()
    print("rel err preds", float((p2.detach().cpu() - p.detach()).abs().max() / p.detach().abs().max()), "loss", abs(l2.item() - l.item()) / abs(l.item()))
    worst = 0
    for (k, a), (_, b_) in zip(model2.named_parameters(), ref.named_parameters()):
        sc = b_.grad.abs().max().item()
        if k.endswith(".nn.0.bias"): continue
        worst = max(worst, (a.grad.cpu() - b_.grad).abs().max().item() / sc)
    print("worst gradient rel err vs fp32 oracle", worst)
