"""Two launches each of the tiled forward / backward aggregation on the config-4 graph (target for
`ncu --launch-skip 2 -c 2 -k regex:gine_aggr_.*tiled`)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import graph as G, kernels as K
from raincast_gnn_b200.utils import synthetic as syn
dev = torch.device("cuda:0")
n, h = 100_000, int(sys.argv[1]) if len(sys.argv) > 1 else 128
ei, ea = G.radius_graph_from_coords(syn.station_coords(n, 1000.0, 0), syn.scaled_graph_radius(n, 1000.0))
sg = G.build_station_graph(ei, ea, n).to(dev)
g = torch.Generator().manual_seed(0)
x, gout = torch.randn(n, h, generator=g).to(dev), torch.randn(n, h, generator=g).to(dev)
w, b, eps = torch.randn(h, generator=g).to(dev), torch.randn(h, generator=g).to(dev), torch.zeros(1, device=dev)
out = torch.empty_like(x)
for _ in range(2):
    K.gine_aggr_fwd(x, sg, w, b, eps, out, tiled=True)
    K.gine_aggr_bwd(gout, x, sg, w, b, eps, None, out, tiled=True)
torch.cuda.synchronize()
