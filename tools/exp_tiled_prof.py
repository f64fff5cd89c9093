"""One forward launch of the instrumented tiled kernel (RC_B200_LIB=tools/librc_prof.so): per-CTA cycle split."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raincast_gnn_b200 import graph as G, kernels as K
from raincast_gnn_b200.utils import synthetic as syn
dev = torch.device("cuda:0")
n, h = 100_000, 128
ei, ea = G.radius_graph_from_coords(syn.station_coords(n, 1000.0, 0), syn.scaled_graph_radius(n, 1000.0))
sg = G.build_station_graph(ei, ea, n).to(dev)
g = torch.Generator().manual_seed(0)
x = torch.randn(n, h, generator=g).to(dev)
w, b, eps = torch.randn(h, generator=g).to(dev), torch.randn(h, generator=g).to(dev), torch.zeros(1, device=dev)
out = torch.empty_like(x)
K.gine_aggr_fwd(x, sg, w, b, eps, out, tiled=True)
torch.cuda.synchronize()
