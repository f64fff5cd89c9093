"""Run the GINE aggregation kernels on the config-4 graph a few times (target for `ncu -k regex:gine_aggr`)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
res, bf, bb, n, e = B.measure_aggregation(torch.device("cuda:0"), iters=int(sys.argv[1]) if len(sys.argv) > 1 else 3)
print(f"n={n} e={e} fwd {res['fwd']*1e3:.1f} us ({bf/res['fwd']/1e6:.0f} GB/s)  bwd {res['bwd']*1e3:.1f} us ({bb/res['bwd']/1e6:.0f} GB/s)")
