"""Chain latency of every block of the reference-shape step (B=8): each block is captured N times in a CUDA graph
(same two-stream schedule as the engine) and replayed back to back.  usage: exp_blocks.py [B]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from raincast_gnn_b200 import kernels as K
from raincast_gnn_b200.engine import TrainEngine
from raincast_gnn_b200.graph import build_station_graph
from raincast_gnn_b200.models import GNN
from raincast_gnn_b200.utils import synthetic as syn

dev = torch.device("cuda:0")
bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ei, ea, ei_b, ea_b = B.static_graph(bsz)
m = bsz * B.N_STATIONS
sg = build_station_graph(ei_b, ea_b, m).to(dev)
eng = TrainEngine(B.seeded_model(GNN).to(dev).train(), sg, m, B.MEMBERS, B.FEATS, use_cuda_graph=False)
x, ens = syn.node_features(m, B.MEMBERS, B.FEATS, seed=1); y = syn.log_precip_targets(m, seed=1)
eng.load_batch(x.to(dev), ens.to(dev), y.to(dev))
blk = eng._blocks
N = 20


def chain(name, fn0):
    def fn():
        with K.on_side():          # fork: the side stream always takes part in the capture, as in the engine's step
            pass
        fn0()
    side = torch.cuda.Stream()
    K.SIDE.stream = eng._side
    try:
        with torch.cuda.stream(side):
            for _ in range(2):
                fn(); K.join_side()
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            before = __import__("raincast_gnn_b200._lib", fromlist=["x"]).launch_count()
            with torch.cuda.graph(g):
                for _ in range(N):
                    fn(); K.join_side()
            launches = (__import__("raincast_gnn_b200._lib", fromlist=["x"]).launch_count() - before) // N
    finally:
        K.SIDE.stream = None
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): g.replay()
    c.record(); c.synchronize()
    t = a.elapsed_time(c) / 10 / N * 1e3
    print(f"{name:34s} {t:7.1f} us  ({launches} launches)")
    return t


Pd, Gd = blk["ds"]; Pr, Gr = blk["dr"]; Ph, Gh = blk["head"]
emb, s_ds = K.deepsets_fwd(Pd, eng.ens)
node, s_dr = K.dimred_fwd(Pr, eng.x, emb)
h = node; saved = []
for i, (Pl, _) in enumerate(blk["layers"]):
    h, s = K.gine_layer_fwd(Pl, h, eng.graph, first=(i == 0), training=True); saved.append(s)
raw, s_h = K.head_fwd(Ph, h)
_, d_raw, _ = K.crps_fwd_bwd(raw, eng.y, eng.kind, raw_input=True, u=eng.u_fixed, xi=eng.xi, t=eng.t, loss_out=eng.loss)
d = K.head_bwd(Ph, s_h, d_raw, Gh)
dl = K.gine_layer_bwd(blk["layers"][1][0], saved[1], eng.graph, d, blk["layers"][1][1], first=False, training=True)
d_emb = K.dimred_bwd(Pr, s_dr, dl, Gr)
torch.cuda.synchronize()
tot = 0.0
tot += chain("deepsets_fwd", lambda: K.deepsets_fwd(Pd, eng.ens))
tot += chain("dimred_fwd", lambda: K.dimred_fwd(Pr, eng.x, emb))
tl = chain("gine_layer_fwd (x4)", lambda: K.gine_layer_fwd(blk["layers"][1][0], node, eng.graph, first=False, training=True)); tot += 4 * tl
tot += chain("head_fwd", lambda: K.head_fwd(Ph, h))
tot += chain("crps_fwd_bwd", lambda: K.crps_fwd_bwd(raw, eng.y, eng.kind, raw_input=True, u=eng.u_fixed, xi=eng.xi, t=eng.t, loss_out=eng.loss))
tot += chain("head_bwd", lambda: K.head_bwd(Ph, s_h, d_raw, Gh))
tb = chain("gine_layer_bwd (x4)", lambda: K.gine_layer_bwd(blk["layers"][1][0], saved[1], eng.graph, d, blk["layers"][1][1], first=False, training=True)); tot += 4 * tb
tot += chain("dimred_bwd", lambda: K.dimred_bwd(Pr, s_dr, dl, Gr))
tot += chain("deepsets_bwd", lambda: K.deepsets_bwd(Pd, s_ds, d_emb, Gd))
print(f"sum of blocks {tot:.1f} us")

# ---- pieces of the DeepSets backward
from raincast_gnn_b200 import _lib
import ctypes as C
L = _lib.lib()
ens_, pooled, s2, r1, bf16 = s_ds
em, f, hdim = ens_.shape[1], ens_.shape[2], Pd["phi0_w"].shape[0]
d_pooled = torch.randn(m, hdim, device=dev)
nb = int(L.rc_deepsets_pool_bwd_nblocks(m, em, f, hdim))
part = torch.empty(nb, hdim * f + hdim, device=dev)
chain("  pool_bwd kernel", lambda: _lib.check(L.rc_deepsets_pool_bwd(ens_.data_ptr(), Pd["phi0_w"].data_ptr(), Pd["phi0_b"].data_ptr(), d_pooled.data_ptr(), part.data_ptr(), m, em, f, hdim, 0, None, torch.cuda.current_stream().cuda_stream)))
def red():
    sink = K.GradSink(dev)
    sink.add(part, Gd["phi0_w"], hdim * f + hdim, nb, hdim * f)
    sink.add(part.reshape(-1)[hdim * f:], Gd["phi0_b"], hdim * f + hdim, nb, hdim)
    sink.flush()
chain(f"  reduce of {nb} pool_bwd partials", red)
chain("  pool_fwd kernel", lambda: _lib.check(L.rc_deepsets_pool_fwd(ens_.data_ptr(), Pd["phi0_w"].data_ptr(), Pd["phi0_b"].data_ptr(), pooled.data_ptr(), m, em, f, hdim, torch.cuda.current_stream().cuda_stream)))
chain("  one dgrad GEMM (mask)", lambda: K.linear_bwd_data(d_emb, Pd["rho2_w"], mask_pos=r1))
def wg():
    sink = K.GradSink(dev)
    K.linear_bwd_weight(K.operand(d_emb, hdim), K.operand(r1, hdim), m, hdim, hdim, Gd["rho2_w"], Gd["rho2_b"], sink)
    sink.flush()
chain("  one wgrad GEMM + reduce", wg)
