"""Graphed training step: the explicit kernel schedule of one train.py iteration (train.py:61-71)

    H2D batch -> DeepSets -> dim_red -> L x GINE layer -> head -> links+CRPS (value and gradient)
              -> backward of every block -> [sum of the ranks' flat gradients] -> fused AdamW

captured once in a CUDA graph (optimiser and peer-memory gradient exchange included) and replayed, with parameters,
gradients and Adam moments in flat buffers.
No autograd, no per-step Python dispatch, no host sync: the loss stays on the device (train.py's per-step
`loss.item()` becomes one read per epoch).  Data-parallel semantics are DDP's: per-rank BatchNorm
statistics, per-rank mean over valid nodes, mean of rank gradients.  The 0.84 MB gradient bucket
is exchanged once per step over NVLink: every rank's flat gradient sits in peer-mapped memory and the AdamW kernel
sums the ranks' gradients itself (rc_p2p_*: barrier, sum + AdamW, barrier); without peer memory (or with
RC_DP_EXCHANGE=nccl) one NCCL all-reduce feeds rc_adamw_step, the 1/world factor folded into the kernel.

The module keeps owning the parameters: `model.parameters()` become views of the flat buffer, so
state_dict() / .ckpt saving keep working while the engine trains.
"""
from __future__ import annotations

import os

import torch

from . import _lib
from . import kernels as K
from .graph import StationGraph
from .models.model_utils import loss_kind


_ALIGN = 4      # floats: every parameter view starts on a 16-byte boundary (the kernels use 128-bit loads)


def _padded(n: int) -> int:
    return (n + _ALIGN - 1) // _ALIGN * _ALIGN


def _flatten_into(tensors, flat):
    off = 0
    views = []
    for t in tensors:
        n = t.numel()
        views.append(flat[off:off + n].view(t.shape))
        off += _padded(n)
    return views


FUSE_HEAD = os.environ.get("RC_FUSE_HEAD", "1") != "0"      # 0: head GEMM, CRPS kernel and head backward as separate launches (A/B, tests)


class TrainEngine:
    def __init__(self, model, graph: StationGraph, num_nodes: int, members: int, feats: int, *, lr: float = 1e-4,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01, process_group=None,
                 use_cuda_graph: bool = True):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise _lib.RcError("TrainEngine needs the model on a CUDA device (there is no CPU path)")
        self.model, self.graph, self.device = model, graph.to(dev), dev
        self.m, self.members, self.feats = num_nodes, members, feats
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.kind = loss_kind(model.loss, model.grad_u)
        self.u_fixed = 0.0 if model.grad_u == "True" else float(model.u)
        self.xi = float(model.xi)
        self.t = float(getattr(model.loss_fn, "t", 5.0))
        # ---- flat parameter / gradient / moment buffers (parameters become views)
        params = list(model.parameters())
        self.names = [n for n, _ in model.named_parameters()]
        total = sum(_padded(p.numel()) for p in params)      # padding stays zero: zero gradient, zero update
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        views = _flatten_into(params, self.flat_p)
        with torch.no_grad():
            for p, v in zip(params, views):
                v.copy_(p.detach().float())
                p.data = v
        # the flat gradient: under data parallelism it lives where the other ranks of the box can read it (NVLink peer
        # memory), and the step sums the peers' gradients inside the AdamW kernel instead of an NCCL all-reduce
        self.p2p = None
        self.flat_g = None
        force = os.environ.get("RC_DP_FORCE_P2P", "0") == "1" and process_group is not None     # (measurements: the exchange path at world 1)
        if (self.world > 1 or force) and os.environ.get("RC_DP_EXCHANGE", "p2p") == "p2p":
            self.flat_g = self._setup_peer_exchange(total, dev)
        if self.flat_g is None:
            self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grads = dict(zip(self.names, _flatten_into(params, self.flat_g)))
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.step_count = torch.zeros((), dtype=torch.int64, device=dev)
        self.n_params = total
        # ---- static inputs / outputs
        self.x = torch.zeros((num_nodes, feats), dtype=torch.float32, device=dev)
        self.ens = torch.zeros((num_nodes, members, feats), dtype=torch.float32, device=dev)
        self.y = torch.zeros((num_nodes,), dtype=torch.float32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float64, device=dev)
        self.loss_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        self._blocks = self._bind(model)
        self.use_cuda_graph = use_cuda_graph
        self._side = torch.cuda.Stream(device=dev)      # weight-gradient work overlaps the data-gradient chain
        self._side2 = torch.cuda.Stream(device=dev) if os.environ.get("RC_SIDE_ALT", "1") != "0" else None   # tail of backward
        # input prefetch: the next batch travels host -> staging buffers on a copy stream while the current step runs
        # In graph mode the staging buffers are a second input set with its own captured graph (same kernels, same
        # pool): taking the prefetched batch is a flip, not a copy.  Other modes copy device -> device.
        self._copy = torch.cuda.Stream(device=dev)
        self._stage = None                              # (x, ensemble, y) of the set being filled, allocated on first use
        self._staged = torch.cuda.Event()               # the staged batch has landed
        self._stage_free = torch.cuda.Event()           # the set being filled is no longer read by a running step
        self._has_staged = False
        self._alt_graph = None                          # graph of the step reading the other input set
        self._graph = None
        self.kernels_per_step = None        # librc launches inside one fwd+bwd (counted at capture)

    # ------------------------------------------------------------------ data-parallel gradient exchange
    def _setup_peer_exchange(self, total: int, dev):
        """Symmetric (peer-mapped) gradient buffer + barrier flags through torch's symmetric-memory rendezvous (CUDA
        peer mappings inside one box).  Returns the gradient buffer, or None when peer memory is not available
        (the step then uses one NCCL all-reduce, RC_DP_EXCHANGE=nccl forces that)."""
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.pg if self.pg is not None else torch.distributed.group.WORLD
            rank = torch.distributed.get_rank(group)
            flat = symm.empty(total, dtype=torch.float32, device=dev)
            flags = symm.empty(2 * 16, dtype=torch.int32, device=dev)
            flags2 = symm.empty(2 * 16, dtype=torch.int32, device=dev)      # align_ranks(): a barrier of its own
            flat.zero_()
            flags.zero_()
            flags2.zero_()
            h_flat = symm.rendezvous(flat, group)
            h_flags = symm.rendezvous(flags, group)
            h_flags2 = symm.rendezvous(flags2, group)
            torch.cuda.synchronize(dev)
            # flag protocol of the exchange kernel: device-scope fences by default (rc_p2p.cu), RC_P2P_STRICT=1 keeps
            # release / acquire at system scope
            _lib.check(_lib.lib().rc_p2p_flag_scope(0 if os.environ.get("RC_P2P_STRICT", "0") == "1" else 1), "rc_p2p_flag_scope")
            torch.distributed.barrier(group)          # every rank's flags are zero before anyone signals
            self.p2p = {"rank": rank, "grads": h_flat.buffer_ptrs_dev, "flags": h_flags.buffer_ptrs_dev,
                        "epochs": torch.zeros(2, dtype=torch.int32, device=dev),
                        "timed_out": torch.zeros(1, dtype=torch.int32, device=dev), "keep": (flat, flags, h_flat, h_flags, flags2, h_flags2),
                        "align_flags": h_flags2.buffer_ptrs_dev, "align_epochs": torch.zeros(2, dtype=torch.int32, device=dev)}
            return flat
        except Exception as exc:                      # no symmetric memory on this box / torch build: NCCL path
            import logging
            logging.getLogger(__name__).warning("peer-memory gradient exchange unavailable (%s): using NCCL all-reduce", exc)
            self.p2p = None
            return None

    # ------------------------------------------------------------------ parameter dictionaries for kernels.py
    def _bind(self, model):
        P = dict(model.named_parameters())
        G = self.grads
        B = dict(model.named_buffers())

        def pick(src, mapping):
            return {k: src[v] for k, v in mapping.items()}
        ds_map = {"phi0_w": "deepset.phi.0.weight", "phi0_b": "deepset.phi.0.bias", "phi2_w": "deepset.phi.2.weight",
                  "phi2_b": "deepset.phi.2.bias", "rho0_w": "deepset.rho.0.weight", "rho0_b": "deepset.rho.0.bias",
                  "rho2_w": "deepset.rho.2.weight", "rho2_b": "deepset.rho.2.bias"}
        dr_map = {"dimred_w": "dim_red.weight", "dimred_b": "dim_red.bias"}
        hd_map = {"aggr_w": "aggr.weight", "aggr_b": "aggr.bias"}
        layers = []
        for i in range(len(model.conv.convolutions)):
            pre = f"conv.convolutions.{i}."
            lmap = {"eps": pre + "eps", "lin_w": pre + "lin.weight", "lin_b": pre + "lin.bias", "nn0_w": pre + "nn.0.weight",
                    "nn0_b": pre + "nn.0.bias", "bn_w": pre + "nn.1.weight", "bn_b": pre + "nn.1.bias",
                    "nn3_w": pre + "nn.3.weight", "nn3_b": pre + "nn.3.bias"}
            lp = pick(P, lmap)
            lp.update(bn_rm=B[pre + "nn.1.running_mean"], bn_rv=B[pre + "nn.1.running_var"],
                      bn_nbt=B[pre + "nn.1.num_batches_tracked"])
            layers.append((lp, pick(G, lmap)))
        return {"ds": (pick(P, ds_map), pick(G, ds_map)), "dr": (pick(P, dr_map), pick(G, dr_map)),
                "layers": layers, "head": (pick(P, hd_map), pick(G, hd_map))}

    # ------------------------------------------------------------------ one forward + loss + backward (no optimiser)
    @property
    def _opt_in_graph(self) -> bool:
        """The optimiser (and the peer-memory gradient exchange) is part of the captured step; only the NCCL all-reduce
        fallback runs eagerly after the replay."""
        return self.world == 1 or self.p2p is not None

    def _fwd_bwd(self):
        """One whole training step on the current stream (+ the side stream): forward, CRPS, backward and - unless the
        gradients travel through NCCL - the gradient exchange, AdamW and the running loss sum.  This is what the CUDA
        graph holds."""
        K.SIDE.stream, K.SIDE.alt = self._side, self._side2
        try:
            self._fwd_bwd_body()
            K.join_side()
        finally:
            K.SIDE.stream = K.SIDE.alt = None
        if self._opt_in_graph:
            self._optimizer()
            self.loss_sum.add_(self.loss)

    def _fwd_bwd_body(self):
        blk = self._blocks
        Pd, Gd = blk["ds"]
        K.dimred_prepack(blk["dr"][0], self.feats, self.x)   # side stream, overlaps the DeepSets kernels
        emb, s_ds = K.deepsets_fwd(Pd, self.ens, bf16=(getattr(self.model.deepset, "compute_dtype", "fp32") == "bf16"))
        Pr, Gr = blk["dr"]
        node, s_dr = K.dimred_fwd(Pr, self.x, emb)
        saved = []
        h = node
        for i, (Pl, _) in enumerate(blk["layers"]):
            h, s = K.gine_layer_fwd(Pl, h, self.graph, first=(i == 0), training=True)
            saved.append(s)
        Ph, Gh = blk["head"]
        if FUSE_HEAD and K.head_crps_blocks(h.shape[0], h.shape[1]) > 0:
            # head Linear + links + CRPS + their backward: one launch instead of three on the chain and one beside it
            _, d, _ = K.head_crps_fwd_bwd(Ph, h, self.y, self.kind, Gh, u=self.u_fixed, xi=self.xi, t=self.t, loss_out=self.loss,
                                          flush=False)         # (reduced with the last GINE layer's partials)
        else:
            raw, s_h = K.head_fwd(Ph, h)
            _, d_raw, _ = K.crps_fwd_bwd(raw, self.y, self.kind, raw_input=True, u=self.u_fixed, xi=self.xi, t=self.t,
                                         loss_out=self.loss)
            d = K.head_bwd(Ph, s_h, d_raw, Gh, flush=False)
        for i in reversed(range(len(blk["layers"]))):
            Pl, Gl = blk["layers"][i]
            d = K.gine_layer_bwd(Pl, saved[i], self.graph, d, Gl, first=(i == 0), training=True)
        d_emb = K.dimred_bwd(Pr, s_dr, d, Gr, flush=False)        # (reduced with the DeepSets block's partials)
        K.deepsets_bwd(Pd, s_ds, d_emb, Gd)

    def _optimizer(self):
        if self.p2p is not None and os.environ.get("RC_DP_DEBUG_LOCAL_ADAMW", "0") == "1":
            return self._adamw_call()         # (measurement only: peer-mapped gradient buffer, no exchange)
        if self.p2p is not None:
            # ONE kernel: wait for every rank's gradients, sum them from peer memory in rank order, AdamW; it ends once every
            # peer is done reading this rank's gradients, so the next step needs no handshake before its backward
            L, P, st = _lib.lib(), self.p2p, torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(L.rc_p2p_step(self.flat_p.data_ptr(), P["grads"], P["flags"], P["epochs"].data_ptr(), P["rank"], self.world,
                                     self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.step_count.data_ptr(), self.n_params,
                                     self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, P["timed_out"].data_ptr(), st),
                       "rc_p2p_step")
            return
        if self.world > 1:
            torch.distributed.all_reduce(self.flat_g, op=torch.distributed.ReduceOp.SUM, group=self.pg)
        self._adamw_call()

    def _adamw_call(self):
        _lib.check(_lib.lib().rc_adamw_step(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(),
                                            self.exp_avg_sq.data_ptr(), self.step_count.data_ptr(), self.n_params, self.lr,
                                            self.betas[0], self.betas[1], self.eps, self.weight_decay, 1.0 / self.world,
                                            torch.cuda.current_stream(self.device).cuda_stream), "rc_adamw_step")

    def capture(self):
        """Build the step: warm up on a side stream and capture fwd+bwd in a CUDA graph.  BatchNorm running statistics
        and Adam state are restored afterwards, so capture has no training effect."""
        snap = {k: v.clone() for k, v in self.model.state_dict().items()}
        opt_snap = (self.exp_avg.clone(), self.exp_avg_sq.clone(), self.step_count.clone())
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._fwd_bwd()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        if self.use_cuda_graph:
            self._graph = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            with torch.cuda.graph(self._graph):
                self._fwd_bwd()
            self.kernels_per_step = _lib.launch_count() - before
        else:
            before = _lib.launch_count()
            self._fwd_bwd()
            self.kernels_per_step = _lib.launch_count() - before
        torch.cuda.synchronize(self.device)
        with torch.no_grad():
            for k, v in self.model.state_dict().items():
                v.copy_(snap[k])
            self.exp_avg.copy_(opt_snap[0])
            self.exp_avg_sq.copy_(opt_snap[1])
            self.step_count.copy_(opt_snap[2])       # (the peer-exchange epoch is NOT restored: it advances in lock step on all ranks)
        self.loss_sum.zero_()
        self.flat_g.zero_()
        if self.p2p is not None:
            torch.cuda.synchronize(self.device)
            torch.distributed.barrier(self.pg if self.pg is not None else torch.distributed.group.WORLD)
        return self

    # ------------------------------------------------------------------ public
    def load_batch(self, x, ensemble, y, non_blocking: bool = True):
        """Copy one batch (host pinned or device tensors) into the static input buffers."""
        self.x.copy_(x, non_blocking=non_blocking)
        self.ens.copy_(ensemble, non_blocking=non_blocking)
        self.y.copy_(y, non_blocking=non_blocking)

    def load_dates(self, split, dates: torch.Tensor):
        """Build the step's batch from a device-resident split (utils.dataset.DeviceSplit): `dates` is a device int64
        tensor of B forecast-date indices; one gather kernel, no host copy."""
        b = int(dates.numel())
        n = split.num_stations
        if b * n != self.m or split.x.shape[2] != self.feats or split.ensemble.shape[2] != self.members:
            raise _lib.RcError(f"load_dates: {b} dates of {n} stations do not make the captured batch of {self.m} nodes")
        if getattr(self, "_bad_date", None) is None:
            self._bad_date = torch.zeros(1, dtype=torch.int32, device=self.device)
        _lib.check(_lib.lib().rc_gather_dates(split.x.data_ptr(), split.ensemble.data_ptr(), split.y.data_ptr(),
                                              dates.data_ptr(), b, len(split), n * self.feats, n * self.members * self.feats, n,
                                              self.x.data_ptr(), self.ens.data_ptr(), self.y.data_ptr(), self._bad_date.data_ptr(),
                                              torch.cuda.current_stream(self.device).cuda_stream), "rc_gather_dates")

    def begin_epoch(self, split, batches: torch.Tensor):
        """Start an epoch over a device-resident split with NO host work per step: `batches` (int64 [n_batches, B], any
        device) is the epoch's order of full batches.  It goes into a device-side table, the epoch base is pinned to the
        optimiser's step counter, and every `step_resident()` replays ONE graph = rc_gather_dates_step (batch number
        step_count - base) + the captured training step.  (train.py:55-62: DataLoader iteration + collate + H2D per step.)"""
        if self._graph is None and self.use_cuda_graph:
            self.capture()
        if self._graph is None:
            raise _lib.RcError("begin_epoch needs an engine with CUDA graphs (use load_dates + step otherwise)")
        n, b = int(batches.shape[0]), int(batches.shape[1])
        stations = split.num_stations
        if b * stations != self.m or split.x.shape[2] != self.feats or split.ensemble.shape[2] != self.members:
            raise _lib.RcError(f"begin_epoch: {b} dates of {stations} stations do not make the captured batch of {self.m} nodes")
        res = getattr(self, "_res", None)
        if res is None or res["split"] is not split or res["order"].shape[0] < n or res["inputs"] != (self.x.data_ptr(), self.ens.data_ptr()):
            cap = max(n, (len(split) + b - 1) // b)
            res = self._res = {"split": split, "order": torch.zeros((cap, b), dtype=torch.int64, device=self.device),
                               "base": torch.zeros(2, dtype=torch.int64, device=self.device), "graph": None,
                               "inputs": (self.x.data_ptr(), self.ens.data_ptr())}
            if getattr(self, "_bad_date", None) is None:
                self._bad_date = torch.zeros(1, dtype=torch.int32, device=self.device)
        res["order"][:n].copy_(batches, non_blocking=True)
        res["base"][:1].copy_(self.step_count.reshape(1))      # [step counter at the start of the epoch, batches in the epoch]
        res["base"][1:].fill_(n)
        if res["graph"] is None:
            self._gather_step()                          # (first launch of this kernel variant outside the capture)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self._graph.pool()):
                self._gather_step()
                self._fwd_bwd()
            res["graph"] = g

    def _gather_step(self):
        res, split = self._res, self._res["split"]
        n = split.num_stations
        _lib.check(_lib.lib().rc_gather_dates_step(split.x.data_ptr(), split.ensemble.data_ptr(), split.y.data_ptr(),
                                                   res["order"].data_ptr(), res["order"].shape[0], self.step_count.data_ptr(),
                                                   res["base"].data_ptr(), res["order"].shape[1], len(split), n * self.feats,
                                                   n * self.members * self.feats, n, self.x.data_ptr(), self.ens.data_ptr(),
                                                   self.y.data_ptr(), self._bad_date.data_ptr(),
                                                   torch.cuda.current_stream(self.device).cuda_stream), "rc_gather_dates_step")

    def step_resident(self):
        """One training step on the next batch of the epoch started by `begin_epoch` (one graph replay)."""
        self._res["graph"].replay()
        if not self._opt_in_graph:
            self._optimizer()
            self.loss_sum.add_(self.loss)
        return self.loss

    def check_dates(self):
        """Raise if a gather saw a date outside the split or a batch number outside the epoch (synchronises the device)."""
        bad = int(self._bad_date.item()) if getattr(self, "_bad_date", None) is not None else 0
        if bad:
            self._bad_date.zero_()
            raise _lib.RcError("a batch gather read a date index outside the split" if bad == 1 else
                               "step_resident ran past the batches given to begin_epoch")

    def align_ranks(self):
        """A device-side barrier of the ranks on the current stream (peer-memory flags of its own, one tiny kernel, no host
        synchronisation): every rank leaves it within a flag round trip of the last one's arrival.  bench.py calls it
        between its un-timed L2 flush and a timed step; without peer memory it falls back to a one-element all-reduce."""
        if self.p2p is not None:
            P = self.p2p
            _lib.check(_lib.lib().rc_p2p_barrier(P["align_flags"], P["align_epochs"].data_ptr(), P["rank"], self.world, 0,
                                                 P["timed_out"].data_ptr(), torch.cuda.current_stream(self.device).cuda_stream),
                       "rc_p2p_barrier")
        elif self.world > 1:
            if getattr(self, "_align", None) is None:
                self._align = torch.zeros(1, device=self.device)
            torch.distributed.all_reduce(self._align, group=self.pg)

    def check_peers(self):
        """Raise if a peer-memory barrier gave up waiting for another rank (synchronises the device)."""
        if self.p2p is not None and int(self.p2p["timed_out"].item()) != 0:
            raise _lib.RcError("a rank did not reach the gradient-exchange barrier within 10 s")

    def prefetch(self, x, ensemble, y):
        """Start the host -> device copy of the NEXT batch (pinned host tensors) on the copy stream; it overlaps the
        step that is running.  `take_prefetched` moves it into the step's inputs."""
        if self._stage is None:
            self._stage = tuple(torch.empty_like(t) for t in (self.x, self.ens, self.y))
            self._stage_free.record(torch.cuda.current_stream(self.device))
        self._copy.wait_event(self._stage_free)
        with torch.cuda.stream(self._copy):
            for dst, src in zip(self._stage, (x, ensemble, y)):
                dst.copy_(src, non_blocking=True)
            self._staged.record(self._copy)
        self._has_staged = True

    def wait_prefetch(self):
        """Make the current stream wait for the copy started by `prefetch` (e.g. before an end-of-step timestamp)."""
        if self._has_staged:
            torch.cuda.current_stream(self.device).wait_event(self._staged)

    def take_prefetched(self):
        """Make the prefetched batch the step's input.  Graph mode: flip to the input set it was copied into (its own
        graph is captured on first use); otherwise a device -> device copy on the current stream."""
        if not self._has_staged:
            raise _lib.RcError("take_prefetched without a prefetch")
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._staged)
        self._has_staged = False
        if self._graph is not None:
            # the current inputs become the set the next prefetch fills, once the step that reads them has run
            # (step() records _stage_free after its replay)
            (self.x, self.ens, self.y), self._stage = self._stage, (self.x, self.ens, self.y)
            self._graph, self._alt_graph = self._alt_graph, self._graph
            if self._graph is None:
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph, pool=self._alt_graph.pool()):
                    self._fwd_bwd()
            # the set just left was read by the replay already on this stream: it may be refilled once that has run
            # (recorded here, not after the coming replay, so that `prefetch` may follow `step` and still overlap it)
            self._stage_free.record(cur)
            return
        for dst, src in zip((self.x, self.ens, self.y), self._stage):
            dst.copy_(src, non_blocking=True)
        self._stage_free.record(cur)

    def step(self):
        """One training step on the batch in the static buffers; returns the device-resident loss (float64 [1])."""
        if self._graph is None and self.use_cuda_graph:
            self.capture()
        if self._graph is not None:
            self._graph.replay()
        else:
            self._fwd_bwd()
        if not self._opt_in_graph:                    # NCCL fallback: all-reduce + AdamW after the captured fwd / bwd
            self._optimizer()
            self.loss_sum.add_(self.loss)
        return self.loss

    def step_eager(self, x, ensemble, y, graph: StationGraph):
        """One training step on a batch of a DIFFERENT size than the captured one (the ragged last batch of an epoch,
        train.py:61-71): the same kernel schedule issued eagerly on `graph` / the given device tensors."""
        keep = (self.x, self.ens, self.y, self.graph, self.m)
        self.x, self.ens, self.y = (_lib.f32c(t) for t in (x, ensemble, y.reshape(-1)))
        self.graph, self.m = graph.to(self.device), int(x.shape[0])
        try:
            self._fwd_bwd()
            if not self._opt_in_graph:
                self._optimizer()
                self.loss_sum.add_(self.loss)
        finally:
            self.x, self.ens, self.y, self.graph, self.m = keep
        return self.loss

    def step_emulated_ranks(self, batches):
        """One data-parallel step of G = len(batches) ranks emulated on ONE GPU (never spin-wait across
        ranks on one device): every micro-batch runs forward / CRPS / backward with its own BatchNorm statistics and its
        own valid-node mean, the G flat gradients are summed in rank order, and ONE AdamW step takes their mean
        (grad_scale = 1/G) - the arithmetic of rc_p2p_step / all-reduce + rc_adamw_step.  Returns the G losses and the
        mean gradients by parameter name."""
        total = torch.zeros_like(self.flat_g)
        losses = []
        for x, ens, y in batches:
            self.load_batch(x, ens, y)
            K.SIDE.stream, K.SIDE.alt = self._side, self._side2
            try:
                self._fwd_bwd_body()
                K.join_side()
            finally:
                K.SIDE.stream = K.SIDE.alt = None
            total += self.flat_g
            losses.append(self.loss.clone())
        _lib.check(_lib.lib().rc_adamw_step(self.flat_p.data_ptr(), total.data_ptr(), self.exp_avg.data_ptr(),
                                            self.exp_avg_sq.data_ptr(), self.step_count.data_ptr(), self.n_params, self.lr,
                                            self.betas[0], self.betas[1], self.eps, self.weight_decay, 1.0 / len(batches),
                                            torch.cuda.current_stream(self.device).cuda_stream), "rc_adamw_step")
        mean = total / len(batches)
        grads = {k: mean[v.data_ptr() // 4 - self.flat_g.data_ptr() // 4:][:v.numel()].view(v.shape).clone() for k, v in self.grads.items()}
        return torch.cat(losses), grads

    @property
    def launches_per_step(self) -> int:
        """librc kernel launches per step."""
        if self._opt_in_graph:                        # AdamW / peer exchange are inside the counted step
            return int(self.kernels_per_step or 0)
        opt = 1                                       # AdamW after the NCCL all-reduce
        return int(self.kernels_per_step or 0) + opt
