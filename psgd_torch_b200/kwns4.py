"""KWNS4 -- drop-in for the reference's torch.optim wrappers, re-pointed at the B200 engine.

    KWNS4         mirrors /root/reference/wrapped_as_torch_optimizer_for_ddp.py:4-176      (single GPU and DDP)
    KWNS4DTensor  mirrors /root/reference/wrapped_as_torch_optimizer_for_dtensor.py:4-185  (FSDP2 / DTensor shards)

Same constructor arguments, defaults, asserts, param_groups keys and per-parameter state keys ("QL", "exprs", "step",
"ema") with the same tensor layouts, so a reference checkpoint's tensors load.  Same RNG discipline: private CPU+CUDA
generator states are swapped in around step() (ddp.py:100-104,172-176) and all draws happen in the reference's order, so
DDP replicas draw identical numbers and stay consistent.

What changes is where the arithmetic runs: the per-parameter loop body (ddp.py:117-157) is
    psgd_kwns4_head  (weight decay + cast + momentum EMA, one pass)
    psgd_kron_whiten_q0p5eq1p5_update
    psgd_kron_precond_grad (sum of squares for the clipping rule fused into the last product)
    psgd_kwns4_tail  (clip + clamp + parameter update, one pass, no host synchronisation -- the reference's
                      `if avg_amp > max_avg_amp` at ddp.py:154 stalls the host once per parameter)
"""
import ctypes as C

import torch

from . import _lib
from . import psgd


class KWNS4(torch.optim.Optimizer):
    """Kronecker-product whitening preconditioner fitted with online Newton-Schulz iteration (dQ = Q^0.5 E Q^1.5).
    See the reference docstring (ddp.py:5-23) for the meaning of every hyper-parameter; they are kept verbatim."""

    def __init__(
            self,
            params,
            whiten_grad=False,
            preconditioner_max_size=float("inf"),
            preconditioner_max_skew=1.0,
            preconditioner_init_scale=1.0,
            lr_params=2e-4,
            lr_preconditioner=0.5,
            betaL=0.9,
            damping=1e-9,
            momentum=0.9,
            weight_decay=0.05,
            decoupled_weight_decay=True,
            grad_clip_max_amps=(2.0, 10.0),
            preconditioner_update_probability=1.0,
            preconditioner_dtype: torch.dtype | None = torch.bfloat16,
            update_preconditioner_first=True,
            resync_every=1000_000,
            shard_preconditioners=False,
            batch_same_shape=False,
            comm_sms=0,
            exchange="all_gather",
    ):
        # ddp.py:45-62, verbatim
        assert whiten_grad in (False, True)
        assert preconditioner_max_size >= 0.0
        assert preconditioner_max_skew >= 0.0
        assert preconditioner_init_scale > 0.0
        assert lr_params > 0.0
        assert 0.0 < lr_preconditioner < 1.0
        assert 0.0 <= betaL <= 1.0
        assert damping >= 0.0
        assert 0.0 <= momentum < 1.0
        assert weight_decay >= 0.0
        assert decoupled_weight_decay in (False, True)
        assert grad_clip_max_amps[1] >= grad_clip_max_amps[0] >= 1.0
        assert 0.0 < preconditioner_update_probability <= 1.0
        assert preconditioner_dtype in (None, torch.bfloat16, torch.float32)
        assert update_preconditioner_first in (False, True)
        assert resync_every > 0
        if not whiten_grad:
            assert momentum > 0.0, "Cannot whiten momentum if momentum setting is zero."

        defaults = {
            "whiten_grad": whiten_grad,
            "preconditioner_max_size": preconditioner_max_size,
            "preconditioner_max_skew": preconditioner_max_skew,
            "preconditioner_init_scale": preconditioner_init_scale,
            "lr_params": lr_params,
            "lr_preconditioner": lr_preconditioner,
            "betaL": betaL,
            "damping": damping,
            "momentum": momentum,
            "weight_decay": weight_decay,
            "decoupled_weight_decay": decoupled_weight_decay,
            "grad_clip_max_amps": grad_clip_max_amps,
            "preconditioner_update_probability": preconditioner_update_probability,
            "preconditioner_dtype": preconditioner_dtype,
            "update_preconditioner_first": update_preconditioner_first,
            "resync_every": resync_every,
        }
        super().__init__(params, defaults)

        # Not in the reference (which replicates all of the optimizer work on every rank, ddp.py:88-96): owner-computes sharding of the
        # per-parameter preconditioners (BASELINE configs[3], SURVEY.md 8e "training-correct mode").  Every parameter is owned by one rank,
        # which alone keeps its (Q, L, ema) and runs its update + apply; the updated parameter is then broadcast to the other ranks
        # (NCCL).  Off by default so that the class stays a drop-in with the reference's replicated semantics.
        self.shard_preconditioners = bool(shard_preconditioners)
        # Not in the reference either: parameters of one shape share batched engine calls (psgd.*_batched: grouped tcgen05 launches, one
        # norm-bound launch per batch).  The random draws then happen batch by batch instead of parameter by parameter, so it is off by
        # default; a transformer's repeated layers are where it pays (64 k/v projections, 65 norm vectors, pairs of MLP matrices).
        self.batch_same_shape = bool(batch_same_shape)
        self.batch_numel_cap = 128 * 1024 * 1024     # elements of gradient per batched call (workspace grows with it)
        # sharded mode: SMs left free for the NCCL broadcast kernels that run beside the engine's persistent kernels (_lib.set_sm_limit)
        self.comm_sms = int(comm_sms)
        self._comm_sms_applied = False
        # sharded + batched mode: how the updated parameters reach the other ranks.  "all_gather": every rank packs the batch it has
        # just finished into a flat buffer and ONE all-gather per round moves all ranks' batches at once (every rank sends and receives
        # concurrently, large messages); "broadcast": one NCCL broadcast per parameter from its owner (roots take turns, so only one
        # rank sends at a time: measured 16 GB in 120 ms at 8 GPUs / 8 channels, profiles/r02_bench_kwns4_n8.json).
        # "p2p" (CUDA only, one node): an owner pushes each updated parameter with peer-to-peer cudaMemcpyAsync on side streams -- copy
        # engines over NVLink / NVSwitch, no SMs, beside the engine's kernels -- into staging buffers it allocated on the peers' GPUs; the
        # peers map those buffers through CUDA IPC and copy the parameters out at the end of step().  One tiny all-reduce at the start
        # of step() (the peers are done with the previous contents) and one at its end (all pushes have landed) are the only collectives.
        assert exchange in ("all_gather", "broadcast", "p2p")
        self.exchange = exchange
        self._xbuf = None
        self._comm_stream = None
        self._peer_views = None
        self._peer_streams = None
        self._owner = None
        self.dQ = "Q0.5EQ1.5"  # ddp.py:84-86
        self.update_precond = psgd.update_precond_kron_whiten_q0p5eq1p5
        self.precond_grad = psgd.precond_grad_kron
        self._sumsq = {}
        self._init_rng_sync()

    # ---- RNG discipline (ddp.py:88-96) ----
    def _needs_rng_sync(self):
        return torch.distributed.is_available() and torch.distributed.is_initialized()

    def _init_rng_sync(self):
        self.is_distributed = self._needs_rng_sync()
        if self.is_distributed:
            on_gpu = torch.distributed.get_backend() != "gloo"  # the reference assumes nccl; gloo keeps CPU tensors (tests)
            state = torch.get_rng_state()
            state = state.cuda() if on_gpu else state
            torch.distributed.broadcast(state, src=0)
            self.cpu_rng_state = state.cpu()
            if torch.cuda.is_available():
                state = torch.cuda.get_rng_state()
                state = state.cuda() if on_gpu else state
                torch.distributed.broadcast(state, src=0)
                self.cuda_rng_state = state.cpu()
            else:
                self.cuda_rng_state = None

    def _rng_enter(self):
        """ddp.py:100-104: swap the private (synchronised) generator states in; returns the caller's states."""
        if not self.is_distributed:
            return None
        ext = [torch.get_rng_state(), None]
        torch.set_rng_state(self.cpu_rng_state)
        if self.cuda_rng_state is not None:
            ext[1] = torch.cuda.get_rng_state()
            torch.cuda.set_rng_state(self.cuda_rng_state)
        return ext

    def _rng_exit(self, ext):
        """ddp.py:172-176"""
        if ext is None:
            return
        self.cpu_rng_state = torch.get_rng_state()
        torch.set_rng_state(ext[0])
        if self.cuda_rng_state is not None:
            self.cuda_rng_state = torch.cuda.get_rng_state()
            torch.cuda.set_rng_state(ext[1])

    # ---- owner-computes sharding (shard_preconditioners=True) ----
    def _sharding_active(self):
        return self.shard_preconditioners and self.is_distributed and torch.distributed.get_world_size() > 1

    def _assign_owners(self):
        """LPT partition of all parameters on the cost model of SURVEY.md 8d; identical on every rank (shapes only)."""
        from . import partition
        plist, costs = [], []
        for group in self.param_groups:
            for p in group["params"]:
                shp = tuple(self._local(p).squeeze().shape)
                numel = 1
                for s_ in shp:
                    numel *= s_
                dense = [not (s_ <= 1 or s_ > group["preconditioner_max_size"] or s_ * s_ > group["preconditioner_max_skew"] * numel) for s_ in shp]
                if len(shp) == 2:
                    c = partition.kron_unit_cost(shp[0], shp[1], dense[0], dense[1])
                else:   # 0/1-D and order >= 3: a pass-count estimate is enough for balancing
                    c = 40.0 * numel * 2 * 250.0 + sum(8.0 * s_ ** 3 + 6.0 * s_ * numel for s_, dn in zip(shp, dense) if dn)
                plist.append(p)
                costs.append(c)
        owners = partition.owner_of(costs, torch.distributed.get_world_size())
        self._owner = {id(p): r for p, r in zip(plist, owners)}
        # the only draws every rank must agree on are the per-group update coins (ddp.py:110): a dedicated generator seeded from the
        # synchronised private state, because each rank now consumes a different amount of the main streams
        seed = int(torch.frombuffer(bytearray(self.cpu_rng_state[:8].numpy().tobytes()), dtype=torch.int64)[0]) & 0x7FFFFFFF
        self._coin_gen = torch.Generator().manual_seed(seed)
        if getattr(self, "_coin_state", None) is not None:     # resumed from a checkpoint
            self._coin_gen.set_state(self._coin_state)
            self._coin_state = None

    # ---- hooks overridden by the DTensor variant ----
    def _local(self, t):
        return t

    def _resync(self, p, state, group, momentum):
        # ddp.py:163-170
        if self.is_distributed and (state["step"] % group["resync_every"] == 0):
            torch.distributed.broadcast(p, src=0)
            if momentum > 0.0:
                torch.distributed.broadcast(state["ema"], src=0)
            for q, ell in zip(*state["QL"]):
                torch.distributed.broadcast(q, src=0)
                torch.distributed.broadcast(ell, src=0)

    def _sumsq_buf(self, device):
        b = self._sumsq.get(device)
        if b is None:
            b = torch.zeros(1, dtype=torch.float32, device=device)
            self._sumsq[device] = b
        return b

    # ---- checkpoints ----
    _EXPRS_TAG = "psgd_exprs/v1"

    def _gather_sharded_state(self):
        """shard_preconditioners=True: every rank holds the state of the parameters it owns only.  Collective (all ranks call
        state_dict()): the owner broadcasts (Q, L, ema, step) of each of its parameters, so that every rank returns the full state and the
        usual rank-0-only checkpoint is complete."""
        dist = torch.distributed
        me = dist.get_rank()
        full = {}
        plist = [p for group in self.param_groups for p in group["params"]]
        has = torch.zeros(len(plist), dtype=torch.int64, device=self._local(plist[0]).device if plist else "cpu")
        for i, p in enumerate(plist):
            if self._owner.get(id(p)) == me and len(self.state.get(p, {})) > 0:
                has[i] = 1 + self.state[p]["step"]
        dist.all_reduce(has, op=dist.ReduceOp.MAX)
        has = has.tolist()
        index_of = {id(p): i for i, p in enumerate(plist)}
        for group in self.param_groups:
            for p in group["params"]:
                i = index_of[id(p)]
                if has[i] == 0:
                    continue
                owner = self._owner[id(p)]
                lp = self._local(p)
                pre = group["preconditioner_dtype"] or lp.dtype
                if owner == me:
                    st = self.state[p]
                else:
                    shape = lp.squeeze().shape
                    QL, exprs = psgd.init_kron(torch.empty(shape, dtype=pre, device=lp.device), Scale=group["preconditioner_init_scale"],
                                               max_size=group["preconditioner_max_size"], max_skew=group["preconditioner_max_skew"], dQ=self.dQ)
                    st = {"QL": QL, "exprs": exprs, "step": has[i] - 1,
                          "ema": None if group["momentum"] == 0.0 else torch.zeros(shape, dtype=pre, device=lp.device)}
                for t in list(st["QL"][0]) + list(st["QL"][1]) + ([st["ema"]] if st["ema"] is not None else []):
                    dist.broadcast(t, src=owner)
                full[p] = st
        return full

    def state_dict(self):
        """The reference's state_dict (keys "QL", "step", "ema" per parameter) with three differences:
          * `exprs` (callables) is replaced by a plain tag -- it is rebuilt from the factor shapes on load, so the checkpoint holds tensors
            and python scalars only (torch.load(weights_only=True) accepts it);
          * the private CPU / CUDA generator states of a distributed run ride in param_groups[0]["psgd_rng"] (ddp.py:92,96 keep them as
            attributes and forget them in checkpoints, so a resumed reference run draws different numbers than an uninterrupted one);
          * with shard_preconditioners=True the call is COLLECTIVE (every rank must make it): owned states are broadcast so that every
            rank returns the complete state."""
        saved = None
        if self._sharding_active() and self._owner is not None:
            saved = dict(self.state)
            for p, st in self._gather_sharded_state().items():
                self.state[p] = st
        try:
            sd = super().state_dict()
        finally:
            if saved is not None:
                self.state.clear()
                self.state.update(saved)
        sd["state"] = {k: {kk: (self._EXPRS_TAG if kk == "exprs" else vv) for kk, vv in st.items()} for k, st in sd["state"].items()}
        if getattr(self, "is_distributed", False):
            sd["param_groups"] = [dict(g) for g in sd["param_groups"]]
            rng = {"cpu": self.cpu_rng_state.clone(), "cuda": None if self.cuda_rng_state is None else self.cuda_rng_state.clone()}
            if getattr(self, "_coin_gen", None) is not None:
                rng["coin"] = self._coin_gen.get_state().clone()
            sd["param_groups"][0]["psgd_rng"] = rng
        return sd

    def load_state_dict(self, state_dict):
        """torch.optim.Optimizer.load_state_dict casts every floating state tensor to the PARAMETER's dtype: a bf16 preconditioner of an fp32
        parameter would come back as fp32, and -- worse -- the fp32 Lipschitz constants (and an fp32 preconditioner) of a bf16 parameter would
        be truncated to 8 mantissa bits.  Q, L and ema are therefore restored from the checkpoint's own tensors: Q / ema in the group's
        preconditioner_dtype, L in fp32 (psgd.py:96-98).  `exprs` is rebuilt from the factor shapes.  A checkpoint written by the reference
        wrapper (same keys, no RNG entry) loads unchanged; with shard_preconditioners=True every rank loads the full state and keeps the
        parameters it owns."""
        raw_state = state_dict["state"]
        state_dict = {"state": {k: dict(v) for k, v in raw_state.items()}, "param_groups": [dict(g) for g in state_dict["param_groups"]],
                      **{k: v for k, v in state_dict.items() if k not in ("state", "param_groups")}}
        rng = state_dict.pop("psgd_rng", None)                        # round-1 checkpoints kept it at the top level
        rng = state_dict["param_groups"][0].pop("psgd_rng", rng) if state_dict["param_groups"] else rng
        for st in state_dict["state"].values():
            st.pop("exprs", None)                                     # callables (reference) or the tag: rebuilt below
        super().load_state_dict(state_dict)
        if rng is not None and getattr(self, "is_distributed", False):
            self.cpu_rng_state = rng["cpu"].clone().cpu()
            if rng.get("cuda") is not None and self.cuda_rng_state is not None:
                self.cuda_rng_state = rng["cuda"].clone().cpu()
            if rng.get("coin") is not None:
                if getattr(self, "_coin_gen", None) is not None:
                    self._coin_gen.set_state(rng["coin"])
                else:
                    self._coin_state = rng["coin"]
        idx = 0
        me = torch.distributed.get_rank() if self._sharding_active() else 0
        if self._sharding_active() and self._owner is None:
            self._assign_owners()
        for group in self.param_groups:
            pd = group["preconditioner_dtype"]
            for p in group["params"]:
                raw = raw_state.get(idx, raw_state.get(str(idx)))
                idx += 1
                st = self.state.get(p)
                if not st or raw is None:
                    continue
                if self._sharding_active() and self._owner[id(p)] != me:
                    del self.state[p]                                 # another rank owns this parameter's preconditioner
                    continue
                lp = self._local(p)
                dt = pd or lp.dtype
                Qr, Lr = raw["QL"]
                st["QL"] = [[q.detach().to(device=lp.device, dtype=dt).contiguous().clone() for q in Qr],
                            [l.detach().to(device=lp.device, dtype=torch.float32).clone() for l in Lr]]
                st["ema"] = None if raw.get("ema") is None else raw["ema"].detach().to(device=lp.device, dtype=dt).contiguous().clone()
                st["exprs"] = psgd.exprs_for_state(st["QL"][0], self.dQ)

    # ---- one parameter's share of the loop body (ddp.py:117-143): state creation, weight decay + cast + EMA (one engine pass) ----
    def _head(self, p, group):
        lib = _lib.load_library()
        momentum = group["momentum"]
        wd, lr_params = group["weight_decay"], group["lr_params"]
        local_p = self._local(p)
        grad = self._local(p.grad)
        if not local_p.is_contiguous():
            raise _lib.EngineError("KWNS4 (B200 engine) needs contiguous parameters")
        grad = grad.contiguous()
        pre_dtype = group["preconditioner_dtype"] or grad.dtype
        sq_shape = grad.squeeze().shape  # ddp.py:124
        dev = grad.device
        h = _lib.handle_for(dev)
        state = self.state[p]
        if len(state) == 0:  # ddp.py:130-137
            QL, exprs = psgd.init_kron(torch.empty(sq_shape, dtype=pre_dtype, device=dev),
                                       Scale=group["preconditioner_init_scale"],
                                       max_size=group["preconditioner_max_size"],
                                       max_skew=group["preconditioner_max_skew"], dQ=self.dQ)
            state["QL"], state["exprs"] = QL, exprs
            state["step"] = 0
            state["ema"] = None if momentum == 0.0 else torch.zeros(sq_shape, dtype=pre_dtype, device=dev)
        t = state["step"]
        beta = min(t / (t + 1), momentum) if momentum > 0.0 else 0.0  # ddp.py:141
        need_g = group["whiten_grad"] or momentum == 0.0
        coupled = wd > 0.0 and not group["decoupled_weight_decay"]
        g_cast = torch.empty(sq_shape, dtype=pre_dtype, device=dev) if (need_g and (grad.dtype != pre_dtype or coupled)) else None
        # head: weight decay (ddp.py:117-122) + cast (125-127) + EMA (139-143), one pass
        rc = lib.psgd_kwns4_head(h, grad.numel(), _lib.ptr(local_p), _lib.dtype_code(local_p), _lib.ptr(grad),
                                 _lib.dtype_code(grad), float(wd), float(lr_params), int(group["decoupled_weight_decay"]),
                                 _lib.ptr(state["ema"]), _lib.ptr(g_cast), _lib._DTYPES[pre_dtype], float(beta),
                                 _lib.stream_ptr(dev))
        _lib.check(h, rc, "psgd_kwns4_head")
        state["step"] += 1
        g_pre = g_cast if g_cast is not None else grad.view(sq_shape)
        return {"p": p, "local_p": local_p, "state": state, "dev": dev, "h": h,
                "whiten": g_pre if group["whiten_grad"] else state["ema"],       # ddp.py:145
                "precond": g_pre if momentum == 0.0 else state["ema"]}           # ddp.py:150

    def _tail(self, job, hh, sumsq, group):
        """clip (ddp.py:153-156) + p -= lr*h (157), one pass, device-side branch"""
        lib = _lib.load_library()
        max_avg_amp, max_element_amp = group["grad_clip_max_amps"]
        lp = job["local_p"]
        rc = lib.psgd_kwns4_tail(job["h"], hh.numel(), hh.numel(), _lib.ptr(lp), _lib.dtype_code(lp), _lib.ptr(hh), _lib.dtype_code(hh),
                                 _lib.ptr(sumsq), float(max_avg_amp), float(max_element_amp), float(group["lr_params"]),
                                 _lib.stream_ptr(job["dev"]))
        _lib.check(job["h"], rc, "psgd_kwns4_tail")

    def _process(self, jobs, group, updateP_first, updateP_last):
        """update (if due) -> apply -> tail -> update (if due last) for one parameter or one same-shape batch of parameters."""
        kw = dict(lr=group["lr_preconditioner"], betaL=group["betaL"], damping=group["damping"])
        if len(jobs) == 1:
            job = jobs[0]
            st = job["state"]
            if updateP_first:  # ddp.py:146-148
                self.update_precond(st["QL"], st["exprs"], job["whiten"], **kw)
            sumsq = self._sumsq_buf(job["dev"])
            hh = self.precond_grad(st["QL"], st["exprs"], job["precond"], sumsq_out=sumsq)  # ddp.py:150-151
            self._tail(job, hh, sumsq, group)
            if updateP_last:  # ddp.py:159-161
                self.update_precond(st["QL"], st["exprs"], job["whiten"], **kw)
            return
        QLs = [j["state"]["QL"] for j in jobs]
        if updateP_first:
            psgd.update_precond_kron_whiten_q0p5eq1p5_batched(QLs, None, [j["whiten"] for j in jobs], **kw)
        sumsq = torch.empty(len(jobs), dtype=torch.float32, device=jobs[0]["dev"])
        hs = psgd.precond_grad_kron_batched(QLs, None, [j["precond"] for j in jobs], sumsq_out=sumsq)
        for i, (job, hh) in enumerate(zip(jobs, hs)):
            self._tail(job, hh, sumsq[i:i + 1], group)
        if updateP_last:
            psgd.update_precond_kron_whiten_q0p5eq1p5_batched(QLs, None, [j["whiten"] for j in jobs], **kw)

    @staticmethod
    def _batch_key(job):
        st = job["state"]
        g = job["whiten"]
        return (tuple(g.shape), g.dtype, tuple(q.dim() for q in st["QL"][0])) if g.dim() <= 2 else None

    @torch.no_grad()
    def step(self):
        external = self._rng_enter()
        sharded = self._sharding_active()
        if sharded and self._owner is None:
            self._assign_owners()
        if sharded and self.exchange == "p2p" and self._peer_views is None:
            self._setup_p2p()
        if sharded and self.comm_sms > 0 and not self._comm_sms_applied:
            total = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
            _lib.set_sm_limit(total - self.comm_sms)
            self._comm_sms_applied = True
        my_rank = torch.distributed.get_rank() if sharded else 0
        pending = []
        self._xbytes_step = 0        # bytes every rank receives through the all-gather exchange this step (padding included)
        p2p = sharded and self.exchange == "p2p"
        if p2p:
            torch.distributed.all_reduce(self._p2p_flag)    # every rank has finished reading its parameters (forward / backward)
        for group in self.param_groups:
            momentum = group["momentum"]
            coin = torch.rand([], generator=self._coin_gen) if sharded else torch.rand([])
            updateP_first, updateP_last = ((group["update_preconditioner_first"], not group["update_preconditioner_first"])
                                           if coin < group["preconditioner_update_probability"] else (False, False))
            # the parameters this rank works on, in parameter order, and (sharded) the order in which every rank meets the broadcasts
            mine, recv = [], []
            for p in group["params"]:
                local_p = self._local(p)
                if sharded:
                    # owner and receivers must take the same decision or the broadcasts dead-lock: the skip tests use what every rank
                    # sees alike (the parameter), never the local gradient, and a missing gradient on the owner is an error
                    if local_p.numel() == 0:
                        continue
                    owner = self._owner[id(p)]
                    if owner != my_rank:
                        recv.append((p, owner))
                        continue
                    if p.grad is None:
                        raise _lib.EngineError("shard_preconditioners=True: the owner rank has no gradient for a parameter the other ranks "
                                               "are waiting for (every parameter handed to KWNS4 must receive a gradient on every step)")
                if p.grad is None or self._local(p.grad).numel() == 0:  # ddp.py:114-115, dtensor.py:124-125
                    continue
                mine.append(p)
            if p2p:
                # owner computes and pushes; nothing to do for the parameters of other owners
                # largest batches first: the push of a batch overlaps the batches computed after it, so the step should end on a small one
                # (an lm_head-sized parameter is 1 GB x 7 peers = 10 ms of NVLink egress)
                batches = self._make_batches(mine) if self.batch_same_shape else [[p] for p in mine]
                batches.sort(key=lambda pl: sum(self._local(p).numel() * self._local(p).element_size() for p in pl), reverse=True)
                for plist in batches:
                    self._process([self._head(p, group) for p in plist], group, updateP_first, updateP_last)
                    self._push_round(plist)
                continue
            if not self.batch_same_shape:
                # the reference's order: one parameter at a time (ddp.py:112-161)
                if sharded:
                    order = [p for p in group["params"] if self._local(p).numel() > 0]
                    mine_set = set(id(p) for p in mine)
                    for p in order:
                        if id(p) in mine_set:
                            self._process([self._head(p, group)], group, updateP_first, updateP_last)
                        pending.append(torch.distributed.broadcast(self._local(p), src=self._owner[id(p)], async_op=True))
                else:
                    for p in mine:
                        job = self._head(p, group)
                        self._process([job], group, updateP_first, updateP_last)
                        self._resync(p, job["state"], group, momentum)
                continue
            # batched: same-shape parameters share engine calls (psgd.*_batched); not the reference's draw order
            batches = self._make_batches(mine)
            if not sharded:
                for plist in batches:
                    jobs = [self._head(p, group) for p in plist]
                    self._process(jobs, group, updateP_first, updateP_last)
                    for p, job in zip(plist, jobs):
                        self._resync(p, job["state"], group, momentum)
                continue
            # sharded + batched: every rank derives every rank's batch list from the shapes and walks the same global sequence
            world = torch.distributed.get_world_size()
            all_params = [p for p in group["params"] if self._local(p).numel() > 0]
            per_rank = [self._make_batches([p for p in all_params if self._owner[id(p)] == r]) for r in range(world)]
            if self.exchange == "all_gather":
                # rounds: round i holds the i-th largest batch of every rank (similar sizes per round -> little padding)
                def nbytes(plist):
                    return sum((self._local(p).numel() * self._local(p).element_size() + 15) // 16 * 16 for p in plist)
                for r in range(world):
                    per_rank[r].sort(key=nbytes, reverse=True)        # stable, shapes only: identical on every rank
                for i in range(max(len(b) for b in per_rank)):
                    plists = [per_rank[r][i] if i < len(per_rank[r]) else [] for r in range(world)]
                    if plists[my_rank]:
                        self._process([self._head(p, group) for p in plists[my_rank]], group, updateP_first, updateP_last)
                    self._exchange_round(plists, max(nbytes(pl) for pl in plists), my_rank, world)
                continue
            # one broadcast per parameter: round-robin over the owners' lists, computing the own batches and posting the broadcasts of all
            # of them in that order
            for i in range(max(len(b) for b in per_rank)):
                for r in range(world):
                    if i >= len(per_rank[r]):
                        continue
                    plist = per_rank[r][i]
                    if r == my_rank:
                        self._process([self._head(p, group) for p in plist], group, updateP_first, updateP_last)
                    for p in plist:
                        pending.append(torch.distributed.broadcast(self._local(p), src=r, async_op=True))
        for w in pending:
            w.wait()
        if self._comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        if p2p:
            for st in self._peer_streams.values():
                torch.cuda.current_stream().wait_stream(st)
            torch.distributed.all_reduce(self._p2p_flag)    # ... and every rank's pushes have landed
            self._unpack_p2p(my_rank)
        self._rng_exit(external)

    def _setup_p2p(self):
        """Collective, once.  Every rank allocates, ON EACH PEER'S GPU, a staging buffer for the parameters it owns (so that its pushes
        are plain peer-to-peer copies into its own allocation: 750 GB/s per direction on the copy engines, measured beside running
        GEMMs, tools/p2p_probe.py) and hands the peer a CUDA IPC handle of it; the peer maps it as device-local memory (3.2 TB/s copy-out,
        tools/p2p_probe2.py).  Writing THROUGH an IPC mapping of another GPU's memory is the slow direction here (26 GB/s), hence the
        staging.  One node, plain CUDA parameters."""
        from torch.multiprocessing import reductions
        dist = torch.distributed
        world, me = dist.get_world_size(), dist.get_rank()
        plist = [p for group in self.param_groups for p in group["params"] if self._local(p).numel() > 0]
        dev = self._local(plist[0]).device
        # layout of every owner's staging buffer: its parameters in parameter order, 16-byte aligned (shapes only: same on every rank)
        self._stage_off, sizes = {}, [0] * world
        for p in plist:
            r = self._owner[id(p)]
            lp = self._local(p)
            self._stage_off[id(p)] = sizes[r]
            sizes[r] += (lp.numel() * lp.element_size() + 15) // 16 * 16
        # rank k drives GPU devs[k] of this node (torchrun: LOCAL_RANK); every rank takes part in every collective below whatever fails
        devs = [None] * world
        dist.all_gather_object(devs, dev.index if dev.type == "cuda" else None)
        err, args = None, None
        try:
            if not all(self._local(p).is_cuda and self._local(p).is_contiguous() for p in plist):
                raise _lib.EngineError("exchange='p2p' needs contiguous CUDA parameters")
            if world > torch.cuda.device_count() or len(set(devs)) != world:
                raise _lib.EngineError("exchange='p2p' works inside one node, one visible GPU per rank")
            self._stage = {k: torch.empty(max(sizes[me], 16), dtype=torch.uint8, device=torch.device("cuda", devs[k]))
                           for k in range(world) if k != me}
            args = {k: reductions.reduce_tensor(t)[1] for k, t in self._stage.items()}
            lib = _lib.load_library()
            for k in self._stage:
                rc = lib.psgd_peer_enable(devs[k])
                if rc != 0:
                    raise _lib.EngineError(f"no peer access from GPU {dev.index} to GPU {devs[k]} ({rc})")
        except Exception as e:
            err = e
        gathered = [None] * world
        dist.all_gather_object(gathered, args)
        inbox = {}
        if err is None and all(g is not None for g in gathered):
            try:
                for r in range(world):
                    if r != me:
                        inbox[r] = reductions.rebuild_cuda_tensor(*gathered[r][me])      # rank r's staging buffer on MY GPU
                        assert inbox[r].device == dev
            except Exception as e:
                err = e
        ok = torch.tensor([0.0 if err is not None or any(g is None for g in gathered) else 1.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) != 1.0:
            raise _lib.EngineError(f"exchange='p2p': setting up the peer staging buffers failed on some rank ({err}); use exchange='all_gather'")
        self._inbox = inbox
        self._peer_views = True
        self._peer_streams = {k: torch.cuda.Stream() for k in self._stage}
        self._p2p_flag = torch.zeros(1, device=dev)
        self._p2p_plist = plist

    def _stage_view(self, buf, p):
        lp = self._local(p)
        off = self._stage_off[id(p)]
        return buf[off:off + lp.numel() * lp.element_size()].view(lp.dtype).view(lp.shape)

    def _push_round(self, plist):
        """Copy the parameters this rank has just updated (current stream) into its staging buffer on every peer's GPU: one side stream
        per peer, ordered after the current stream; psgd_peer_copy_async = cudaMemcpyAsync on THIS device's stream only (copy engines
        over NVLink).  Not torch's cross-device copy_: that records / waits events on the destination device's stream in this process,
        and a GPU with work from two processes' contexts time-slices between them (measured: 255 instead of 205 ms per step at 2 GPUs)."""
        lib = _lib.load_library()
        ev = torch.cuda.Event()
        ev.record()
        for k, st in self._peer_streams.items():
            st.wait_event(ev)
            base = self._stage[k].data_ptr()
            for p in plist:
                lp = self._local(p)
                rc = lib.psgd_peer_copy_async(base + self._stage_off[id(p)], lp.data_ptr(), lp.numel() * lp.element_size(), st.cuda_stream)
                if rc != 0:
                    raise _lib.EngineError(f"psgd_peer_copy_async failed ({rc})")

    def _unpack_p2p(self, my_rank):
        """After the end-of-step barrier: the peers' staging buffers on this GPU hold their updated parameters; copy them out (local)."""
        dst, src = [], []
        for p in self._p2p_plist:
            r = self._owner[id(p)]
            if r != my_rank:
                dst.append(self._local(p).detach())
                src.append(self._stage_view(self._inbox[r], p))
        with torch.no_grad():
            if dst:
                torch._foreach_copy_(dst, src)

    def _exchange_round(self, plists, maxb, my_rank, world):
        """One all-gather of this round's batches (plists[r] = the parameters rank r has just updated; maxb = the largest packed size).
        On CUDA everything here runs on a side stream ordered after the compute stream, so that packing, the collective and unpacking
        of round i overlap the engine calls of round i + 1; step() joins the side stream at its end."""
        dev = None
        for pl in plists:
            if pl:
                dev = self._local(pl[0]).device
                break
        if dev is None or maxb == 0:
            return
        need = (world + 1) * maxb
        if self._xbuf is None or self._xbuf.numel() < need or self._xbuf.device != dev:
            if self._comm_stream is not None:
                self._comm_stream.synchronize()
            self._xbuf = torch.empty(need, dtype=torch.uint8, device=dev)
        send, recv = self._xbuf[:maxb], self._xbuf[maxb:need]
        self._xbytes_step += world * maxb

        def views(buf, plist):
            out, off = [], 0
            for p in plist:
                lp = self._local(p)
                nb = lp.numel() * lp.element_size()
                out.append(buf[off:off + nb].view(lp.dtype).view(lp.shape))
                off += (nb + 15) // 16 * 16
            return out

        def run():
            with torch.no_grad():
                if plists[my_rank]:
                    torch._foreach_copy_(views(send, plists[my_rank]), [self._local(p).detach() for p in plists[my_rank]])
                torch.distributed.all_gather_into_tensor(recv, send)
                dst, src = [], []
                for r in range(world):
                    if r != my_rank and plists[r]:
                        dst += [self._local(p).detach() for p in plists[r]]
                        src += views(recv[r * maxb:(r + 1) * maxb], plists[r])
                if dst:
                    torch._foreach_copy_(dst, src)

        if dev.type == "cuda":
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=dev)
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream):
                run()
        else:
            run()

    def _make_batches(self, plist):
        """Consecutive-in-bucket grouping of parameters by (squeezed shape, dtype, dense/diagonal pattern), at most _lib.MAX_BATCH per
        batch and at most ~1.5 GB of gradients per batch; order-3+ tensors stay alone.  Depends on shapes only: identical on every rank."""
        buckets, order = {}, []
        for p in plist:
            lp = self._local(p)
            shp = tuple(lp.squeeze().shape)
            key = (shp, lp.dtype) if len(shp) <= 2 else ("single", id(p))
            if key not in buckets:
                buckets[key] = []
                order.append(key)
            buckets[key].append(p)
        out = []
        for key in order:
            ps = buckets[key]
            numel = max(1, self._local(ps[0]).numel())
            cap = 1 if key[0] == "single" else max(1, min(_lib.MAX_BATCH, int(self.batch_numel_cap // numel)))
            for i in range(0, len(ps), cap):
                out.append(ps[i:i + cap])
        return out


class KWNS4DTensor(KWNS4):
    """dtensor.py:4-185: every rank preconditions its LOCAL shard of each DTensor parameter independently."""

    def _needs_rng_sync(self):
        return True  # dtensor.py:89-96 syncs unconditionally

    def _local(self, t):
        return t.to_local() if hasattr(t, "to_local") else t

    def _resync(self, p, state, group, momentum):
        # dtensor.py:167-179: resync along replicated mesh dims only (sharded dims: "NOT implemented" in the reference)
        if state["step"] % group["resync_every"] != 0 or not hasattr(p, "placements"):
            return
        from torch.distributed.tensor.placement_types import Replicate
        for mesh_dim, placement in enumerate(p.placements):
            if isinstance(placement, Replicate):
                pg = p.device_mesh.get_group(mesh_dim)
                src = torch.distributed.get_process_group_ranks(pg)[0]
                torch.distributed.broadcast(p.to_local(), src=src, group=pg)
                if momentum > 0.0:
                    torch.distributed.broadcast(state["ema"], src=src, group=pg)
                for q, ell in zip(*state["QL"]):
                    torch.distributed.broadcast(q, src=src, group=pg)
                    torch.distributed.broadcast(ell, src=src, group=pg)
