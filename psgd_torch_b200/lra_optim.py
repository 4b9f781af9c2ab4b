"""LRAWhitenOptimizer -- a torch.optim.Optimizer for the PSGD low-rank-approximation whitening preconditioner.

The reference ships NO torch.optim wrapper for LRA: only the closure-style class psgd.LRAWhiten
(/root/reference/psgd.py:1075-1190).  This wrapper is therefore authored here: its constructor follows LRAWhiten.__init__
(psgd.py:1094-1128), its step() follows LRAWhiten.step (psgd.py:1131-1190) minus the closure (gradients are read from
p.grad like KWNS4 does, ddp.py:113), and its structure (param_groups, RNG discipline) follows KWNS4 (ddp.py:98-176).
One global preconditioner Q = (I + U V^T) diag(d) acts on the concatenation of all gradients (psgd.py:1142).
"""
import torch

from . import _lib
from . import psgd
from .kwns4 import KWNS4


class LRAWhitenOptimizer(torch.optim.Optimizer):
    def __init__(self, params, rank_of_approximation: int = 10, preconditioner_init_scale: float | None = None, lr_params=0.001,
                 lr_preconditioner=0.1, betaL=0.9, damping=1e-9, momentum=0.0, grad_clip_max_amps=(2.0, 10.0),
                 preconditioner_update_probability=1.0, update_preconditioner_first=True, whiten_grad=True,
                 preconditioner_dtype: torch.dtype | None = None):
        defaults = dict(lr_params=lr_params, lr_preconditioner=lr_preconditioner, betaL=betaL, damping=damping,
                        momentum=momentum if (0 < momentum < 1) else 0.0, grad_clip_max_amps=grad_clip_max_amps,
                        preconditioner_update_probability=preconditioner_update_probability,
                        update_preconditioner_first=update_preconditioner_first)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("LRAWhitenOptimizer fits ONE global preconditioner: pass a single parameter group")
        ps = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        self._params = ps
        dtype = preconditioner_dtype or ps[0].dtype
        device = ps[0].device
        self._sizes = [p.numel() for p in ps]
        n = sum(self._sizes)
        r = rank_of_approximation
        assert 0 < r < n and r <= 64, "the engine supports ranks 1..64 (rank 0 = diagonal preconditioner is not built)"
        if not whiten_grad:
            assert defaults["momentum"] > 0, "Cannot whiten momentum if the momentum setting is invalid."  # psgd.py:1126-1127
        self._whiten_grad = whiten_grad
        # psgd.py:1114-1123
        U = torch.randn(n, r, dtype=dtype, device=device)
        U *= 0.1 ** 0.5 / torch.linalg.vector_norm(U)
        V = torch.randn(n, r, dtype=dtype, device=device)
        V *= 0.1 ** 0.5 / torch.linalg.vector_norm(V)
        self._UVd = [U, V]
        if preconditioner_init_scale is not None:
            self._UVd.append(torch.ones(n, 1, dtype=dtype, device=device) * preconditioner_init_scale)
        self._Luvd = [torch.zeros([], dtype=torch.float32, device=device) for _ in range(3)]
        self._m, self._counter_m = None, 0
        self._dtype, self._device, self._n = dtype, device, n
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=device)
        # RNG discipline of KWNS4 (ddp.py:88-96): under torch.distributed the ranks draw randn_like(g) and the U/V coin from private
        # generator states synchronised at construction, so that the replicated preconditioners stay identical
        self._init_rng_sync()

    _needs_rng_sync = KWNS4._needs_rng_sync
    _init_rng_sync = KWNS4._init_rng_sync
    _rng_enter = KWNS4._rng_enter
    _rng_exit = KWNS4._rng_exit

    # ---- checkpoints: the global preconditioner, its Lipschitz constants, the momentum buffer and its counter live outside the
    # per-parameter state of torch.optim.Optimizer; they ride under the reserved state key "psgd_lra" (tensors + python scalars) ----
    def state_dict(self):
        sd = super().state_dict()
        sd["state"] = dict(sd["state"])
        blob = {"UVd": [t.detach().clone() for t in self._UVd], "Luvd": [t.detach().clone() for t in self._Luvd],
                "m": None if self._m is None else self._m.detach().clone(), "counter_m": int(self._counter_m)}
        if self.is_distributed:
            blob["rng_cpu"] = self.cpu_rng_state.clone()
            blob["rng_cuda"] = None if self.cuda_rng_state is None else self.cuda_rng_state.clone()
        sd["state"]["psgd_lra"] = blob
        return sd

    def load_state_dict(self, state_dict):
        state_dict = {**state_dict, "state": dict(state_dict["state"])}
        blob = state_dict["state"].pop("psgd_lra", None)
        super().load_state_dict(state_dict)
        if blob is None:
            return
        mv = lambda t, dt: t.detach().to(device=self._device, dtype=dt).contiguous().clone()
        UVd = [mv(t, self._dtype) for t in blob["UVd"]]
        if UVd[0].shape != (self._n, self._UVd[0].shape[1]):
            raise ValueError(f"checkpoint holds an LRA preconditioner of shape {tuple(UVd[0].shape)}, this optimizer needs "
                             f"{(self._n, self._UVd[0].shape[1])}")
        self._UVd = UVd
        self._Luvd = [mv(t, torch.float32) for t in blob["Luvd"]]
        self._m = None if blob["m"] is None else mv(blob["m"], self._dtype)
        self._counter_m = int(blob["counter_m"])
        if self.is_distributed and blob.get("rng_cpu") is not None:
            self.cpu_rng_state = blob["rng_cpu"].clone().cpu()
            if blob.get("rng_cuda") is not None and self.cuda_rng_state is not None:
                self.cuda_rng_state = blob["rng_cuda"].clone().cpu()

    @torch.no_grad()
    def step(self):
        ext = self._rng_enter()
        try:
            self._step()
        finally:
            self._rng_exit(ext)

    def _step(self):
        g = self.param_groups[0]
        lib = _lib.load_library()
        h = _lib.handle_for(self._device)
        grads = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1, 1) for p in self._params]
        grad = torch.cat(grads).to(self._dtype)  # psgd.py:1142
        if len(self._UVd) < 3:  # psgd.py:1144-1145 (one-off, host-side torch ops)
            self._UVd.append(((torch.mean(grad.float() ** 4) + g["damping"] ** 4) ** (-1 / 8)).to(self._dtype) * torch.ones_like(grad))
        momentum = g["momentum"]
        if momentum > 0:  # psgd.py:1147-1153
            beta = min(self._counter_m / (1 + self._counter_m), momentum)
            self._counter_m += 1
            if self._m is None:
                self._m = torch.zeros_like(grad)
            rc = lib.psgd_kwns4_head(h, grad.numel(), _lib.ptr(grad), _lib.dtype_code(grad), _lib.ptr(grad), _lib.dtype_code(grad),
                                     0.0, 0.0, 1, _lib.ptr(self._m), None, _lib.dtype_code(grad), float(beta),
                                     _lib.stream_ptr(self._device))
            _lib.check(h, rc, "psgd_kwns4_head")
        else:
            self._m, self._counter_m = None, 0
        if torch.rand([]) < g["preconditioner_update_probability"]:  # psgd.py:1157-1160
            first, last = g["update_preconditioner_first"], not g["update_preconditioner_first"]
        else:
            first, last = False, False
        whiten = grad if self._whiten_grad else self._m
        if first:
            psgd.update_precond_lra_whiten(self._UVd, self._Luvd, whiten, lr=g["lr_preconditioner"], betaL=g["betaL"], damping=g["damping"])
        pre = psgd.precond_grad_lra(self._UVd, self._m if momentum > 0 else grad, sumsq_out=self._sumsq)  # psgd.py:1168-1171
        if last:
            psgd.update_precond_lra_whiten(self._UVd, self._Luvd, whiten, lr=g["lr_preconditioner"], betaL=g["betaL"], damping=g["damping"])
        # psgd.py:1179-1187: one global clipping factor, then scatter into the parameters
        max_avg_amp, max_elem_amp = g["grad_clip_max_amps"]
        off = 0
        flat = pre.view(-1)
        for p, sz in zip(self._params, self._sizes):
            if not p.is_contiguous():
                raise _lib.EngineError("LRAWhitenOptimizer needs contiguous parameters")
            sl = flat[off:off + sz]
            rc = lib.psgd_kwns4_tail(h, sz, self._n, _lib.ptr(p), _lib.dtype_code(p), _lib.ptr(sl), _lib.dtype_code(sl),
                                     _lib.ptr(self._sumsq), float(max_avg_amp), float(max_elem_amp), float(g["lr_params"]),
                                     _lib.stream_ptr(self._device))
            _lib.check(h, rc, "psgd_kwns4_tail")
            off += sz
