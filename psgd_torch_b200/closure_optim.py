"""Closure-style PSGD optimizers on top of the engine's functional API: KronWhiten, KronNewton, LRAWhiten, LRANewton.

These are the callers of the hot path in the reference (SURVEY.md 8b "Callers", component inventory rows 7-8): host-side autograd glue
with no arithmetic of their own beyond momentum / clipping.  Same constructor arguments, mutable attributes (lr_params,
lr_preconditioner, betaL, damping, momentum, ...), `step(closure)` contract (returns whatever the closure returns) and RNG consumption
order as /root/reference/psgd.py:516-654 (KronWhiten), 832-978 (KronNewton), 1075-1190 (LRAWhiten), 1201-1330 (LRANewton); the
preconditioner math is the engine's (psgd_torch_b200.psgd -> libpsgd_b200.so).  Written around one shared skeleton instead of four
stand-alone classes.  Real CUDA (sm_100a) parameters only; complex parameters are not supported.
"""
import torch

from . import psgd

_WHITEN = {"QUAD4P": "quad4p", "PRO4P": "pro4p", "QEP": "qep", "EQ": "eq", "QEQ": "qeq", "QUAD": "quad", "Q0.5EQ1.5": "q0p5eq1p5",
           "Q0p5EQ1p5": "q0p5eq1p5"}


class _ClosureOptimizer:
    """Shared skeleton: parameter bookkeeping, closure evaluation (gradients, optional Hessian-vector products), momentum."""

    def __init__(self, params_with_grad, lr_params, lr_preconditioner, betaL, damping, momentum, preconditioner_update_probability):
        self.lr_params = lr_params
        self.lr_preconditioner = lr_preconditioner
        self.betaL = betaL
        self.damping = damping
        self.momentum = momentum if (0 < momentum < 1) else 0.0
        self.preconditioner_update_probability = preconditioner_update_probability
        params_with_grad = [params_with_grad] if isinstance(params_with_grad, torch.Tensor) else params_with_grad
        self._params_with_grad = [p for p in params_with_grad if p.requires_grad]
        if any(torch.is_complex(p) for p in self._params_with_grad):
            raise NotImplementedError("complex parameters are not supported by the engine")
        self._counter_m = 0

    # --- closure evaluation -------------------------------------------------------------------
    @staticmethod
    def _loss_of(returns):
        return returns if isinstance(returns, torch.Tensor) else returns[0]

    def _grads(self, closure):
        with torch.enable_grad():
            returns = closure()
            grads = torch.autograd.grad(self._loss_of(returns), self._params_with_grad)
        return returns, grads

    def _grads_and_hvps(self, closure, exact, delta):
        """(closure returns, grads, probes vs, Hessian-vector products): psgd.py:916-938 / 1266-1288."""
        ps = self._params_with_grad
        if exact:
            with torch.enable_grad():
                returns = closure()
                grads = torch.autograd.grad(self._loss_of(returns), ps, create_graph=True)
                vs = [torch.randn_like(p) for p in ps]
                Hvs = torch.autograd.grad(grads, ps, vs)
            return returns, [g.detach() for g in grads], vs, Hvs
        returns, grads = self._grads(closure)
        vs = [torch.randn_like(p) for p in ps]
        for p, v in zip(ps, vs):
            p.add_(v, alpha=delta)
        _, pgrads = self._grads(closure)
        Hvs = [(pg - g) / delta for pg, g in zip(pgrads, grads)]
        for p, v in zip(ps, vs):
            p.sub_(v, alpha=delta)
        return returns, grads, vs, Hvs

    def _ema_beta(self):
        beta = min(self._counter_m / (1 + self._counter_m), self.momentum)
        self._counter_m += 1
        return beta

    def _coin_update_order(self, first):
        """psgd.py:623-626: one CPU coin decides whether this step updates the preconditioner, before or after preconditioning."""
        if torch.rand([]) < self.preconditioner_update_probability:
            return first, not first
        return False, False


def _clip_amps_(g, max_avg_amp, max_element_amp):
    """psgd.py:645-651: scale down if the average amplitude exceeds max_avg_amp (host-side branch as in the reference), clamp elements."""
    avg_amp = torch.sqrt(torch.mean(g * g))
    if avg_amp > max_avg_amp:
        g *= max_avg_amp / avg_amp
    g.clamp_(min=-max_element_amp, max=max_element_amp)


# ------------------------------------------------------------------------------------------------
# Kron
# ------------------------------------------------------------------------------------------------
class _KronBase(_ClosureOptimizer):
    def _setup_kron(self, max_size, max_skew, init_scale, dQ, family):
        if dQ not in _WHITEN:
            raise AssertionError("Invalid choice for dQ")
        self._preconditioner_max_size, self._preconditioner_max_skew, self._dQ = max_size, max_skew, dQ
        self._update_precond = getattr(psgd, f"update_precond_kron_{family}_{_WHITEN[dQ]}")
        if dQ in ("QUAD4P", "PRO4P"):   # P is fitted directly: precondition with exprA(*Q, G)   psgd.py:573 / 908
            if max(torch.finfo(p.dtype).eps for p in self._params_with_grad) > 1e-6:
                print("Fitting P directly with half precision is risky.")
            self._precond_grad = lambda QL, exprs, G: exprs[0](*QL[0], G)
        else:
            self._precond_grad = psgd.precond_grad_kron
        if init_scale is None:
            self._QLs_exprs = None
            print("FYI: Will set the preconditioner initial scale on the fly. Recommend to set it manually.")
        else:
            self._QLs_exprs = [self._init(p.squeeze(), init_scale) for p in self._params_with_grad]
        self._ms = None

    def _init(self, t, scale):
        return psgd.init_kron(t, float(scale), self._preconditioner_max_size, self._preconditioner_max_skew, self._dQ)

    def _momentum_or(self, grads):
        if self.momentum > 0:
            beta = self._ema_beta()
            if self._ms is None:
                self._ms = [torch.zeros_like(g) for g in grads]
            for m, g in zip(self._ms, grads):
                m.mul_(beta).add_(g, alpha=1 - beta)
            return self._ms
        self._ms, self._counter_m = None, 0
        return grads


class KronWhiten(_KronBase):
    """PSGD with the Kronecker-product gradient/momentum whitening preconditioner (psgd.py:516-654)."""

    def __init__(self, params_with_grad, preconditioner_max_size=float("inf"), preconditioner_max_skew=1.0,
                 preconditioner_init_scale: float | None = None, lr_params=0.001, lr_preconditioner=0.1, betaL=0.9, damping=1e-9,
                 momentum=0.0, grad_clip_max_amps=(2.0, 10.0), preconditioner_update_probability=1.0, update_preconditioner_first=True,
                 whiten_grad=True, dQ="Q0.5EQ1.5"):
        super().__init__(params_with_grad, lr_params, lr_preconditioner, betaL, damping, momentum, preconditioner_update_probability)
        self.grad_clip_max_amps = grad_clip_max_amps
        self.update_preconditioner_first = update_preconditioner_first
        self._whiten_grad = whiten_grad
        if not whiten_grad:
            assert self.momentum > 0, "Cannot whiten momentum if the momentum setting is invalid."
        self._setup_kron(preconditioner_max_size, preconditioner_max_skew, preconditioner_init_scale, dQ, "whiten")

    @torch.no_grad()
    def step(self, closure):
        returns, grads = self._grads(closure)
        grads = [g.squeeze().contiguous() for g in grads]
        if self._QLs_exprs is None:  # psgd.py:606-609
            scale = max(torch.mean(torch.abs(g) ** 4) for g in grads)
            scale = (scale + self.damping ** 4) ** (-1 / 8)
            self._QLs_exprs = [self._init(g, scale) for g in grads]
        ms = self._momentum_or(grads)
        first, last = self._coin_update_order(self.update_preconditioner_first)
        targets = grads if self._whiten_grad else ms
        if first:
            for (QL, exprs), t in zip(self._QLs_exprs, targets):
                self._update_precond(QL, exprs, t, lr=self.lr_preconditioner, betaL=self.betaL, damping=self.damping)
        pre_grads = [self._precond_grad(QL, exprs, x) for (QL, exprs), x in zip(self._QLs_exprs, ms)]
        if last:
            for (QL, exprs), t in zip(self._QLs_exprs, targets):
                self._update_precond(QL, exprs, t, lr=self.lr_preconditioner, betaL=self.betaL, damping=self.damping)
        for p, g in zip(self._params_with_grad, pre_grads):
            _clip_amps_(g, *self.grad_clip_max_amps)
            p.subtract_(g.view_as(p), alpha=self.lr_params)
        return returns


class KronNewton(_KronBase):
    """PSGD with the Kronecker-product Newton-type preconditioner fitted on (vector, Hessian-vector-product) pairs (psgd.py:832-978)."""

    def __init__(self, params_with_grad, preconditioner_max_size=float("inf"), preconditioner_max_skew=1.0,
                 preconditioner_init_scale: float | None = None, lr_params=0.01, lr_preconditioner=0.1, betaL=0.9, damping=1e-9,
                 momentum=0.0, grad_clip_max_norm=float("inf"), preconditioner_update_probability=1.0, exact_hessian_vector_product=True,
                 dQ="Q0.5EQ1.5"):
        super().__init__(params_with_grad, lr_params, lr_preconditioner, betaL, damping, momentum, preconditioner_update_probability)
        self.grad_clip_max_norm = grad_clip_max_norm
        self._exact_hessian_vector_product = exact_hessian_vector_product
        self._delta_param_scale = max(torch.finfo(p.dtype).eps for p in self._params_with_grad) ** 0.5
        self._setup_kron(preconditioner_max_size, preconditioner_max_skew, preconditioner_init_scale, dQ, "newton")

    @torch.no_grad()
    def step(self, closure):
        if (torch.rand([]) < self.preconditioner_update_probability) or (self._QLs_exprs is None):
            returns, grads, vs, Hvs = self._grads_and_hvps(closure, self._exact_hessian_vector_product, self._delta_param_scale)
            if self._QLs_exprs is None:  # psgd.py:940-943
                scale = (sum(torch.sum(v * v) for v in vs) / sum(v.numel() for v in vs)) ** (1 / 4)
                scale = scale * (max(torch.mean(torch.abs(h) ** 4) for h in Hvs) + self.damping ** 4) ** (-1 / 8)
                self._QLs_exprs = [self._init(h.squeeze(), scale) for h in Hvs]
            for (QL, exprs), v, h in zip(self._QLs_exprs, vs, Hvs):
                self._update_precond(QL, exprs, v.squeeze().contiguous(), h.squeeze().contiguous(), lr=self.lr_preconditioner,
                                     betaL=self.betaL, damping=self.damping)
        else:
            returns, grads = self._grads(closure)
        grads = [g.squeeze().contiguous() for g in grads]
        xs = self._momentum_or(grads)
        pre_grads = [self._precond_grad(QL, exprs, x) for (QL, exprs), x in zip(self._QLs_exprs, xs)]
        lr = self.lr_params
        if self.grad_clip_max_norm < float("inf"):  # psgd.py:965-968
            grad_norm = torch.sqrt(sum(torch.sum(g * g) for g in pre_grads))
            if grad_norm > self.grad_clip_max_norm:
                lr = lr * self.grad_clip_max_norm / grad_norm
        for p, g in zip(self._params_with_grad, pre_grads):
            p.subtract_(lr * g.view_as(p))
        return returns


# ------------------------------------------------------------------------------------------------
# LRA: one global preconditioner on the concatenation of all gradients (psgd.py:1142)
# ------------------------------------------------------------------------------------------------
class _LRABase(_ClosureOptimizer):
    def _setup_lra(self, rank, init_scale):
        ps = self._params_with_grad
        dtype, device = ps[0].dtype, ps[0].device
        self._param_sizes = [p.numel() for p in ps]
        n = sum(self._param_sizes)
        assert 0 < rank < n and rank <= 64, "the engine supports ranks 1..64 (rank 0, a purely diagonal Q, is not built)"
        U = torch.randn(n, rank, dtype=dtype, device=device)
        U *= 0.1 ** 0.5 / torch.linalg.vector_norm(U)
        V = torch.randn(n, rank, dtype=dtype, device=device)
        V *= 0.1 ** 0.5 / torch.linalg.vector_norm(V)
        self._UVd = [U, V]
        if init_scale is None:
            print("FYI: Will set the preconditioner initial scale on the fly. Recommend to set it manually.")
        else:
            self._UVd.append(torch.ones(n, 1, dtype=dtype, device=device) * init_scale)
        self._Luvd = [torch.zeros([], dtype=torch.float32, device=device) for _ in range(3)]
        self._m = None

    @staticmethod
    def _cat(ts):
        return torch.cat([t.reshape(-1, 1) for t in ts])

    def _momentum_or(self, grad):
        if self.momentum > 0:
            beta = self._ema_beta()
            if self._m is None:
                self._m = torch.zeros_like(grad)
            self._m.mul_(beta).add_(grad, alpha=1 - beta)
            return self._m
        self._m, self._counter_m = None, 0
        return grad

    def _scatter(self, pre_grad, lr):
        off = 0
        for p, sz in zip(self._params_with_grad, self._param_sizes):
            p.subtract_(lr * pre_grad[off:off + sz].view_as(p))
            off += sz


class LRAWhiten(_LRABase):
    """PSGD with the low-rank-approximation gradient/momentum whitening preconditioner (psgd.py:1075-1190)."""

    def __init__(self, params_with_grad, rank_of_approximation: int = 10, preconditioner_init_scale: float | None = None, lr_params=0.001,
                 lr_preconditioner=0.1, betaL=0.9, damping=1e-9, momentum=0.0, grad_clip_max_amps=(2.0, 10.0),
                 preconditioner_update_probability=1.0, update_preconditioner_first=True, whiten_grad=True):
        super().__init__(params_with_grad, lr_params, lr_preconditioner, betaL, damping, momentum, preconditioner_update_probability)
        self.grad_clip_max_amps = grad_clip_max_amps
        self.update_preconditioner_first = update_preconditioner_first
        self._whiten_grad = whiten_grad
        if not whiten_grad:
            assert self.momentum > 0, "Cannot whiten momentum if the momentum setting is invalid."
        self._setup_lra(rank_of_approximation, preconditioner_init_scale)

    @torch.no_grad()
    def step(self, closure):
        returns, grads = self._grads(closure)
        grad = self._cat(grads)
        if len(self._UVd) < 3:  # psgd.py:1144-1145
            self._UVd.append((torch.mean(grad ** 4) + self.damping ** 4) ** (-1 / 8) * torch.ones_like(grad))
        x = self._momentum_or(grad)
        first, last = self._coin_update_order(self.update_preconditioner_first)
        target = grad if self._whiten_grad else self._m
        if first:
            psgd.update_precond_lra_whiten(self._UVd, self._Luvd, target, lr=self.lr_preconditioner, betaL=self.betaL, damping=self.damping)
        pre_grad = psgd.precond_grad_lra(self._UVd, x)
        if last:
            psgd.update_precond_lra_whiten(self._UVd, self._Luvd, target, lr=self.lr_preconditioner, betaL=self.betaL, damping=self.damping)
        _clip_amps_(pre_grad, *self.grad_clip_max_amps)
        self._scatter(pre_grad, self.lr_params)
        return returns


class LRANewton(_LRABase):
    """PSGD with the low-rank-approximation Newton-type preconditioner (psgd.py:1201-1330)."""

    def __init__(self, params_with_grad, rank_of_approximation: int = 10, preconditioner_init_scale: float | None = None, lr_params=0.01,
                 lr_preconditioner=0.1, betaL=0.9, damping=1e-9, momentum=0.0, grad_clip_max_norm=float("inf"),
                 preconditioner_update_probability=1.0, exact_hessian_vector_product=True):
        super().__init__(params_with_grad, lr_params, lr_preconditioner, betaL, damping, momentum, preconditioner_update_probability)
        self.grad_clip_max_norm = grad_clip_max_norm
        self._exact_hessian_vector_product = exact_hessian_vector_product
        self._delta_param_scale = torch.finfo(self._params_with_grad[0].dtype).eps ** 0.5
        self._setup_lra(rank_of_approximation, preconditioner_init_scale)

    @torch.no_grad()
    def step(self, closure):
        if (torch.rand([]) < self.preconditioner_update_probability) or (len(self._UVd) < 3):
            returns, grads, vs, Hvs = self._grads_and_hvps(closure, self._exact_hessian_vector_product, self._delta_param_scale)
            v, h = self._cat(vs), self._cat(Hvs)
            if len(self._UVd) < 3:  # psgd.py:1292-1293
                self._UVd.append(torch.mean(v * v) ** (1 / 4) * (torch.mean(h ** 4) + self.damping ** 4) ** (-1 / 8) * torch.ones_like(v))
            psgd.update_precond_lra_newton(self._UVd, self._Luvd, v, h, lr=self.lr_preconditioner, betaL=self.betaL, damping=self.damping)
        else:
            returns, grads = self._grads(closure)
        pre_grad = psgd.precond_grad_lra(self._UVd, self._momentum_or(self._cat(grads)))
        lr = self.lr_params
        if self.grad_clip_max_norm < float("inf"):  # psgd.py:1320-1323
            grad_norm = torch.linalg.vector_norm(pre_grad)
            if grad_norm > self.grad_clip_max_norm:
                lr = lr * self.grad_clip_max_norm / grad_norm
        self._scatter(pre_grad, lr)
        return returns
