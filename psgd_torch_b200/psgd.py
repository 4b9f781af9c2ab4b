"""Functional API of the reference's psgd.py, served by the B200 engine (libpsgd_b200.so).

Same names, argument meaning, defaults, in-place state mutation and RNG consumption order as
/root/reference/psgd.py for the hot path of BASELINE.json:north_star:

    init_kron                               psgd.py:161-263   (host-side; state layout identical)
    update_precond_kron_whiten_q0p5eq1p5    psgd.py:394-419   -> psgd_kron_whiten_q0p5eq1p5_update
    precond_grad_kron                       psgd.py:322-327   -> psgd_kron_precond_grad
    norm_lower_bound_spd / _skh             psgd.py:46-93     -> psgd_norm_lower_bound_*
    procrustes_step2                        psgd.py:101-124   -> psgd_procrustes_step2
    balance_kron_precond                    psgd.py:266-275   -> psgd_kron_balance
    update_precond_lra / _lra_whiten        psgd.py:994-1072  -> psgd_lra_update / psgd_lra_whiten_update
    update_precond_kron_whiten_{eq,qep,qeq,pro4p,quad,quad4p}            psgd.py:330-513  -> psgd_kron_update
    update_precond_kron_eq, update_precond_kron_newton_{eq,...,quad4p}   psgd.py:278-319, 657-829 -> psgd_kron_update
    procrustes_step3                        psgd.py:127-155   -> psgd_procrustes_step3
    precond_grad_lra                        psgd.py:1055-1063 -> psgd_lra_precond_grad

All tensors must live on a B200 (sm_100a) device; there is no CPU / PyTorch fallback.  Random numbers are drawn
here with torch, on the same generators and in the same order as the reference draws them (SURVEY.md 8b "RNG
contract"), and handed to the kernels as inputs -- so a DDP job whose ranks synchronise their generators (KWNS4
does, ddp.py:88-96) stays replica-consistent, and a parity test can pass `noise=` explicitly.
"""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import EngineError, KronNoiseT, KronT, LraT

_K_PROBES = 32  # psgd.py:46 (k=32; the code always calls with the default)


def lift2single(x):
    """psgd.py:96-98"""
    return x.to(torch.float32) if torch.finfo(x.dtype).eps > 1e-6 else x


# ------------------------------------------------------------------------------------------------
# contraction "expressions" (opaque to callers; the reference passes them back verbatim)
# ------------------------------------------------------------------------------------------------
class _ExprP:
    """exprP(*Q_conj, *Q, G) -> (kron_i Q_i^H Q_i) G   (psgd.py:251-252). Picklable, stateless."""

    def __init__(self, order):
        self.order = order

    def __call__(self, *ops):
        n = max(self.order, 1)
        assert len(ops) == 2 * n + 1, "exprP expects (*Q_conj, *Q, G)"
        Q, G = list(ops[n:2 * n]), ops[-1]
        return _apply_kron(Q, G)


class _ExprG:
    """exprGs[i](X, Y_conj): contraction of two tensors keeping only dim i (psgd.py:221-223, 240-243)."""

    def __init__(self, i, dense, order):
        self.i, self.dense, self.order = i, dense, order

    def __call__(self, X, Y):
        if self.order == 0:
            return X * Y
        if self.order == 1:
            return gemm(X.reshape(-1, 1), Y.reshape(-1, 1), trans_b=True) if self.dense else X * Y
        if self.order > 2:
            Mx = X.movedim(self.i, 0).reshape(X.shape[self.i], -1)
            My = Y.movedim(self.i, 0).reshape(Y.shape[self.i], -1)
            return gemm(Mx, My, trans_b=True) if self.dense else torch.sum(Mx * My, dim=1)
        if self.dense:
            return gemm(X, Y, trans_b=True) if self.i == 0 else gemm(X, Y, trans_a=True)
        return torch.sum(X * Y, dim=1 - self.i)


class _ExprA:
    """exprA(*Q, G) -> every factor applied once, Q_L G Q_R^T for a matrix (psgd.py:248-249). What KronWhiten / KronNewton use as the
    preconditioning step when P is fitted directly (psgd.py:573, 908)."""

    def __init__(self, order):
        self.order = order

    def __call__(self, *ops):
        n = max(self.order, 1)
        assert len(ops) == n + 1, "exprA expects (*Q, G)"
        return _apply_factors(list(ops[:n]), ops[-1])


class _ExprQ:
    """exprQs[i](q, T): the i-th factor applied along dim i of T (psgd.py:225-226, 245-246)."""

    def __init__(self, i, dense, order):
        self.i, self.dense, self.order = i, dense, order

    def __call__(self, q, T):
        if not self.dense or self.order == 0:
            shp = [1] * T.dim()
            if T.dim() > 0:
                shp[self.i] = -1
            return T * q.reshape(shp)
        if self.order == 1:
            return gemm(q, T.reshape(-1, 1)).reshape(T.shape)
        if self.order == 2:
            return gemm(q, T) if self.i == 0 else gemm(T, q, trans_b=True)
        return _mode_product(q, T, self.i, False)


def init_kron(t, Scale=1.0, max_size=float("inf"), max_skew=1.0, dQ="Q0.5EQ1.5"):
    """psgd.py:161-263. Returns [[Q, L], exprs] with the reference's state layout: Q[i] = scale*eye(s) (dense) or
    scale*ones(s) (diagonal, psgd.py:208) in t's dtype, L[i] = fp32 0-dim zeros, exprs as psgd.py:255-263 lists them per geometry."""
    if dQ not in _lib.DQ_CODES:
        raise AssertionError("Invalid choice for dQ")  # psgd.py:262
    if dQ in ("QUAD4P", "PRO4P"):  # psgd.py:186-187: the two geometries that fit P directly
        Scale = Scale ** 2
    shape = t.shape
    if len(shape) == 0:  # psgd.py:189-195
        Q = [Scale * torch.ones_like(t)]
        L = [lift2single(torch.zeros_like(t))]
        return [[Q, L], _exprs_for(dQ, 0, (_ExprG(0, False, 0),), (_ExprQ(0, False, 0),))]
    if len(shape) > 26:
        raise ValueError(f"Got tensor with dim {len(t.shape)}; einsum runs out of letters; replace 26 with larger numbers.")
    scale = Scale ** (1 / len(shape))
    Q, L, exprGs, exprQs = [], [], [], []
    for i, size in enumerate(shape):
        L.append(lift2single(torch.zeros([], dtype=t.dtype, device=t.device)))
        if size <= 1 or size > max_size or size ** 2 > max_skew * t.numel():
            Q.append(scale * torch.ones(size, dtype=t.dtype, device=t.device))
            exprGs.append(_ExprG(i, False, len(shape)))
            exprQs.append(_ExprQ(i, False, len(shape)))
        else:
            Q.append(scale * torch.eye(size, dtype=t.dtype, device=t.device))
            exprGs.append(_ExprG(i, True, len(shape)))
            exprQs.append(_ExprQ(i, True, len(shape)))
    return [[Q, L], _exprs_for(dQ, len(shape), tuple(exprGs), tuple(exprQs))]


def _exprs_for(dQ, order, exprGs, exprQs):
    """psgd.py:255-263: which expressions each geometry carries."""
    if dQ == "QEP":
        return (_ExprP(order), exprGs, exprQs)
    if dQ == "EQ":
        return (_ExprP(order), exprGs, _ExprA(order))
    if dQ in ("QUAD4P", "PRO4P"):
        return (_ExprA(order), exprGs)
    return (_ExprP(order), exprGs)


def exprs_for_state(Q, dQ="Q0.5EQ1.5"):
    """Rebuild the `exprs` tuple of init_kron from the factors of a loaded checkpoint (dense factor = 2-D, diagonal = 1-D, scalar tensor =
    one 0-dim factor): the expressions depend on the order and on which factors are dense, nothing else."""
    if len(Q) == 1 and Q[0].dim() == 0:
        return _exprs_for(dQ, 0, (_ExprG(0, False, 0),), (_ExprQ(0, False, 0),))
    order = len(Q)
    return _exprs_for(dQ, order, tuple(_ExprG(i, q.dim() == 2, order) for i, q in enumerate(Q)),
                      tuple(_ExprQ(i, q.dim() == 2, order) for i, q in enumerate(Q)))


# ------------------------------------------------------------------------------------------------
# engine plumbing
# ------------------------------------------------------------------------------------------------
def _kron_desc(Q, L, G):
    order = G.dim()
    assert order <= 2, "order >= 3 tensors take the host-side composition (_update_kron_nd / _apply_kron_nd)"
    if len(Q) != max(order, 1):
        raise EngineError("Q does not match the tensor order")
    for q in Q:
        if q.dtype != G.dtype or q.device != G.device or not q.is_contiguous():
            raise EngineError("Q factors must be contiguous and share G's dtype and device")
    k = KronT()
    if order <= 1:
        k.m, k.n, k.has_r = max(G.numel(), 1), 1, 0
        k.kind_l, k.kind_r = (_lib.PSGD_DENSE if Q[0].dim() == 2 else _lib.PSGD_DIAG), _lib.PSGD_DIAG
        k.QL, k.QR, k.LL, k.LR = Q[0].data_ptr(), None, L[0].data_ptr() if L is not None else None, None
    else:
        k.m, k.n, k.has_r = G.shape[0], G.shape[1], 1
        k.kind_l = _lib.PSGD_DENSE if Q[0].dim() == 2 else _lib.PSGD_DIAG
        k.kind_r = _lib.PSGD_DENSE if Q[1].dim() == 2 else _lib.PSGD_DIAG
        k.QL, k.QR = Q[0].data_ptr(), Q[1].data_ptr()
        k.LL, k.LR = (L[0].data_ptr(), L[1].data_ptr()) if L is not None else (None, None)
    k.dtype = _lib.dtype_code(G)
    return k


_dummy_L = {}


def _dummy(device):
    d = _dummy_L.get(device)
    if d is None:
        d = torch.zeros(2, dtype=torch.float32, device=device)
        _dummy_L[device] = d
    return d


def _mode_product(q, X, i, transpose):
    """Apply a dense factor along dim i of an order >= 3 tensor: a GEMM on the mode-i matricisation (engine), permutes by torch."""
    Xi = X.movedim(i, 0)
    shp = Xi.shape
    Y = gemm(q, Xi.reshape(shp[0], -1), trans_a=transpose)
    return Y.reshape(shp).movedim(0, i)


def _apply_kron_nd(Q, G):
    """psgd.py:322-327 for tensors of order >= 3: host-side composition of engine GEMMs, one mode at a time."""
    X = G
    for i, q in enumerate(Q):
        if q.dim() == 2:
            X = _mode_product(q, X, i, False)
    for i, q in enumerate(Q):
        if q.dim() == 2:
            X = _mode_product(q, X, i, True)
        else:
            shp = [1] * X.dim()
            shp[i] = -1
            X = X * (q * q).reshape(shp)
    return X.contiguous()


def _update_kron_nd(Q, L, G, lr, betaL, damping, noise):
    """psgd.py:394-419 for order >= 3: Pg by mode products, then per factor the engine's factor update (bound, L, Newton-Schulz,
    procrustes) on the mode-i Gram."""
    lib = _lib.load_library()
    h = _lib.handle_for(G.device)
    dt = _lib.dtype_code(G)
    damp = damping + torch.finfo(G.dtype).eps * G.abs()
    Pg = _apply_kron_nd(Q, G + damp * noise["N"])
    numel = G.numel()
    for i, q in enumerate(Q):
        M = Pg.movedim(i, 0).reshape(Pg.shape[i], -1).contiguous()
        s = q.shape[0]
        if q.dim() == 2:
            term1 = gemm(M, M, trans_b=True)
            ws = _lib.workspace(G.device, lib.psgd_helper_workspace_bytes(h, s, dt))
            rc = lib.psgd_kron_factor_update(h, dt, _lib.PSGD_DENSE, s, _lib.ptr(q), _lib.ptr(L[i]), _lib.ptr(term1), float(numel / s),
                                             float(lr), float(betaL), _lib.ptr(noise["spd"][i]), _lib.ptr(noise["skh"][i]), _lib.ptr(ws),
                                             ws.numel(), _lib.stream_ptr(G.device))
        else:
            term1 = torch.sum(M.float() * M.float(), dim=1).contiguous()
            rc = lib.psgd_kron_factor_update(h, dt, _lib.PSGD_DIAG, s, _lib.ptr(q), _lib.ptr(L[i]), _lib.ptr(term1), float(numel / s),
                                             float(lr), float(betaL), None, None, None, 0, _lib.stream_ptr(G.device))
        _lib.check(h, rc, "psgd_kron_factor_update")
    if noise.get("balance", False):
        balance_kron_precond(Q)


def _apply_kron(Q, G, sumsq_out=None):
    if not G.is_cuda:
        raise EngineError("psgd_torch_b200 runs on CUDA (sm_100a) tensors only")
    G = G.contiguous()
    if G.dim() > 2:
        out = _apply_kron_nd(Q, G)
        if sumsq_out is not None:
            sumsq_out.copy_(torch.sum(out.float() ** 2).reshape(sumsq_out.shape))
        return out
    k = _kron_desc(Q, None, G)
    d = _dummy(G.device)  # the apply never touches L; the descriptor validator wants non-null pointers
    k.LL, k.LR = d.data_ptr(), d.data_ptr() + 4
    h = _lib.handle_for(G.device)
    lib = _lib.load_library()
    nbytes = lib.psgd_kron_workspace_bytes(h, C.byref(k))
    ws = _lib.workspace(G.device, nbytes)
    out = torch.empty_like(G)
    rc = lib.psgd_kron_precond_grad(h, C.byref(k), _lib.ptr(G), _lib.ptr(out), _lib.ptr(sumsq_out), _lib.ptr(ws), ws.numel(),
                                    _lib.stream_ptr(G.device))
    _lib.check(h, rc, "psgd_kron_precond_grad")
    return out


def precond_grad_kron(QL, exprs, G, sumsq_out=None):
    """psgd.py:322-327: returns a new tensor (kron_i Q_i^T Q_i) G. `sumsq_out` (optional fp32 1-element CUDA tensor)
    receives sum(out^2), fused into the last product -- KWNS4's clipping rule (ddp.py:153) reads it."""
    return _apply_kron(QL[0], G, sumsq_out)


def draw_kron_noise(G, Q):
    """Draw the random inputs of one update in the reference's order: randn_like(G) (psgd.py:403), then per
    dense factor randn(32,s) for norm_lower_bound_spd (psgd.py:62) and randn(32,s) for norm_lower_bound_skh inside
    procrustes_step2 (psgd.py:87), finally the CPU coin torch.rand([]) < 0.01 of psgd.py:418."""
    noise = {"N": torch.randn_like(G), "spd": [], "skh": []}
    for q in Q:
        if q.dim() == 2:
            noise["spd"].append(torch.randn(_K_PROBES, q.shape[1], dtype=q.dtype, device=q.device))
            noise["skh"].append(torch.randn(_K_PROBES, q.shape[1], dtype=q.dtype, device=q.device))
        else:
            noise["spd"].append(None)
            noise["skh"].append(None)
    noise["balance"] = bool(torch.rand([]) < 0.01)
    return noise


def update_precond_kron_whiten_q0p5eq1p5(QL, exprs, G, lr=0.1, betaL=0.9, damping=1e-9, noise=None):
    """psgd.py:394-419: update the Kron preconditioner Q as dQ = Q^0.5 E Q^1.5, in place on Q and L.
    `noise` (optional) = dict from draw_kron_noise; by default it is drawn here like the reference draws it."""
    Q, L = QL
    if not G.is_cuda:
        raise EngineError("psgd_torch_b200 runs on CUDA (sm_100a) tensors only")
    G = G.contiguous()
    if noise is None:
        noise = draw_kron_noise_philox(G, Q) if (_NOISE_MODE == "philox" and G.dim() <= 2) else draw_kron_noise(G, Q)
    if G.dim() > 2:
        return _update_kron_nd(Q, L, G, lr, betaL, damping, noise)
    k = _kron_desc(Q, L, G)
    nz = KronNoiseT()
    _fill_noise(nz, noise, len(Q))
    h = _lib.handle_for(G.device)
    lib = _lib.load_library()
    nbytes = lib.psgd_kron_workspace_bytes(h, C.byref(k))
    ws = _lib.workspace(G.device, nbytes)
    rc = lib.psgd_kron_whiten_q0p5eq1p5_update(h, C.byref(k), _lib.ptr(G), float(lr), float(betaL), float(damping), C.byref(nz),
                                               int(bool(noise.get("balance", False))), _lib.ptr(ws), ws.numel(),
                                               _lib.stream_ptr(G.device))
    _lib.check(h, rc, "psgd_kron_whiten_q0p5eq1p5_update")


_NOISE_MODE = "torch"
_philox_calls = 0


def set_noise_mode(mode):
    """"torch" (default): every random number of an update is drawn by torch on the host side, in the reference's order (parity with the
    reference's RNG streams; SURVEY.md 8b "RNG contract").  "philox": performance mode -- the damping noise and the norm-bound probes are
    drawn inside the engine's kernels (counter-based Philox4x32-10); the host only draws a 62-bit seed per update and the balancing coin
    from torch's CPU generator, so DDP replicas whose CPU generators are synchronised (KWNS4 does that) still see identical noise."""
    global _NOISE_MODE
    if mode not in ("torch", "philox"):
        raise ValueError("noise mode must be 'torch' or 'philox'")
    _NOISE_MODE = mode


def draw_kron_noise_philox(G, Q):
    """Performance-mode counterpart of draw_kron_noise: no tensors, a seed for the in-kernel generator + the CPU coin of psgd.py:418."""
    global _philox_calls
    _philox_calls += 1
    return {"N": None, "spd": [None] * len(Q), "skh": [None] * len(Q), "seed": int(torch.randint(0, 1 << 62, (1,))),
            "offset": _philox_calls, "balance": bool(torch.rand([]) < 0.01)}


def _fill_noise(nz, noise, nfactors):
    nz.N = noise["N"].data_ptr() if noise["N"] is not None else None
    nz.philox_seed, nz.philox_offset = int(noise.get("seed", 0)), int(noise.get("offset", 0))
    spd, skh = noise["spd"], noise["skh"]
    nz.V0_spd_l = spd[0].data_ptr() if spd[0] is not None else None
    nz.V0_skh_l = skh[0].data_ptr() if skh[0] is not None else None
    if nfactors > 1:
        nz.V0_spd_r = spd[1].data_ptr() if spd[1] is not None else None
        nz.V0_skh_r = skh[1].data_ptr() if skh[1] is not None else None


def _batch_descs(QLs, Gs, with_L):
    n = len(QLs)
    if not 1 <= n <= _lib.MAX_BATCH:
        raise EngineError(f"a batched call takes 1..{_lib.MAX_BATCH} units, got {n}")
    ks = (KronT * n)()
    for u, (QL, G) in enumerate(zip(QLs, Gs)):
        if G.dim() > 2 or not G.is_cuda or not G.is_contiguous():
            raise EngineError("batched calls take contiguous CUDA tensors of order <= 2")
        k = _kron_desc(QL[0], QL[1] if with_L else None, G)
        if not with_L:
            d = _dummy(G.device)
            k.LL, k.LR = d.data_ptr(), d.data_ptr() + 4
        ks[u] = k
    return ks


def update_precond_kron_whiten_q0p5eq1p5_batched(QLs, exprs, Gs, lr=0.1, betaL=0.9, damping=1e-9, noises=None):
    """psgd.py:394-419 for a list of same-shape tensors in ONE engine call (psgd_kron_whiten_q0p5eq1p5_update_batched): what the reference's
    per-parameter loop (ddp.py:112-161) does one tensor at a time.  QLs = [QL_0, QL_1, ...] (each as init_kron returns it), Gs = list of
    gradients of one shape / dtype; `exprs` is accepted for symmetry with the single-tensor call and not used.  Random draws: per unit,
    in list order, exactly the single-tensor call's draws (`noises` = list of dicts from draw_kron_noise replays them)."""
    if noises is None:
        draw = draw_kron_noise_philox if _NOISE_MODE == "philox" else draw_kron_noise
        noises = [draw(G, QL[0]) for QL, G in zip(QLs, Gs)]
    n = len(QLs)
    ks = _batch_descs(QLs, Gs, True)
    nzs = (KronNoiseT * n)()
    for u in range(n):
        _fill_noise(nzs[u], noises[u], len(QLs[u][0]))
    gp = (C.c_void_p * n)(*[G.data_ptr() for G in Gs])
    bal = (C.c_int * n)(*[int(bool(nz.get("balance", False))) for nz in noises])
    dev = Gs[0].device
    h = _lib.handle_for(dev)
    lib = _lib.load_library()
    ws = _lib.workspace(dev, lib.psgd_kron_batch_workspace_bytes(h, ks, n))
    rc = lib.psgd_kron_whiten_q0p5eq1p5_update_batched(h, ks, n, gp, float(lr), float(betaL), float(damping), nzs, bal, _lib.ptr(ws),
                                                       ws.numel(), _lib.stream_ptr(dev))
    _lib.check(h, rc, "psgd_kron_whiten_q0p5eq1p5_update_batched")


def precond_grad_kron_batched(QLs, exprs, Gs, sumsq_out=None):
    """psgd.py:322-327 for a list of same-shape tensors in one engine call; returns the list of preconditioned gradients.  `sumsq_out`
    (optional fp32 CUDA tensor with len(Gs) elements) receives sum(out_u^2) per unit."""
    n = len(QLs)
    ks = _batch_descs(QLs, Gs, False)
    outs = [torch.empty_like(G) for G in Gs]
    xp = (C.c_void_p * n)(*[G.data_ptr() for G in Gs])
    hp = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
    dev = Gs[0].device
    h = _lib.handle_for(dev)
    lib = _lib.load_library()
    ws = _lib.workspace(dev, lib.psgd_kron_batch_workspace_bytes(h, ks, n))
    rc = lib.psgd_kron_precond_grad_batched(h, ks, n, xp, hp, _lib.ptr(sumsq_out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(h, rc, "psgd_kron_precond_grad_batched")
    return outs


def balance_kron_precond(Q):
    """psgd.py:266-275, in place."""
    if len(Q) <= 1:
        return
    if len(Q) > 2:  # order >= 3: host-side composition (fp32 factors, cf. k_balance_scale)
        norms = torch.stack([q.abs().max().float() for q in Q])
        gmean = torch.prod(norms) ** (1 / len(Q))
        for q, nrm in zip(Q, norms):
            q.mul_(gmean / nrm)
        return
    G = torch.empty(Q[0].shape[0], Q[1].shape[0], dtype=Q[0].dtype, device="meta")
    k = KronT()
    k.m, k.n, k.has_r = Q[0].shape[0], Q[1].shape[0], 1
    k.kind_l = _lib.PSGD_DENSE if Q[0].dim() == 2 else _lib.PSGD_DIAG
    k.kind_r = _lib.PSGD_DENSE if Q[1].dim() == 2 else _lib.PSGD_DIAG
    k.dtype = _lib.dtype_code(Q[0])
    d = _dummy(Q[0].device)
    k.QL, k.QR, k.LL, k.LR = Q[0].data_ptr(), Q[1].data_ptr(), d.data_ptr(), d.data_ptr() + 4
    h = _lib.handle_for(Q[0].device)
    lib = _lib.load_library()
    ws = _lib.workspace(Q[0].device, lib.psgd_kron_workspace_bytes(h, C.byref(k)))
    rc = lib.psgd_kron_balance(h, C.byref(k), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(Q[0].device))
    _lib.check(h, rc, "psgd_kron_balance")


def _bound(A, V0, k, half_iters, spd):
    if k != _K_PROBES or half_iters != 2:
        raise NotImplementedError("the engine implements the configuration the reference always uses: k=32, half_iters=2")
    if A.dim() != 2 or A.shape[0] != A.shape[1] or not A.is_cuda:
        raise EngineError("A must be a square CUDA matrix")
    A = A.contiguous()
    s = A.shape[0]
    if V0 is None:
        V0 = torch.randn(k, s, dtype=A.dtype, device=A.device)  # psgd.py:62 / 87
    out = torch.empty([], dtype=torch.float32, device=A.device)
    h = _lib.handle_for(A.device)
    lib = _lib.load_library()
    dt = _lib.dtype_code(A)
    ws = _lib.workspace(A.device, lib.psgd_helper_workspace_bytes(h, s, dt))
    fn = lib.psgd_norm_lower_bound_spd if spd else lib.psgd_norm_lower_bound_skh
    rc = fn(h, dt, _lib.ptr(A), s, _lib.ptr(V0.contiguous()), _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(A.device))
    _lib.check(h, rc, "psgd_norm_lower_bound")
    return out.to(A.dtype)


def norm_lower_bound_spd(A, k=32, half_iters=2, V0=None):
    """psgd.py:46-68"""
    return _bound(A, V0, k, half_iters, True)


def norm_lower_bound_skh(A, k=32, half_iters=2, V0=None):
    """psgd.py:71-93"""
    return _bound(A, V0, k, half_iters, False)


norm4 = norm_lower_bound_spd  # the north_star's name for the spectral-norm bound


def procrustes_step2(Q, max_step_size=1 / 8, V0=None):
    """psgd.py:101-124, in place on Q."""
    if Q.dim() != 2 or Q.shape[0] != Q.shape[1] or not Q.is_cuda or not Q.is_contiguous():
        raise EngineError("Q must be a contiguous square CUDA matrix")
    s = Q.shape[0]
    if V0 is None:
        V0 = torch.randn(_K_PROBES, s, dtype=Q.dtype, device=Q.device)
    h = _lib.handle_for(Q.device)
    lib = _lib.load_library()
    dt = _lib.dtype_code(Q)
    ws = _lib.workspace(Q.device, lib.psgd_helper_workspace_bytes(h, s, dt))
    rc = lib.psgd_procrustes_step2(h, dt, _lib.ptr(Q), s, _lib.ptr(V0.contiguous()), float(max_step_size), _lib.ptr(ws), ws.numel(),
                                   _lib.stream_ptr(Q.device))
    _lib.check(h, rc, "psgd_procrustes_step2")


def gemm(A, B, trans_a=False, trans_b=False, alpha=1.0, D=None, beta=0.0, out_dtype=None, path=0):
    """C = alpha * op(A) op(B) (+ beta * D) through the engine's GEMM (building block; tests and benchmarks)."""
    A, B = A.contiguous(), B.contiguous()
    M, K = (A.shape[1], A.shape[0]) if trans_a else A.shape
    N = B.shape[0] if trans_b else B.shape[1]
    out_dtype = out_dtype or A.dtype
    Cm = torch.empty(M, N, dtype=out_dtype, device=A.device)
    h = _lib.handle_for(A.device)
    lib = _lib.load_library()
    odt = _lib.PSGD_BF16 if out_dtype == torch.bfloat16 else _lib.PSGD_F32
    rc = lib.psgd_gemm(h, path, _lib.dtype_code(A), odt, int(trans_a), int(trans_b), M, N, K, _lib.ptr(A), A.shape[1], _lib.ptr(B),
                       B.shape[1], _lib.ptr(Cm), N, float(alpha), _lib.ptr(D.contiguous()) if D is not None else None,
                       N, float(beta), _lib.stream_ptr(A.device))
    _lib.check(h, rc, "psgd_gemm")
    return Cm


# ------------------------------------------------------------------------------------------------
# LRA
# ------------------------------------------------------------------------------------------------
def IpUVtmatvec(U, V, x):
    """psgd.py:987-991 (host-side helper kept for API parity; the engine fuses it)."""
    return x + U.mm(V.t().mm(x))


def _lra_desc(UVd, Luvd):
    U, V, d = UVd
    if not (U.is_cuda and U.is_contiguous() and V.is_contiguous() and d.is_contiguous()):
        raise EngineError("U, V, d must be contiguous CUDA tensors")
    if U.shape != V.shape or d.numel() != U.shape[0] or V.dtype != U.dtype or d.dtype != U.dtype:
        raise EngineError("inconsistent LRA state")
    l = LraT()
    l.n, l.r, l.dtype = U.shape[0], U.shape[1], _lib.dtype_code(U)
    l.U, l.V, l.d = U.data_ptr(), V.data_ptr(), d.data_ptr()
    if Luvd is not None:
        l.Lu, l.Lv, l.Ld = Luvd[0].data_ptr(), Luvd[1].data_ptr(), Luvd[2].data_ptr()
    else:
        z = _dummy(U.device)
        l.Lu = l.Lv = l.Ld = z.data_ptr()
    return l


def update_precond_lra(UVd, Luvd, v, h, lr=0.1, betaL=0.9, update_U=None):
    """psgd.py:994-1052, in place. `update_U` overrides the coin flip torch.rand([]) < 0.5 of line 1035."""
    if update_U is None:
        update_U = bool(torch.rand([]) < 0.5)
    l = _lra_desc(UVd, Luvd)
    dev = UVd[0].device
    hd = _lib.handle_for(dev)
    lib = _lib.load_library()
    ws = _lib.workspace(dev, lib.psgd_lra_workspace_bytes(hd, C.byref(l)))
    rc = lib.psgd_lra_update(hd, C.byref(l), _lib.ptr(v.contiguous()), _lib.ptr(h.contiguous()), float(lr), float(betaL),
                             int(update_U), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(hd, rc, "psgd_lra_update")


def update_precond_lra_whiten(UVd, Luvd, g, lr=0.1, betaL=0.9, damping=1e-9, noise=None):
    """psgd.py:1066-1072. RNG order: randn_like(g) (1070) then the CPU coin of update_precond_lra (1035)."""
    if noise is None:
        noise = {"v": torch.randn_like(g), "update_U": bool(torch.rand([]) < 0.5)}
    l = _lra_desc(UVd, Luvd)
    dev = UVd[0].device
    hd = _lib.handle_for(dev)
    lib = _lib.load_library()
    ws = _lib.workspace(dev, lib.psgd_lra_workspace_bytes(hd, C.byref(l)))
    rc = lib.psgd_lra_whiten_update(hd, C.byref(l), _lib.ptr(g.contiguous()), _lib.ptr(noise["v"].contiguous()), float(lr),
                                    float(betaL), float(damping), int(noise["update_U"]), _lib.ptr(ws), ws.numel(),
                                    _lib.stream_ptr(dev))
    _lib.check(hd, rc, "psgd_lra_whiten_update")


def precond_grad_lra(UVd, g, sumsq_out=None):
    """psgd.py:1055-1063: returns d * (I + V U^T) (I + U V^T) (d * g)."""
    l = _lra_desc(UVd, None)
    dev = UVd[0].device
    hd = _lib.handle_for(dev)
    lib = _lib.load_library()
    ws = _lib.workspace(dev, lib.psgd_lra_workspace_bytes(hd, C.byref(l)))
    g = g.contiguous()
    out = torch.empty_like(g)
    rc = lib.psgd_lra_precond_grad(hd, C.byref(l), _lib.ptr(g), _lib.ptr(out), _lib.ptr(sumsq_out), _lib.ptr(ws), ws.numel(),
                                   _lib.stream_ptr(dev))
    _lib.check(hd, rc, "psgd_lra_precond_grad")
    return out


def update_precond_lra_newton(UVd, Luvd, v, h, lr=0.1, betaL=0.9, damping=1e-9, update_U=None, noise=None):
    """psgd.py:1193-1198: LRA Newton update = update_precond_lra on (v, h + (damping + eps|h|) * randn_like(h)): independent noise on the
    Hessian-vector product.  RNG order: randn_like(h) then the CPU coin of update_precond_lra.  The damped Hvp is formed by the engine
    (psgd_lra_newton_update: the damping pass of the whitening update with a separate probe); `noise` = {"z", "update_U"} replays draws."""
    if noise is None:
        noise = {"z": torch.randn_like(h), "update_U": bool(torch.rand([]) < 0.5) if update_U is None else bool(update_U)}
    l = _lra_desc(UVd, Luvd)
    dev = UVd[0].device
    hd = _lib.handle_for(dev)
    lib = _lib.load_library()
    ws = _lib.workspace(dev, lib.psgd_lra_workspace_bytes(hd, C.byref(l)))
    rc = lib.psgd_lra_newton_update(hd, C.byref(l), _lib.ptr(v.contiguous()), _lib.ptr(h.contiguous()), _lib.ptr(noise["z"].contiguous()),
                                    float(lr), float(betaL), float(damping), int(noise["update_U"]), _lib.ptr(ws), ws.numel(),
                                    _lib.stream_ptr(dev))
    _lib.check(hd, rc, "psgd_lra_newton_update")


# north_star's names for the LRA functions (old.py:657,744): thin aliases of the psgd.py math
update_precond_UVd = update_precond_lra
precond_grad_UVd = precond_grad_lra


# ------------------------------------------------------------------------------------------------
# The other geometries and the Newton-pair updates -> psgd_kron_update
# ------------------------------------------------------------------------------------------------
class NoiseTape:
    """Random draws of one update in the reference's draw order.  Default: drawn lazily from the torch generators exactly where the
    reference draws (randn_like(G), randn(32, s) per norm bound, the CPU coin torch.rand([])); a list of pre-drawn items replays them
    (parity tests feed the numbers the reference consumed)."""

    def __init__(self, items=None, device=None, rounds=None):
        self.replay = items is not None
        self.items = list(items) if items is not None else []
        self.pos = 0
        self.device = device
        # PRO4P: procrustes_step3 round counts per dense factor (psgd.py:444-449).  `rounds` pins them (the stopping test is skipped) so
        # that a bf16 parity test runs the rounds the reference ran; `self.rounds` records what was done.
        self.fixed_rounds = list(rounds) if rounds is not None else None
        self.rounds = []

    def _next(self, make):
        if self.replay:
            v = self.items[self.pos]
            self.pos += 1
            if isinstance(v, torch.Tensor) and self.device is not None:
                v = v.to(self.device)
            return v
        v = make()
        self.items.append(v)
        return v

    def randn_like(self, x):
        return self._next(lambda: torch.randn_like(x))

    def randn(self, k, s, like):
        return self._next(lambda: torch.randn(k, s, dtype=like.dtype, device=like.device))

    def rand(self):
        return self._next(lambda: float(torch.rand([])))


def _as_tape(noise, device):
    if isinstance(noise, NoiseTape):
        return noise
    return NoiseTape(noise, device=device)


def _apply_factors(Q, G, sumsq_out=None):
    """exprA(*Q, G) on the engine (psgd_kron_apply_factors)."""
    if not G.is_cuda:
        raise EngineError("psgd_torch_b200 runs on CUDA (sm_100a) tensors only")
    G = G.contiguous()
    if G.dim() > 2:
        X = G
        for i, q in enumerate(Q):
            if q.dim() == 2:
                X = _mode_product(q, X, i, False)
            else:
                shp = [1] * X.dim()
                shp[i] = -1
                X = X * q.reshape(shp)
        return X.contiguous()
    k = _kron_desc(Q, None, G)
    d = _dummy(G.device)
    k.LL, k.LR = d.data_ptr(), d.data_ptr() + 4
    h = _lib.handle_for(G.device)
    lib = _lib.load_library()
    ws = _lib.workspace(G.device, lib.psgd_kron_update_workspace_bytes(h, C.byref(k), _lib.DQ_CODES["QUAD4P"]))
    out = torch.empty_like(G)
    rc = lib.psgd_kron_apply_factors(h, C.byref(k), _lib.ptr(G), _lib.ptr(out), _lib.ptr(sumsq_out), _lib.ptr(ws), ws.numel(),
                                     _lib.stream_ptr(G.device))
    _lib.check(h, rc, "psgd_kron_apply_factors")
    return out


def solve_kron_factors(Q, V):
    """conjB of psgd.py:297-303 for real tensors: kron_i(Q_i^{-T}) V with upper-triangular dense factors (fp32-accurate blocked
    triangular inverse on the engine) and divisions for diagonal factors.  Exposed for tests."""
    V = V.contiguous()
    if V.dim() > 2:
        raise NotImplementedError("triangular solves for tensors of order >= 3 are not built")
    k = _kron_desc(Q, None, V)
    d = _dummy(V.device)
    k.LL, k.LR = d.data_ptr(), d.data_ptr() + 4
    h = _lib.handle_for(V.device)
    lib = _lib.load_library()
    ws = _lib.workspace(V.device, lib.psgd_kron_update_workspace_bytes(h, C.byref(k), _lib.DQ_CODES["EQ"]))
    out = torch.empty_like(V)
    rc = lib.psgd_kron_solve_factors(h, C.byref(k), _lib.ptr(V), _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(V.device))
    _lib.check(h, rc, "psgd_kron_solve_factors")
    return out


def procrustes_step3(Q, max_step_size=1 / 3, V0=None):
    """psgd.py:127-155, in place on Q (the branch of line 149 is taken on the device)."""
    if Q.dim() != 2 or Q.shape[0] != Q.shape[1] or not Q.is_cuda or not Q.is_contiguous():
        raise EngineError("Q must be a contiguous square CUDA matrix")
    s = Q.shape[0]
    if V0 is None:
        V0 = torch.randn(_K_PROBES, s, dtype=Q.dtype, device=Q.device)  # psgd.py:87 via 142
    h = _lib.handle_for(Q.device)
    lib = _lib.load_library()
    dt = _lib.dtype_code(Q)
    # its own scratch: a staged PRO4P update keeps the Grams of the next factor in the shared workspace
    ws = torch.empty(lib.psgd_helper_workspace_bytes(h, s, dt), dtype=torch.uint8, device=Q.device)
    rc = lib.psgd_procrustes_step3(h, dt, _lib.ptr(Q), s, _lib.ptr(V0.contiguous()), float(max_step_size), _lib.ptr(ws), ws.numel(),
                                   _lib.stream_ptr(Q.device))
    _lib.check(h, rc, "psgd_procrustes_step3")


def _almost_symmetric(q):
    """(q.H - q).abs().amax() < 0.001 * q.abs().amax()  (psgd.py:448): two device reductions, one host read like the reference."""
    s = q.shape[0]
    h = _lib.handle_for(q.device)
    lib = _lib.load_library()
    dt = _lib.dtype_code(q)
    ws = torch.empty(lib.psgd_helper_workspace_bytes(h, s, dt), dtype=torch.uint8, device=q.device)
    out = torch.empty(2, dtype=torch.float32, device=q.device)
    rc = lib.psgd_symmetry_gap(h, dt, _lib.ptr(q), s, _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(q.device))
    _lib.check(h, rc, "psgd_symmetry_gap")
    gap, amax = out.tolist()
    return gap < 0.001 * amax


def _pro4p_rounds(q, tape):
    """psgd.py:444-449: rotate until the factor is almost symmetric (at most ten rounds; host-side branch like the reference's)."""
    fixed = tape.fixed_rounds.pop(0) if tape.fixed_rounds is not None else None
    done = 0
    for _ in range(10 if fixed is None else fixed):
        procrustes_step3(q, V0=tape.randn(_K_PROBES, q.shape[1], q))
        done += 1
        if fixed is None and _almost_symmetric(q):
            break
    tape.rounds.append(done)


def _kron_update(dQ, QL, X, V, lr, betaL, damping, noise, damp=True):
    """One update through psgd_kron_update.  Draw order = the reference's: randn_like(X) (if damped), then per dense factor the
    norm-bound probe randn(32, s) (psgd.py:62) followed by that factor's procrustes probes (Q0.5EQ1.5: one, psgd.py:87; PRO4P: one
    per procrustes_step3 round), finally the balancing coin (none for QEP)."""
    Q, L = QL
    if not X.is_cuda:
        raise EngineError("psgd_torch_b200 runs on CUDA (sm_100a) tensors only")
    X = X.contiguous()
    if V is not None:
        V = V.contiguous()
        if V.shape != X.shape or V.dtype != X.dtype:
            raise EngineError("V and Hvp must agree in shape and dtype")
    tape = _as_tape(noise, X.device)
    if X.dim() > 2:
        return _kron_update_nd(dQ, QL, X, V, lr, betaL, damping, tape, damp)
    code = _lib.DQ_CODES[dQ]
    k = _kron_desc(Q, L, X)
    nz = KronNoiseT()
    keep = []
    if damp:
        N = tape.randn_like(X)
        keep.append(N)
        nz.N = N.data_ptr()
    h = _lib.handle_for(X.device)
    lib = _lib.load_library()
    ws = _lib.workspace(X.device, lib.psgd_kron_update_workspace_bytes(h, C.byref(k), code))

    def call(stages):
        rc = lib.psgd_kron_update(h, C.byref(k), code, _lib.ptr(X), _lib.ptr(V), float(lr), float(betaL), float(damping), C.byref(nz),
                                  stages, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(X.device))
        _lib.check(h, rc, "psgd_kron_update")

    fields = (("V0_spd_l", "V0_skh_l", _lib.STAGE_FACTOR_L), ("V0_spd_r", "V0_skh_r", _lib.STAGE_FACTOR_R))
    staged = dQ == "PRO4P"
    stages = _lib.STAGE_PREPARE
    for i, q in enumerate(Q):
        spd_f, skh_f, bit = fields[i]
        if q.dim() == 2:
            p = tape.randn(_K_PROBES, q.shape[1], q)
            keep.append(p)
            setattr(nz, spd_f, p.data_ptr())
            if code == 0:
                p2 = tape.randn(_K_PROBES, q.shape[1], q)
                keep.append(p2)
                setattr(nz, skh_f, p2.data_ptr())
        stages |= bit
        if staged:  # psgd.py:444-449: rotate until the factor is almost symmetric; host-side branch like the reference
            call(stages)
            stages = 0
            if q.dim() == 2:
                _pro4p_rounds(q, tape)
    if dQ != "QEP" and tape.rand() < 0.01:  # psgd.py:318, 390, 418 ...
        stages |= _lib.STAGE_BALANCE
    if stages:
        call(stages)
    del keep


def _unfold(T, i):
    return T.movedim(i, 0).reshape(T.shape[i], -1).contiguous()


def _factor_on_mode(q, T, i):
    """exprQs[i](q, T): psgd.py:225-226 / 245-246"""
    if q.dim() == 2:
        return _mode_product(q, T, i, False)
    shp = [1] * T.dim()
    shp[i] = -1
    return T * q.reshape(shp)


def _solve_factors_nd(Q, V):
    """conjB of psgd.py:297-303 for order >= 3: one engine triangular solve per dense mode (on the mode-i unfolding), divisions for the rest."""
    X = V
    for i, q in enumerate(Q):
        if q.dim() == 2:
            Xi = X.movedim(i, 0)
            shp = Xi.shape
            M = Xi.reshape(shp[0], -1).contiguous()
            ones = torch.ones(M.shape[1], dtype=M.dtype, device=M.device)
            X = solve_kron_factors([q, ones], M).reshape(shp).movedim(0, i)
        else:
            shp = [1] * X.dim()
            shp[i] = -1
            X = X / q.reshape(shp)
    return X.contiguous()


def _kron_update_nd(dQ, QL, X, V, lr, betaL, damping, tape, damp):
    """Any geometry, whitening or Newton pair, for tensors of order >= 3: host-side composition -- the contractions
    are engine GEMMs on mode unfoldings (permutes / elementwise glue by torch), each factor's bound + Lipschitz update + step is one
    engine call (psgd_kron_factor_step).  Same draw order as the order <= 2 path."""
    Q, L = QL
    lib = _lib.load_library()
    h = _lib.handle_for(X.device)
    dt = _lib.dtype_code(X)
    code = _lib.DQ_CODES[dQ]
    newton = V is not None
    H = X
    if damp:
        N = tape.randn_like(X)
        H = X + (damping + torch.finfo(X.dtype).eps * X.abs()) * N
        if dQ == "EQ" and not newton:
            V = N                                   # psgd.py:334: the probe is the damping noise
    if dQ == "QEP":
        balance_kron_precond(Q)                     # psgd.py:347 / 674
    Pg = _apply_factors(Q, H) if dQ in ("EQ", "QUAD4P", "PRO4P") else _apply_kron_nd(Q, H)
    conjB = _solve_factors_nd(Q, V) if dQ == "EQ" else None
    numel = X.numel()
    for i, q in enumerate(Q):
        s, dense = q.shape[0], q.dim() == 2
        t2 = float(numel / s)
        src1 = _factor_on_mode(q, Pg, i) if dQ == "QEP" else Pg
        src2 = conjB if dQ == "EQ" else ((_factor_on_mode(q, V, i) if dQ == "QEP" else V) if newton else None)
        M1 = _unfold(src1, i)
        if dense:
            term1 = gemm(M1, M1, trans_b=True)
            if src2 is not None:
                M2 = _unfold(src2, i)
                term2 = gemm(M2, M2, trans_b=True)
            else:
                term2 = gemm(q, q, trans_b=True, alpha=t2) if dQ == "QEP" else None   # psgd.py:361
            spd = tape.randn(_K_PROBES, s, q)
            skh = tape.randn(_K_PROBES, s, q) if code == 0 else None
            ws = torch.empty(lib.psgd_helper_workspace_bytes(h, s, dt), dtype=torch.uint8, device=X.device)
            rc = lib.psgd_kron_factor_step(h, dt, code, _lib.PSGD_DENSE, s, _lib.ptr(q), _lib.ptr(L[i]), _lib.ptr(term1), _lib.ptr(term2), t2,
                                           float(lr), float(betaL), _lib.ptr(spd), _lib.ptr(skh), _lib.ptr(ws), ws.numel(),
                                           _lib.stream_ptr(X.device))
            _lib.check(h, rc, "psgd_kron_factor_step")
            if dQ == "PRO4P":                        # psgd.py:444-449
                _pro4p_rounds(q, tape)
        else:
            term1 = torch.sum(M1.float() * M1.float(), dim=1).contiguous()
            if src2 is not None:
                M2 = _unfold(src2, i)
                term2 = torch.sum(M2.float() * M2.float(), dim=1).contiguous()
            else:
                term2 = (t2 * q.float() * q.float()).contiguous() if dQ == "QEP" else None   # psgd.py:357
            rc = lib.psgd_kron_factor_step(h, dt, code, _lib.PSGD_DIAG, s, _lib.ptr(q), _lib.ptr(L[i]), _lib.ptr(term1), _lib.ptr(term2), t2,
                                           float(lr), float(betaL), None, None, None, 0, _lib.stream_ptr(X.device))
            _lib.check(h, rc, "psgd_kron_factor_step")
    if dQ != "QEP" and tape.rand() < 0.01:
        balance_kron_precond(Q)


def update_precond_kron_eq(QL, exprs, V, Hvp, lr=0.1, betaL=0.9, noise=None):
    """psgd.py:278-319: the raw dQ = E*Q update with the pair (V, Hvp); no damping."""
    _kron_update("EQ", QL, Hvp, V, lr, betaL, 0.0, noise, damp=False)


def _whiten(dQ):
    def f(QL, exprs, G, lr=0.1, betaL=0.9, damping=1e-9, noise=None):
        _kron_update(dQ, QL, G, None, lr, betaL, damping, noise)
    return f


def _newton(dQ):
    def f(QL, exprs, V, Hvp, lr=0.1, betaL=0.9, damping=1e-9, noise=None):
        _kron_update(dQ, QL, Hvp, V, lr, betaL, damping, noise)
    return f


for _n, _dq in (("eq", "EQ"), ("qep", "QEP"), ("qeq", "QEQ"), ("pro4p", "PRO4P"), ("quad", "QUAD"), ("quad4p", "QUAD4P")):
    globals()[f"update_precond_kron_whiten_{_n}"] = _whiten(_dq)       # psgd.py:330-391, 422-513
    globals()[f"update_precond_kron_whiten_{_n}"].__name__ = f"update_precond_kron_whiten_{_n}"
for _n, _dq in (("eq", "EQ"), ("qep", "QEP"), ("qeq", "QEQ"), ("q0p5eq1p5", "Q0.5EQ1.5"), ("pro4p", "PRO4P"), ("quad", "QUAD"),
                ("quad4p", "QUAD4P")):
    globals()[f"update_precond_kron_newton_{_n}"] = _newton(_dq)       # psgd.py:657-829
    globals()[f"update_precond_kron_newton_{_n}"].__name__ = f"update_precond_kron_newton_{_n}"


# closure-style optimizer classes of the reference (psgd.py:516-654, 832-978, 1075-1190, 1201-1330): re-authored host-side glue on top of
# the functions above (psgd_torch_b200/closure_optim.py); resolved lazily to keep this module importable on its own
def __getattr__(name):
    if name in ("KronWhiten", "KronNewton", "LRAWhiten", "LRANewton"):
        from . import closure_optim
        return getattr(closure_optim, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
