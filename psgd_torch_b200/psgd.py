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


def init_kron(t, Scale=1.0, max_size=float("inf"), max_skew=1.0, dQ="Q0.5EQ1.5"):
    """psgd.py:161-263. Returns [[Q, L], exprs] with the reference's state layout: Q[i] = scale*eye(s) (dense) or
    scale*ones(s) (diagonal, psgd.py:208) in t's dtype, L[i] = fp32 0-dim zeros, exprs = (exprP, exprGs)."""
    if dQ not in ("Q0.5EQ1.5", "Q0p5EQ1p5"):
        raise NotImplementedError(f"dQ={dQ!r}: only the Q0.5EQ1.5 geometry (what KWNS4 runs, ddp.py:84-86) is built; "
                                  "the other geometries are SURVEY.md 8(f) items 1 and 3")
    shape = t.shape
    if len(shape) == 0:  # psgd.py:189-195
        Q = [Scale * torch.ones_like(t)]
        L = [lift2single(torch.zeros_like(t))]
        return [[Q, L], (_ExprP(0), (_ExprG(0, False, 0),))]
    if len(shape) > 26:
        raise ValueError(f"Got tensor with dim {len(t.shape)}; einsum runs out of letters; replace 26 with larger numbers.")
    scale = Scale ** (1 / len(shape))
    Q, L, exprGs = [], [], []
    for i, size in enumerate(shape):
        L.append(lift2single(torch.zeros([], dtype=t.dtype, device=t.device)))
        if size <= 1 or size > max_size or size ** 2 > max_skew * t.numel():
            Q.append(scale * torch.ones(size, dtype=t.dtype, device=t.device))
            exprGs.append(_ExprG(i, False, len(shape)))
        else:
            Q.append(scale * torch.eye(size, dtype=t.dtype, device=t.device))
            exprGs.append(_ExprG(i, True, len(shape)))
    return [[Q, L], (_ExprP(len(shape)), tuple(exprGs))]


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
        noise = draw_kron_noise(G, Q)
    if G.dim() > 2:
        return _update_kron_nd(Q, L, G, lr, betaL, damping, noise)
    k = _kron_desc(Q, L, G)
    nz = KronNoiseT()
    nz.N = noise["N"].data_ptr()
    spd, skh = noise["spd"], noise["skh"]
    nz.V0_spd_l = spd[0].data_ptr() if spd[0] is not None else None
    nz.V0_skh_l = skh[0].data_ptr() if skh[0] is not None else None
    if len(Q) > 1:
        nz.V0_spd_r = spd[1].data_ptr() if spd[1] is not None else None
        nz.V0_skh_r = skh[1].data_ptr() if skh[1] is not None else None
    h = _lib.handle_for(G.device)
    lib = _lib.load_library()
    nbytes = lib.psgd_kron_workspace_bytes(h, C.byref(k))
    ws = _lib.workspace(G.device, nbytes)
    rc = lib.psgd_kron_whiten_q0p5eq1p5_update(h, C.byref(k), _lib.ptr(G), float(lr), float(betaL), float(damping), C.byref(nz),
                                               int(bool(noise.get("balance", False))), _lib.ptr(ws), ws.numel(),
                                               _lib.stream_ptr(G.device))
    _lib.check(h, rc, "psgd_kron_whiten_q0p5eq1p5_update")


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


def update_precond_lra_newton(UVd, Luvd, v, h, lr=0.1, betaL=0.9, damping=1e-9, update_U=None):
    """psgd.py:1193-1198: LRA Newton update = update_precond_lra on (v, h + damping * randn_like(h)) (independent noise on the Hvp).
    RNG order: randn_like(h) then the CPU coin of update_precond_lra."""
    damping = damping + torch.finfo(h.dtype).eps * h.abs()   # psgd.py:1197
    update_precond_lra(UVd, Luvd, v, h + damping * torch.randn_like(h), lr=lr, betaL=betaL, update_U=update_U)


# north_star's names for the LRA functions (old.py:657,744): thin aliases of the psgd.py math
update_precond_UVd = update_precond_lra
precond_grad_UVd = precond_grad_lra


# ------------------------------------------------------------------------------------------------
# names the reference exports that are outside this round's scope: fail loudly, never silently fall back
# ------------------------------------------------------------------------------------------------
def _not_built(name, row):
    def f(*a, **k):
        raise NotImplementedError(f"{name} is not built yet ({row}); only the Q0.5EQ1.5 whitening path is served by the engine")
    f.__name__ = name
    return f


for _n in ("eq", "qep", "qeq", "pro4p", "quad", "quad4p"):
    globals()[f"update_precond_kron_whiten_{_n}"] = _not_built(f"update_precond_kron_whiten_{_n}", "SURVEY.md 8a K8/K9")
for _n in ("eq", "qep", "qeq", "q0p5eq1p5", "pro4p", "quad", "quad4p"):
    globals()[f"update_precond_kron_newton_{_n}"] = _not_built(f"update_precond_kron_newton_{_n}", "SURVEY.md 8a K10")
update_precond_kron_eq = _not_built("update_precond_kron_eq", "SURVEY.md 8a K8")
procrustes_step3 = _not_built("procrustes_step3", "SURVEY.md 8a K9")
