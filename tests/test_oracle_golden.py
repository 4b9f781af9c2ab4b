"""CPU tests: the oracle restatement (oracle/psgd_oracle.py) against the golden vectors that
tests/golden/make_golden.py produced by running the unmodified reference (psgd.py, ddp.py)."""
import glob
import os

import pytest
import torch

from conftest import GOLDEN, load_golden, relerr
from oracle import psgd_oracle as orc

# bf16: the oracle's explicit matmul order differs from the einsum path the reference ran with, which
# changes bf16 rounding only (measured <= 1.4e-2 when the fixtures were made); fp32: <= 6e-7.
TOL = {"torch.float32": 5e-6, "torch.bfloat16": 3e-2}

KRON = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "kron_*.pt")))
LRA = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "lra_*.pt")))
KWNS4 = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "kwns4_*.pt")))
GEOM = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "geom_*.pt")))


def test_fixtures_present():
    assert len(KRON) >= 10 and len(LRA) >= 3 and len(KWNS4) >= 3


@pytest.mark.parametrize("fname", KRON)
def test_kron_oracle_matches_reference(fname):
    torch.set_num_threads(1)
    case = load_golden(fname)
    tol = TOL[case["dtype"]]
    Q = [q.clone() for q in case["Q0"]]
    L = [l.clone() for l in case["L0"]]
    for st in case["steps"]:
        orc.update_precond_kron_whiten_q0p5eq1p5([Q, L], st["G"], st["noise"], lr=case["lr"], betaL=case["betaL"],
                                                 damping=case["damping"])
        for q, qr in zip(Q, st["Q"]):
            assert relerr(q, qr) < tol
        for l, lr_ in zip(L, st["L"]):
            assert relerr(l, lr_) < tol
        assert relerr(orc.precond_grad_kron(Q, st["G"]), st["Pg"]) < tol


@pytest.mark.parametrize("fname", LRA)
def test_lra_oracle_matches_reference(fname):
    torch.set_num_threads(1)
    case = load_golden(fname)
    tol = TOL[case["dtype"]]
    UVd = [case["U0"].clone(), case["V0"].clone(), case["d0"].clone()]
    Luvd = [torch.zeros([], dtype=torch.float32) for _ in range(3)]
    for st in case["steps"]:
        orc.update_precond_lra_whiten(UVd, Luvd, st["g"], st["noise"], lr=case["lr"], betaL=case["betaL"],
                                      damping=case["damping"])
        for x, xr in zip(UVd, (st["U"], st["V"], st["d"])):
            assert relerr(x, xr) < tol
        for l, lr_ in zip(Luvd, st["L"]):
            assert relerr(l, lr_) < tol
        assert relerr(orc.precond_grad_lra(UVd, st["g"]), st["Pg"]) < tol


@pytest.mark.parametrize("fname", KWNS4)
def test_kwns4_oracle_matches_reference(fname):
    torch.set_num_threads(1)
    case = load_golden(fname)
    pdtype = {"torch.float32": torch.float32, "torch.bfloat16": torch.bfloat16}[case["pdtype"]]
    p = case["p0"].clone()
    state = {}
    kw = {k: v for k, v in case["kw"].items() if k != "preconditioner_update_probability"}
    for st in case["steps"]:
        orc.kwns4_param_step(p, st["grad"].clone(), state, st["noise"], preconditioner_dtype=pdtype,
                             do_update=st["do_update"], **kw)
        assert relerr(p, st["p"]) < (1e-6 if pdtype == torch.float32 else 2e-3)


@pytest.mark.parametrize("fname", GEOM)
def test_geometry_oracle_matches_reference(fname):
    """The other Kron geometries and the Newton-pair updates: the oracle replays the stored NoiseTape."""
    torch.set_num_threads(1)
    cases = load_golden(fname)
    assert len(cases) >= 3
    for case in cases:
        tol = TOL[case["dtype"]]
        if case["dQ"] == "PRO4P" and case["dtype"] == "torch.float32":
            tol *= 10   # the procrustes_step3 loop amplifies contraction-order rounding (1.2e-5 measured when the fixtures were made)
        dq = case["dQ"]
        Q = [q.clone() for q in case["Q0"]]
        L = [torch.zeros([], dtype=torch.float32) for _ in Q]
        for st in case["steps"]:
            tape = orc.NoiseTape(st["tape"])
            if case["mode"] == "whiten":
                orc.update_precond_kron_whiten(dq, [Q, L], st["G"], tape, lr=case["lr"], betaL=case["betaL"], damping=case["damping"])
            else:
                orc.update_precond_kron_newton(dq, [Q, L], st["V"], st["Hvp"], tape, lr=case["lr"], betaL=case["betaL"],
                                               damping=case["damping"])
            assert tape.pos == len(tape.items), "the oracle must consume exactly the draws the reference made"
            for q, qr in zip(Q, st["Q"]):
                assert relerr(q, qr) < tol, (dq, case["mode"], case["shape"])
            for l, lr_ in zip(L, st["L"]):
                assert relerr(l, lr_) < tol
            assert relerr(orc.precond_grad_kron_dq(dq, Q, st["X"]), st["Pg"]) < tol


LRAN = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "lranewton_*.pt")))


@pytest.mark.parametrize("fname", LRAN)
def test_lra_newton_oracle_matches_reference(fname):
    """update_precond_lra_newton (psgd.py:1193-1198): independent damping noise on the Hessian-vector product."""
    torch.set_num_threads(1)
    case = load_golden(fname)
    tol = TOL[case["dtype"]]
    UVd = [case["U0"].clone(), case["V0"].clone(), case["d0"].clone()]
    Luvd = [torch.zeros([], dtype=torch.float32) for _ in range(3)]
    for st in case["steps"]:
        orc.update_precond_lra_newton(UVd, Luvd, st["v"], st["h"], st["noise"], lr=case["lr"], betaL=case["betaL"], damping=case["damping"])
        for x, xr in zip(UVd, (st["U"], st["V"], st["d"])):
            assert relerr(x, xr) < tol
        for l, lr_ in zip(Luvd, st["L"]):
            assert relerr(l, lr_) < tol or float(lr_) == 0.0
        assert relerr(orc.precond_grad_lra(UVd, st["h"]), st["Pg"]) < tol


def test_pro4p_bf16_fixture_records_the_reference_round_counts():
    """geom_pro4p_bf16.pt: the procrustes_step3 round counts (psgd.py:444-449) ride with every step so that the GPU parity test can pin them."""
    cases = load_golden("geom_pro4p_bf16.pt")
    assert len(cases) == 3
    for case in cases:
        for st in case["steps"]:
            n_dense = sum(1 for q in st["Q"] if q.dim() == 2)
            assert len(st["rounds"]) == n_dense and all(1 <= r <= 10 for r in st["rounds"])
            n_probe = sum(1 for x in st["tape"] if isinstance(x, torch.Tensor) and x.dim() == 2 and x.shape[0] == 32)
            assert n_probe == n_dense + sum(st["rounds"])      # one norm-bound probe per dense factor + one per rotation round


def test_norm_lower_bound_is_a_lower_bound():
    """misc/tightness_of_spectral_norm_bound.py:45-48 property: bound <= ||A||_2 (and not absurdly loose)."""
    g = torch.Generator().manual_seed(0)
    for s in (8, 33, 100):
        W = torch.randn(s, s + 5, generator=g)
        A = W @ W.T / s
        b = float(orc.norm_lower_bound_spd(A, torch.randn(32, s, generator=g)))
        n2 = float(torch.linalg.matrix_norm(A, 2))
        assert 0.5 * n2 <= b <= n2 * (1 + 1e-5)
        R = torch.randn(s, s, generator=g)
        R = R - R.T
        b = float(orc.norm_lower_bound_skh(R, torch.randn(32, s, generator=g)))
        n2 = float(torch.linalg.matrix_norm(R, 2))
        assert 0.4 * n2 <= b <= n2 * (1 + 1e-5)


def test_kron_whitening_converges():
    """misc/psgd_kron_verification.py:184-205 property (statistical KAT): with G = H^{1/2}-coloured noise the fitted
    P whitens the gradient: the Gram of P g approaches identity scale. Loose threshold."""
    torch.manual_seed(0)
    m, n = 6, 9
    WL = torch.randn(m, m) / m ** 0.5 + torch.eye(m)
    WR = torch.randn(n, n) / n ** 0.5 + torch.eye(n)
    QL = orc.init_kron(torch.zeros(m, n), Scale=1.0)
    N = 1500
    for i in range(N):
        G = WL @ torch.randn(m, n) @ WR
        noise = orc.draw_kron_noise(G, QL[0])
        orc.update_precond_kron_whiten_q0p5eq1p5(QL, G, noise, lr=0.3 * (1 - i / N) + 0.01, betaL=0.9, damping=0.0)
    acc = torch.zeros(m, m)
    for _ in range(300):
        G = WL @ torch.randn(m, n) @ WR
        Pg = orc.precond_grad_kron(QL[0], G)
        acc += Pg @ Pg.T / n
    acc /= 300
    # E[Pg Pg^T]/n should be close to I when P = (E[g g^T])^{-1/2}
    assert relerr(acc, torch.eye(m)) < 0.35
