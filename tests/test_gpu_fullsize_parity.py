"""Oracle parity at the sizes bench.py times (BASELINE.json configs[1], [2]): one update + apply of every Llama-3-8B shape bucket and of
the LRA unit, engine (through the C-ABI) against the reference arithmetic.

For each case the state is first warmed by the engine (so Q != c*I, L > 0), then ONE step is run three ways from the same state, with the
same injected noise:
  * the engine (bf16, B200),
  * the oracle in bf16 on the host CPU (the reference's arithmetic: psgd.py:394-419 / 322-327 / 994-1072),
  * the oracle in fp64 on the GPU (torch ops, checker only) -- the yardstick that says how far the reference's own bf16 arithmetic is
    from the exact step.
Asserted: north_star's bf16 tolerance 1e-2 on Q / U / V / d and on the preconditioned gradient against the reference arithmetic; the
Lipschitz constants (the output of a 32-probe power iteration run in bf16) 3e-2; and the engine is never further from fp64 than 1.5x the
reference arithmetic + 2e-3.  Every measured number goes to the parity log (conftest.parity_log -> profiles/r02_parity_errors.log), incl.
the error of the STEP dQ = Q_new - Q_old itself, which is the sensitive quantity (Q moves by ~1e-2 of its norm per update, so an error
confined to the step is invisible in relerr(Q)).
"""
import os

import pytest
import torch

from conftest import check, parity_log, relerr

pytestmark = pytest.mark.gpu


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return torch.device("cuda:0")


def _noise_to(noise, dev, dtype=None):
    cv = (lambda v: None if v is None else (v.to(dev) if dtype is None else v.to(dev).to(dtype)))
    return {"N": cv(noise["N"]), "balance": noise["balance"], "spd": [cv(v) for v in noise["spd"]], "skh": [cv(v) for v in noise["skh"]]}


def _grad(m, n, seed, dev):
    """Correlated gradient (rows and columns coloured by smooth spectra) so that the fitted Q is far from a multiple of I; built on the GPU."""
    g = torch.Generator(device=dev).manual_seed(seed)
    Z = torch.randn(m, n, device=dev, generator=g)
    rs = (1.0 + 3.0 * torch.rand(m, 1, device=dev, generator=g))
    cs = (1.0 + 3.0 * torch.rand(1, n, device=dev, generator=g))
    k = 64
    A = torch.randn(m, k, device=dev, generator=g) / k ** 0.5
    B = torch.randn(k, n, device=dev, generator=g)
    return (0.01 * (rs * Z * cs + 2.0 * A @ B)).bfloat16()


KRON_SHAPES = [(4096, 4096), (4096, 14336), (14336, 4096), (1024, 4096), (128256, 4096)]


@pytest.mark.parametrize("shape", KRON_SHAPES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_kron_update_and_apply_match_oracle_at_bench_sizes(shape):
    from psgd_torch_b200 import psgd
    from oracle import psgd_oracle as orc
    dev = _dev()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    m, n = shape
    case = f"kron q0.5eq1.5 bf16 {m}x{n}"
    QLe, exprs = psgd.init_kron(torch.zeros(m, n, dtype=torch.bfloat16, device=dev))
    warm = 1 if m > 20000 else 2
    for w in range(warm):
        torch.manual_seed(10 + w)
        psgd.update_precond_kron_whiten_q0p5eq1p5(QLe, exprs, _grad(m, n, 500 + w, dev), lr=0.5)
    G = _grad(m, n, 777, dev)
    torch.manual_seed(4242)
    noise = orc.draw_kron_noise(G.cpu(), [q.cpu() for q in QLe[0]])      # CPU generator: the same numbers on every box
    noise["balance"] = False
    Q0 = [q.detach().clone() for q in QLe[0]]
    # reference arithmetic (bf16, host CPU)
    Qo = [q.detach().cpu().clone() for q in QLe[0]]
    Lo = [l.detach().cpu().clone() for l in QLe[1]]
    orc.update_precond_kron_whiten_q0p5eq1p5([Qo, Lo], G.cpu(), noise, lr=0.5, betaL=0.9, damping=1e-9)
    # fp64 yardstick (GPU, torch ops)
    Q64 = [q.detach().double() for q in QLe[0]]
    L64 = [l.detach().double() for l in QLe[1]]
    orc.update_precond_kron_whiten_q0p5eq1p5([Q64, L64], G.double(), _noise_to(noise, dev, torch.float64), lr=0.5, betaL=0.9, damping=1e-9)
    # engine
    psgd.update_precond_kron_whiten_q0p5eq1p5(QLe, exprs, G, lr=0.5, betaL=0.9, damping=1e-9, noise=_noise_to(noise, dev))
    for i, (qe, qo, q64, q0) in enumerate(zip(QLe[0], Qo, Q64, Q0)):
        kind = "dense" if qe.dim() == 2 else "diag"
        check(case, f"Q[{i}] ({kind})", qe, qo, 1e-2, yard=q64)
        # the step itself: engine and reference arithmetic each against the fp64 step
        d64 = q64 - q0.double()
        de, do = qe.double() - q0.double(), qo.to(dev).double() - q0.double()
        e_e, e_o = relerr(de, d64), relerr(do, d64)
        parity_log(case, f"dQ[{i}] = Q_new - Q_old ({kind}), engine vs fp64 step", e_e, None, e_o)
        assert e_e <= 1.5 * e_o + 0.1, (case, i, e_e, e_o)
    for i, (le, lo, l64) in enumerate(zip(QLe[1], Lo, L64)):
        check(case, f"L[{i}]", le, lo, 3e-2, yard=l64, floor=1e-2)
    X = _grad(m, n, 888, dev)
    Pe = psgd.precond_grad_kron(QLe, exprs, X)
    Qc = [q.detach().cpu() for q in QLe[0]]
    Po = orc.precond_grad_kron(Qc, X.cpu())
    P64 = orc.precond_grad_kron([q.detach().double() for q in QLe[0]], X.double())
    check(case, "precond_grad", Pe, Po, 1e-2, yard=P64)


@pytest.mark.parametrize("n_log2,r", [(24, 32)])
def test_lra_update_and_apply_match_oracle_at_2p24(n_log2, r):
    """psgd.py:994-1072 at n = 2^24, r = 32 bf16 (the tensor-core sweeps + the bulk-copy apply), one U step and one V step."""
    from psgd_torch_b200 import psgd
    from oracle import psgd_oracle as orc
    dev = _dev()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    n = 1 << n_log2
    bf = torch.bfloat16
    case = f"lra whiten bf16 n=2^{n_log2} r={r}"
    g0 = torch.Generator(device=dev).manual_seed(5)
    sc = (0.1 / (n * r)) ** 0.5          # psgd.py:1115-1118: ||U||_F = ||V||_F = sqrt(0.1)
    UVe = [(sc * torch.randn(n, r, device=dev, generator=g0)).to(bf), (sc * torch.randn(n, r, device=dev, generator=g0)).to(bf),
           torch.ones(n, 1, device=dev, dtype=bf)]
    Le = [torch.zeros([], device=dev) for _ in range(3)]
    scale = (1.0 + (torch.arange(n, device=dev).reshape(n, 1) % 11)).float()

    def grad(seed):
        gg = torch.Generator(device=dev).manual_seed(seed)
        return (0.01 * scale * torch.randn(n, 1, device=dev, generator=gg)).to(bf)

    def probe(seed):
        gg = torch.Generator(device=dev).manual_seed(seed)
        return torch.randn(n, 1, device=dev, generator=gg).to(bf)

    for w in range(2):   # warm: one U step, one V step
        psgd.update_precond_lra_whiten(UVe, Le, grad(20 + w), lr=0.1, noise={"v": probe(30 + w), "update_U": w == 0})
    for step, upd_u in enumerate((True, False)):
        g, v = grad(40 + step), probe(50 + step)
        UVo = [x.detach().cpu().clone() for x in UVe]
        Lo = [l.detach().cpu().clone() for l in Le]
        orc.update_precond_lra_whiten(UVo, Lo, g.cpu(), {"v": v.cpu(), "update_U": upd_u}, lr=0.1, betaL=0.9, damping=1e-9)
        UV64 = [x.detach().double() for x in UVe]
        L64 = [l.detach().double() for l in Le]
        orc.update_precond_lra_whiten(UV64, L64, g.double(), {"v": v.double(), "update_U": upd_u}, lr=0.1, betaL=0.9, damping=1e-9)
        psgd.update_precond_lra_whiten(UVe, Le, g, lr=0.1, betaL=0.9, damping=1e-9, noise={"v": v, "update_U": upd_u})
        tag = f"{case} step {step} ({'U' if upd_u else 'V'} side)"
        for name, xe, xo, x64 in zip("UVd", UVe, UVo, UV64):
            check(tag, name, xe, xo, 1e-2, yard=x64)
        for name, le, lo, l64 in zip(("Lu", "Lv", "Ld"), Le, Lo, L64):
            if float(l64) == 0.0:
                continue
            # the reference's bf16 Lipschitz constants are far from their exact values at this length (profiles/r01_lra_rank_dtype_sweep.log);
            # the engine keeps every r-sized quantity in fp32: judged against fp64
            ee, er = relerr(le, l64), relerr(lo, l64)
            parity_log(tag, name + " vs fp64", ee, 1.5 * er + 1e-2, er)
            assert ee <= 1.5 * er + 1e-2, (tag, name, ee, er)
        Pe = psgd.precond_grad_lra(UVe, g)
        P64 = orc.precond_grad_lra([x.detach().double() for x in UVe], g.double())
        Po = orc.precond_grad_lra([x.detach().cpu() for x in UVe], g.cpu())
        check(tag, "precond_grad", Pe, Po, 1e-2, yard=P64)
        del UV64, P64
        torch.cuda.empty_cache()


@pytest.mark.parametrize("fname", ["lranewton_r8_f32.pt", "lranewton_r16_bf16.pt"])
def test_lra_newton_engine_matches_reference_golden(fname):
    """L5: psgd.update_precond_lra_newton (psgd.py:1193-1198) through psgd_lra_newton_update against the unmodified reference's outputs."""
    from conftest import load_golden
    from psgd_torch_b200 import psgd
    dev = _dev()
    case = load_golden(fname)
    bf = case["dtype"] == "torch.bfloat16"
    tol = 3e-2 if bf else 2e-5      # bf16: four accumulated steps at n = 1000 (single steps are checked below against the oracle)
    UVd = [case["U0"].clone().to(dev), case["V0"].clone().to(dev), case["d0"].clone().to(dev)]
    Luvd = [torch.zeros([], dtype=torch.float32, device=dev) for _ in range(3)]
    for si, st in enumerate(case["steps"]):
        noise = {"z": st["noise"]["z"].to(dev), "update_U": st["noise"]["update_U"]}
        psgd.update_precond_lra_newton(UVd, Luvd, st["v"].to(dev), st["h"].to(dev), lr=case["lr"], betaL=case["betaL"], damping=case["damping"],
                                       noise=noise)
        tag = f"lra newton golden {fname} step {si}"
        for name, x, xr in zip("UVd", UVd, (st["U"], st["V"], st["d"])):
            check(tag, name, x, xr, tol)
        for name, l, lr_ in zip(("Lu", "Lv", "Ld"), Luvd, st["L"]):
            if float(lr_) != 0.0:
                check(tag, name, l, lr_, tol)
        check(tag, "precond_grad", psgd.precond_grad_lra(UVd, st["h"].to(dev)), st["Pg"], tol)


@pytest.mark.parametrize("n,r,dtype", [(70000, 32, torch.bfloat16), (5000, 12, torch.float32)])
def test_lra_newton_engine_matches_oracle(n, r, dtype):
    from psgd_torch_b200 import psgd
    from oracle import psgd_oracle as orc
    dev = _dev()
    g0 = torch.Generator().manual_seed(n + r)
    sc = (0.1 / (n * r)) ** 0.5
    U, V = (3 * sc * torch.randn(n, r, generator=g0)).to(dtype), (3 * sc * torch.randn(n, r, generator=g0)).to(dtype)
    d = (1.0 + 0.2 * torch.rand(n, 1, generator=g0)).to(dtype)
    hd = 0.3 + torch.rand(n, 1, generator=g0)
    UVe = [U.to(dev), V.to(dev), d.to(dev)]
    Le = [torch.zeros([], device=dev) for _ in range(3)]
    bf = dtype == torch.bfloat16
    for step in range(4):
        v = torch.randn(n, 1, generator=g0)
        h = (hd * v).to(dtype)
        v = v.to(dtype)
        noise = {"z": torch.randn(n, 1, generator=g0).to(dtype), "update_U": step % 2 == 0}
        UVo = [x.detach().cpu().clone() for x in UVe]
        Lo = [l.detach().cpu().clone() for l in Le]
        UV64 = [x.double() for x in UVo]
        L64 = [l.double() for l in Lo]
        orc.update_precond_lra_newton(UVo, Lo, v, h, noise, lr=0.1)
        orc.update_precond_lra_newton(UV64, L64, v.double(), h.double(), {"z": noise["z"].double(), "update_U": noise["update_U"]}, lr=0.1)
        psgd.update_precond_lra_newton(UVe, Le, v.to(dev), h.to(dev), lr=0.1, noise={"z": noise["z"].to(dev), "update_U": noise["update_U"]})
        tag = f"lra newton {dtype} n={n} r={r} step {step}"
        for name, xe, xo, x64 in zip("UVd", UVe, UVo, UV64):
            check(tag, name, xe, xo, 1e-2 if bf else 1e-5, yard=x64 if bf else None)
