"""Generate the golden input/output vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference/psgd.py and wrapped_as_torch_optimizer_for_ddp.py) on CPU through
oracle/opt_einsum_shim.py.  Only runnable where /root/reference exists (the build container); the
resulting *.pt files are committed and are what travels to the GPU box.

    python tests/golden/make_golden.py

For every case the reference call is made after torch.manual_seed(seed); the random numbers it
consumed are then re-drawn with the same seed by oracle.psgd_oracle.draw_*_noise (same call order,
shapes and dtypes) and stored, so that the oracle restatement and the CUDA engine can be driven with
exactly the numbers the reference used.
"""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import opt_einsum_shim  # noqa: E402
from oracle import psgd_oracle as orc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ref = opt_einsum_shim.load_reference("/root/reference")


def structured_grad(shape, dtype, gen_seed):
    """G = H_L^{1/2} Z H_R^{1/2}-like correlated gradient (cf. misc/psgd_kron_verification.py:185-194) so
    that the preconditioner has something to whiten."""
    g = torch.Generator().manual_seed(gen_seed)
    Z = torch.randn(*shape, generator=g)
    if len(shape) == 2:
        m, n = shape
        WL = torch.randn(m, m, generator=g) / m ** 0.5 + 0.5 * torch.eye(m)
        WR = torch.randn(n, n, generator=g) / n ** 0.5 + 0.5 * torch.eye(n)
        Z = WL @ Z @ WR
    return (0.1 * Z).to(dtype)


def kron_case(name, shape, dtype, steps=3, max_skew=1.0, lr=0.5, want_balance_step=None):
    t0 = torch.zeros(*shape, dtype=dtype)
    QL_ref, exprs = ref.init_kron(t0, Scale=1.0, max_size=float("inf"), max_skew=max_skew, dQ="Q0.5EQ1.5")
    QL_o = orc.init_kron(t0, Scale=1.0, max_size=float("inf"), max_skew=max_skew)
    for a, b in zip(QL_ref[0], QL_o[0]):
        assert torch.equal(a, b)
    case = {"name": name, "shape": list(shape), "dtype": str(dtype), "max_skew": max_skew, "lr": lr,
            "betaL": 0.9, "damping": 1e-9, "Q0": [q.clone() for q in QL_ref[0]], "L0": [l.clone() for l in QL_ref[1]],
            "steps": []}
    worst = 0.0
    for s in range(steps):
        G = structured_grad(shape, dtype, 1000 + s)
        seed = 4242 + 17 * s
        if want_balance_step == s:  # find a seed whose 4th draw triggers balance_kron_precond
            while True:
                torch.manual_seed(seed)
                if orc.draw_kron_noise(G, QL_ref[0])["balance"]:
                    break
                seed += 1
        torch.manual_seed(seed)
        ref.update_precond_kron_whiten_q0p5eq1p5(QL_ref, exprs, G, lr=lr, betaL=0.9, damping=1e-9)
        Pg = ref.precond_grad_kron(QL_ref, exprs, G)
        torch.manual_seed(seed)
        noise = orc.draw_kron_noise(G, QL_o[0])
        orc.update_precond_kron_whiten_q0p5eq1p5(QL_o, G, noise, lr=lr, betaL=0.9, damping=1e-9)
        Pg_o = orc.precond_grad_kron(QL_o[0], G)
        for a, b in zip(QL_ref[0], QL_o[0]):
            worst = max(worst, float((a.float() - b.float()).norm() / a.float().norm()))
        worst = max(worst, float((Pg.float() - Pg_o.float()).norm() / Pg.float().norm()))
        case["steps"].append({"G": G, "seed": seed, "noise": noise,
                              "Q": [q.clone() for q in QL_ref[0]], "L": [l.clone() for l in QL_ref[1]],
                              "Pg": Pg.clone()})
    print(f"{name:28s} shape={tuple(shape)} {dtype}: oracle-vs-reference worst normwise rel err {worst:.3e}")
    torch.save(case, os.path.join(OUT, f"kron_{name}.pt"))
    return worst


def lra_case(name, n, r, dtype, steps=4, lr=0.1):
    g0 = torch.Generator().manual_seed(7)
    U = torch.randn(n, r, generator=g0)
    U = (U * (0.1 ** 0.5 / torch.linalg.vector_norm(U))).to(dtype)  # psgd.py:1115-1118
    V = torch.randn(n, r, generator=g0)
    V = (V * (0.1 ** 0.5 / torch.linalg.vector_norm(V))).to(dtype)
    d = torch.ones(n, 1, dtype=dtype)
    Luvd_r = [torch.zeros([], dtype=torch.float32) for _ in range(3)]
    UVd_r = [U.clone(), V.clone(), d.clone()]
    UVd_o, Luvd_o = [U.clone(), V.clone(), d.clone()], [l.clone() for l in Luvd_r]
    case = {"name": name, "n": n, "r": r, "dtype": str(dtype), "lr": lr, "betaL": 0.9, "damping": 1e-9,
            "U0": U, "V0": V, "d0": d, "steps": []}
    worst = 0.0
    for s in range(steps):
        g = structured_grad((n, 1), dtype, 2000 + s)
        g = g * (1 + torch.arange(n).reshape(n, 1) % 7).to(dtype)
        seed = 999 + 31 * s
        torch.manual_seed(seed)
        ref.update_precond_lra_whiten(UVd_r, Luvd_r, g, lr=lr, betaL=0.9, damping=1e-9)
        Pg = ref.precond_grad_lra(UVd_r, g)
        torch.manual_seed(seed)
        noise = orc.draw_lra_noise(g)
        orc.update_precond_lra_whiten(UVd_o, Luvd_o, g, noise, lr=lr, betaL=0.9, damping=1e-9)
        Pg_o = orc.precond_grad_lra(UVd_o, g)
        for a, b in zip(UVd_r + [Pg], UVd_o + [Pg_o]):
            worst = max(worst, float((a.float() - b.float()).norm() / a.float().norm()))
        case["steps"].append({"g": g, "seed": seed, "noise": noise, "U": UVd_r[0].clone(), "V": UVd_r[1].clone(),
                              "d": UVd_r[2].clone(), "L": [l.clone() for l in Luvd_r], "Pg": Pg.clone()})
    print(f"{name:28s} n={n} r={r} {dtype}: oracle-vs-reference worst normwise rel err {worst:.3e}")
    torch.save(case, os.path.join(OUT, f"lra_{name}.pt"))
    return worst


def kwns4_case(name, shape, pdtype, steps=5, **kw):
    """Run the reference KWNS4 (ddp.py) single-process on CPU and the oracle step glue side by side."""
    sys.modules["psgd"] = ref  # ddp.py does `import psgd`
    import importlib.util
    spec = importlib.util.spec_from_file_location("kwns4_reference", "/root/reference/wrapped_as_torch_optimizer_for_ddp.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    del sys.modules["psgd"]
    g0 = torch.Generator().manual_seed(5)
    p_ref = torch.nn.Parameter(torch.randn(*shape, generator=g0))
    p_orc = p_ref.detach().clone()
    opt = mod.KWNS4([p_ref], preconditioner_dtype=pdtype, **kw)
    state = {}
    case = {"name": name, "shape": list(shape), "pdtype": str(pdtype), "kw": kw, "p0": p_ref.detach().clone(), "steps": []}
    worst = 0.0
    for s in range(steps):
        grad = structured_grad(shape, torch.float32, 3000 + s)
        seed = 31337 + s
        p_ref.grad = grad.clone()
        torch.manual_seed(seed)
        opt.step()
        # oracle side: same RNG order: group coin flip (ddp.py:110) then the update's draws
        torch.manual_seed(seed)
        do_update = bool(torch.rand([]) < kw.get("preconditioner_update_probability", 1.0))
        gq = grad.squeeze().to(pdtype) if pdtype else grad.squeeze()
        if len(state) == 0:
            Qtmp = orc.init_kron(gq, Scale=kw.get("preconditioner_init_scale", 1.0))[0]
        else:
            Qtmp = state["QL"][0]
        noise = orc.draw_kron_noise(gq, Qtmp) if do_update else None
        okw = {k: v for k, v in kw.items() if k not in ("preconditioner_update_probability",)}
        orc.kwns4_param_step(p_orc, grad.clone(), state, noise, preconditioner_dtype=pdtype, do_update=do_update, **okw)
        worst = max(worst, float((p_ref.detach() - p_orc).norm() / p_ref.detach().norm()))
        st = opt.state[p_ref]
        case["steps"].append({"grad": grad, "seed": seed, "do_update": do_update, "noise": noise,
                              "p": p_ref.detach().clone(), "Q": [q.clone() for q in st["QL"][0]],
                              "L": [l.clone() for l in st["QL"][1]], "ema": None if st["ema"] is None else st["ema"].clone()})
    print(f"{name:28s} shape={tuple(shape)} {pdtype}: oracle-vs-reference worst param rel err {worst:.3e}")
    torch.save(case, os.path.join(OUT, f"kwns4_{name}.pt"))
    return worst


DQ_FUNCS = {"EQ": "eq", "QEP": "qep", "QEQ": "qeq", "Q0.5EQ1.5": "q0p5eq1p5", "PRO4P": "pro4p", "QUAD": "quad", "QUAD4P": "quad4p"}


def spd_pair(shape, dtype, gen_seed):
    """(V, Hvp) with Hvp = H_L V H_R for SPD H_L, H_R: a Hessian-vector-product pair for the Newton-type updates."""
    g = torch.Generator().manual_seed(gen_seed)
    V = torch.randn(*shape, generator=g)
    if len(shape) == 2:
        m, n = shape
        WL = torch.randn(m, m, generator=g) / m ** 0.5
        WR = torch.randn(n, n, generator=g) / n ** 0.5
        H = (WL @ WL.T + 0.3 * torch.eye(m)) @ V @ (WR @ WR.T + 0.3 * torch.eye(n))
    else:
        d = 0.3 + torch.rand(*shape, generator=g)
        H = d * V
    return V.to(dtype), (0.5 * H).to(dtype)


def geom_cases(dq, specs=None, suffix=""):
    """All golden cases of one geometry: whitening and Newton-pair updates by the unmodified reference (psgd.py:330-513,
    657-829), with the oracle run side by side in record mode on the same seed (its NoiseTape is stored for replay)."""
    f32, bf16 = torch.float32, torch.bfloat16
    fn = DQ_FUNCS[dq]
    if specs is not None:
        return _geom_cases(dq, fn, specs, suffix)
    specs = [("whiten", (24, 40), f32, 3), ("whiten", (8, 300), f32, 3), ("whiten", (300, 8), f32, 2), ("whiten", (50,), f32, 3),
             ("whiten", (136, 200), f32, 2), ("newton", (24, 40), f32, 3), ("newton", (8, 300), f32, 2)]
    specs += [("whiten", (5, 6, 7), f32, 2)]                                   # order 3: all dense
    if dq != "PRO4P":   # (PRO4P's data-dependent procrustes_step3 loop makes this one chaotic: two contraction orders differ by 5e-4)
        specs += [("newton", (3, 4, 30), f32, 2)]                              # order 3: dense, dense, diagonal
    if dq not in ("PRO4P",):
        specs += [("whiten", (24, 40), bf16, 3), ("whiten", (136, 200), bf16, 2), ("newton", (24, 40), bf16, 2)]
    if dq == "Q0.5EQ1.5":
        specs = [s for s in specs if s[0] == "newton"]  # the whitening form has its own fixtures (kron_*.pt)
    return _geom_cases(dq, fn, specs, suffix)


def _geom_cases(dq, fn, specs, suffix):
    out = []
    for (mode, shape, dtype, steps) in specs:
        t0 = torch.zeros(*shape, dtype=dtype)
        QL_ref, exprs = ref.init_kron(t0, Scale=1.0, max_size=float("inf"), max_skew=1.0, dQ=dq)
        QL_o = orc.init_kron_dq(t0, Scale=1.0, dQ=dq)
        lr = 0.5 if dq not in ("PRO4P", "QUAD4P") else 0.2
        case = {"dQ": dq, "mode": mode, "shape": list(shape), "dtype": str(dtype), "lr": lr, "betaL": 0.9, "damping": 1e-9,
                "Q0": [q.clone() for q in QL_ref[0]], "steps": []}
        worst = 0.0
        for s in range(steps):
            seed = 777 + 13 * s
            if mode == "whiten":
                G = structured_grad(shape, dtype, 5000 + s)
                torch.manual_seed(seed)
                getattr(ref, f"update_precond_kron_whiten_{fn}")(QL_ref, exprs, G, lr=lr, betaL=0.9, damping=1e-9)
                torch.manual_seed(seed)
                tape = orc.NoiseTape()
                orc.update_precond_kron_whiten(dq, QL_o, G, tape, lr=lr, betaL=0.9, damping=1e-9)
                inputs = {"G": G}
            else:
                V, Hvp = spd_pair(shape, dtype, 6000 + s)
                torch.manual_seed(seed)
                getattr(ref, f"update_precond_kron_newton_{fn}")(QL_ref, exprs, V, Hvp, lr=lr, betaL=0.9, damping=1e-9)
                torch.manual_seed(seed)
                tape = orc.NoiseTape()
                orc.update_precond_kron_newton(dq, QL_o, V, Hvp, tape, lr=lr, betaL=0.9, damping=1e-9)
                inputs = {"V": V, "Hvp": Hvp}
            X = structured_grad(shape, dtype, 7000 + s)
            Pg = exprs[0](*QL_ref[0], X) if dq in ("PRO4P", "QUAD4P") else ref.precond_grad_kron(QL_ref, exprs, X)
            Pg_o = orc.precond_grad_kron_dq(dq, QL_o[0], X)
            for a, b in zip(list(QL_ref[0]) + [Pg], list(QL_o[0]) + [Pg_o]):
                worst = max(worst, float((a.float() - b.float()).norm() / a.float().norm()))
            for a, b in zip(QL_ref[1], QL_o[1]):
                worst = max(worst, float((a - b).abs() / a.abs()))
            case["steps"].append({**inputs, "seed": seed, "tape": list(tape.items), "X": X, "Q": [q.clone() for q in QL_ref[0]],
                                  "L": [l.clone() for l in QL_ref[1]], "Pg": Pg.clone(), "rounds": list(tape.rounds)})
        print(f"geom {dq:10s} {mode:6s} shape={tuple(shape)} {dtype}: oracle-vs-reference worst rel err {worst:.3e}")
        case["oracle_vs_reference"] = worst
        out.append(case)
    torch.save(out, os.path.join(OUT, f"geom_{fn}{suffix}.pt"))


def lra_newton_case(name, n, r, dtype, steps=4, lr=0.1):
    """update_precond_lra_newton (psgd.py:1193-1198) by the unmodified reference on (v, Hvp) pairs of a diagonal-plus-low-rank SPD
    Hessian; the oracle side by side on the same seed."""
    g0 = torch.Generator().manual_seed(11)
    U = torch.randn(n, r, generator=g0)
    U = (U * (0.1 ** 0.5 / torch.linalg.vector_norm(U))).to(dtype)
    V = torch.randn(n, r, generator=g0)
    V = (V * (0.1 ** 0.5 / torch.linalg.vector_norm(V))).to(dtype)
    d = torch.ones(n, 1, dtype=dtype)
    hd = 0.3 + torch.rand(n, 1, generator=g0)
    W = torch.randn(n, 3, generator=g0) / n ** 0.5
    UVd_r, Luvd_r = [U.clone(), V.clone(), d.clone()], [torch.zeros([], dtype=torch.float32) for _ in range(3)]
    UVd_o, Luvd_o = [U.clone(), V.clone(), d.clone()], [torch.zeros([], dtype=torch.float32) for _ in range(3)]
    case = {"name": name, "n": n, "r": r, "dtype": str(dtype), "lr": lr, "betaL": 0.9, "damping": 1e-9, "U0": U, "V0": V, "d0": d,
            "steps": []}
    worst = 0.0
    for s in range(steps):
        v = torch.randn(n, 1, generator=g0)
        h = (hd * v + W @ (W.T @ v)).to(dtype)
        v = v.to(dtype)
        seed = 555 + 29 * s
        torch.manual_seed(seed)
        ref.update_precond_lra_newton(UVd_r, Luvd_r, v, h, lr=lr, betaL=0.9, damping=1e-9)
        Pg = ref.precond_grad_lra(UVd_r, h)
        torch.manual_seed(seed)
        noise = orc.draw_lra_newton_noise(h)
        orc.update_precond_lra_newton(UVd_o, Luvd_o, v, h, noise, lr=lr, betaL=0.9, damping=1e-9)
        Pg_o = orc.precond_grad_lra(UVd_o, h)
        for a, b in zip(UVd_r + [Pg], UVd_o + [Pg_o]):
            worst = max(worst, float((a.float() - b.float()).norm() / a.float().norm()))
        case["steps"].append({"v": v, "h": h, "seed": seed, "noise": noise, "U": UVd_r[0].clone(), "V": UVd_r[1].clone(),
                              "d": UVd_r[2].clone(), "L": [l.clone() for l in Luvd_r], "Pg": Pg.clone()})
    print(f"{name:28s} n={n} r={r} {dtype}: oracle-vs-reference worst normwise rel err {worst:.3e}")
    case["oracle_vs_reference"] = worst
    torch.save(case, os.path.join(OUT, f"lranewton_{name}.pt"))
    return worst


def kwns4_checkpoint_case(name, shape, pdtype, ptype, save_at=3, steps=5, **kw):
    """A checkpoint WRITTEN BY THE REFERENCE WRAPPER (ddp.py:131-137 state keys) in the middle of a run, plus the rest of the reference's
    trajectory: the drop-in must load it and continue like the reference does.  `exprs` (closures of the opt_einsum stand-in; with the real
    opt_einsum: ContractExpression objects) is dropped from the saved state -- every other key is the reference's own."""
    sys.modules["psgd"] = ref
    import importlib.util
    spec = importlib.util.spec_from_file_location("kwns4_reference", "/root/reference/wrapped_as_torch_optimizer_for_ddp.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    del sys.modules["psgd"]
    g0 = torch.Generator().manual_seed(8)
    p_ref = torch.nn.Parameter(torch.randn(*shape, generator=g0).to(ptype))
    opt = mod.KWNS4([p_ref], preconditioner_dtype=pdtype, **kw)
    case = {"name": name, "shape": list(shape), "pdtype": str(pdtype), "ptype": str(ptype), "kw": kw, "save_at": save_at, "steps": []}
    for s in range(steps):
        if s == save_at:
            sd = opt.state_dict()
            sd = {"state": {k: {kk: vv for kk, vv in st.items() if kk != "exprs"} for k, st in sd["state"].items()},
                  "param_groups": sd["param_groups"]}
            case["checkpoint"] = copy.deepcopy(sd)
            case["p_at_save"] = p_ref.detach().clone()
        grad = structured_grad(shape, torch.float32, 3100 + s).to(ptype)
        seed = 4141 + s
        p_ref.grad = grad.clone()
        torch.manual_seed(seed)
        opt.step()
        torch.manual_seed(seed)
        do_update = bool(torch.rand([]) < kw.get("preconditioner_update_probability", 1.0))
        st = opt.state[p_ref]
        gq = grad.squeeze().to(pdtype) if pdtype else grad.squeeze()
        # the draws of this step, re-drawn in the reference's order (the Q shapes are all draw_kron_noise needs)
        noise = orc.draw_kron_noise(gq, st["QL"][0]) if do_update else None
        case["steps"].append({"grad": grad, "seed": seed, "do_update": do_update, "noise": noise, "p": p_ref.detach().clone(),
                              "Q": [q.clone() for q in st["QL"][0]], "L": [l.clone() for l in st["QL"][1]],
                              "ema": None if st["ema"] is None else st["ema"].clone()})
    print(f"{name:28s} reference checkpoint at step {save_at}, keys {sorted(case['checkpoint']['state'][0].keys())}")
    torch.save(case, os.path.join(OUT, f"refckpt_{name}.pt"))


def round2_cases():
    """Fixtures added in round 2 (existing files are left untouched): update_precond_lra_newton, and PRO4P in bf16 with the
    procrustes_step3 round counts recorded (a parity test replays exactly the rounds the reference ran, psgd.py:444-449)."""
    f32, bf16 = torch.float32, torch.bfloat16
    lra_newton_case("r8_f32", 500, 8, f32)
    lra_newton_case("r16_bf16", 1000, 16, bf16)
    geom_cases("PRO4P", specs=[("whiten", (24, 40), bf16, 3), ("whiten", (136, 200), bf16, 2), ("newton", (24, 40), bf16, 2)], suffix="_bf16")
    kwns4_checkpoint_case("f32", (16, 24), f32, f32, lr_params=1e-2)
    kwns4_checkpoint_case("bf16_param", (16, 24), f32, bf16, lr_params=1e-2)    # bf16 parameter with an fp32 preconditioner


def rosenbrock(x):
    f = x.reshape(-1)
    x1, x2 = f[0::2], f[1::2]
    return torch.sum(100.0 * (x2 - x1 ** 2) ** 2 + (1.0 - x1) ** 2)


CLOSURE_CASES = [   # (class, kwargs): loss trajectories of the reference's closure-style optimizers on Rosenbrock (10x10 + 6 parameters), CPU fp32
    ("KronWhiten", dict(preconditioner_init_scale=None, lr_params=0.02, lr_preconditioner=0.3, momentum=0.9, dQ="Q0.5EQ1.5")),
    ("KronWhiten", dict(preconditioner_init_scale=1.0, lr_params=0.02, lr_preconditioner=0.3, momentum=0.9, whiten_grad=False, dQ="EQ",
                        update_preconditioner_first=False, preconditioner_update_probability=0.7)),
    ("KronWhiten", dict(preconditioner_init_scale=None, lr_params=0.02, lr_preconditioner=0.2, dQ="QUAD4P")),
    ("KronNewton", dict(preconditioner_init_scale=None, lr_params=0.3, lr_preconditioner=0.3, grad_clip_max_norm=1.0, dQ="Q0.5EQ1.5")),
    ("KronNewton", dict(preconditioner_init_scale=0.1, lr_params=0.3, lr_preconditioner=0.3, grad_clip_max_norm=1.0, momentum=0.5, dQ="QEP",
                        exact_hessian_vector_product=False, preconditioner_update_probability=0.8)),
    ("LRAWhiten", dict(rank_of_approximation=5, preconditioner_init_scale=None, lr_params=0.02, lr_preconditioner=0.3, momentum=0.9)),
    ("LRANewton", dict(rank_of_approximation=5, preconditioner_init_scale=0.1, lr_params=0.3, lr_preconditioner=0.3, grad_clip_max_norm=1.0)),
]


def closure_cases(steps=40):
    """The reference's KronWhiten / KronNewton / LRAWhiten / LRANewton (psgd.py:516-654, 832-978, 1075-1330), unmodified, on CPU."""
    out = []
    for name, kw in CLOSURE_CASES:
        torch.manual_seed(2024)
        xs = [torch.zeros(10, 10, requires_grad=True), torch.full((6,), 0.5, requires_grad=True)]
        opt = getattr(ref, name)(xs, **kw)
        losses = [float(opt.step(lambda: rosenbrock(xs[0]) + rosenbrock(xs[1]))) for _ in range(steps)]
        print(f"closure {name:10s} {kw.get('dQ', 'LRA'):10s}: loss {losses[0]:.4e} -> {losses[-1]:.4e}")
        out.append({"cls": name, "kw": kw, "losses": losses, "x": [x.detach().clone() for x in xs]})
    torch.save(out, os.path.join(OUT, "closures.pt"))


if __name__ == "__main__":
    torch.set_num_threads(1)  # deterministic reduction order
    if "--round2" in sys.argv:
        round2_cases()
        sys.exit(0)
    f32, bf16 = torch.float32, torch.bfloat16
    kron_case("dd_f32", (24, 40), f32)
    kron_case("dd_bf16", (24, 40), bf16)
    kron_case("dd_f32_balance", (16, 16), f32, steps=2, want_balance_step=1)
    kron_case("dense_diag_f32", (8, 300), f32)
    kron_case("diag_dense_f32", (300, 8), f32)
    kron_case("dense_diag_bf16", (8, 300), bf16)
    kron_case("vec_f32", (50,), f32)
    kron_case("vec_bf16", (64,), bf16)
    kron_case("big_dd_f32", (136, 200), f32, steps=2)
    kron_case("big_dd_bf16", (136, 200), bf16, steps=2)
    kron_case("order3_f32", (2, 3, 4), f32)
    lra_case("r4_f32", 300, 4, f32)
    lra_case("r4_bf16", 300, 4, bf16)
    lra_case("r16_f32", 1000, 16, f32)
    kwns4_case("f32", (16, 24), f32, steps=5, lr_params=1e-2)
    kwns4_case("bf16", (16, 24), bf16, steps=5, lr_params=1e-2)
    kwns4_case("bf16_squeeze", (1, 12, 1, 20), bf16, steps=4, lr_params=1e-2, momentum=0.0, whiten_grad=True,
               weight_decay=0.0, update_preconditioner_first=False)
    for dq in DQ_FUNCS:
        geom_cases(dq)
    closure_cases()
    round2_cases()
