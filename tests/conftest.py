import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def relerr(a, b):
    """Normwise relative error ||a-b||_F / ||b||_F in fp64 (elementwise rel. error is meaningless near zeros)."""
    import torch

    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def load_golden(name):
    import torch

    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


# ---------------------------------------------------------------------------------------------------------------------
# measured-error log: every parity comparison of the GPU suite appends one line (case, quantity, measured error, tolerance and --
# for bf16 -- the reference arithmetic's own distance from an fp64 evaluation of the same step) to gpurun_out/parity_errors.log;
# the copy committed as profiles/r02_parity_errors.log is that file from a B200 run.
# ---------------------------------------------------------------------------------------------------------------------
PARITY_LOG = os.environ.get("PSGD_PARITY_LOG", os.path.join(ROOT, "gpurun_out", "parity_errors.log"))


def pytest_sessionstart(session):
    markexpr = getattr(session.config.option, "markexpr", "") or ""
    if "gpu" in markexpr and "not gpu" not in markexpr and not session.config.option.collectonly:
        try:
            import torch
            if not torch.cuda.is_available():
                return
            os.makedirs(os.path.dirname(PARITY_LOG), exist_ok=True)
            with open(PARITY_LOG, "w") as f:
                f.write("# case | quantity | measured normwise rel. error | tolerance | reference arithmetic's own error vs fp64 (bf16 cases)\n")
        except OSError:
            pass


def parity_log(case, what, err, tol=None, ref=None):
    try:
        with open(PARITY_LOG, "a") as f:
            f.write(f"{case} | {what} | {err:.3e} | {'' if tol is None else f'{tol:.1e}'} | {'' if ref is None else f'{ref:.3e}'}\n")
    except OSError:
        pass


def check(case, what, a, b, tol, yard=None, b_ref=None, slack=1.5, floor=2e-3):
    """assert relerr(a, b) < tol and log the measured error.  With `yard` (an fp64 evaluation of the same step) the engine's distance from
    it is logged next to the reference arithmetic's own (b_ref, default b) and must satisfy err(a, yard) <= slack * err(b_ref, yard) + floor."""
    e = relerr(a, b)
    parity_log(case, what, e, tol)
    assert e < tol, (case, what, e, tol)
    if yard is not None:
        ey, er = relerr(a, yard), relerr(b if b_ref is None else b_ref, yard)
        parity_log(case, what + " vs fp64", ey, slack * er + floor, er)
        assert ey <= slack * er + floor, (case, what, ey, er)
    return e
