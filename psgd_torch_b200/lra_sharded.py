"""Row-sharded LRA preconditioner: Q = (I + U V^T) diag(d) with the rows of U, V, d spread over the GPUs of one box.

The reference keeps the whole preconditioner on every device (psgd.py:1075-1190 builds U, V of n x r for n = all parameters: 67 GB in bf16
for the Llama-3-8B embedding at r = 32).  Every cross-row quantity of the update and of the apply is a short sum or a maximum
(psgd.py:1006-1052: U^T U, V^T V, V^T U, the r-vectors V^T x / U^T x, two maxima; psgd.py:1055-1063: two r-vectors), so a rank that
owns rows [lo, hi) runs the engine's sweeps on its rows and the ranks all-reduce ~13 KB per update and ~0.3 KB per apply over NCCL /
NVLink in between (psgd_lra_update_staged / psgd_lra_precond_grad_staged, include/psgd_b200.h).  The Lipschitz constants Lu, Lv, Ld
are replicated: every rank computes them from the same all-reduced sums.  Results equal the single-device preconditioner up to the
order of the fp32 partial sums.
"""
import ctypes as C

import torch

from . import _lib
from .psgd import _lra_desc

ST_SWEEP1, ST_SWEEP2, ST_FINISH = 1, 2, 4


class ShardedLRA:
    def __init__(self, UVd, Luvd, group=None):
        """UVd = [U, V, d]: this rank's rows; Luvd = [Lu, Lv, Ld] fp32 scalars (replicated); group: torch.distributed process group
        (None: the default group if torch.distributed is initialised, else a single shard)."""
        self.UVd, self.Luvd, self.group = UVd, Luvd, group
        self.dev = UVd[0].device
        self._lib = _lib.load_library()
        self._h = _lib.handle_for(self.dev)
        self._desc = _lra_desc(UVd, Luvd)
        nbytes = self._lib.psgd_lra_workspace_bytes(self._h, C.byref(self._desc))
        self._ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.dev)   # own scratch: lives across the stages
        offs, cnts = (C.c_size_t * 4)(), (C.c_size_t * 4)()
        _lib.check(self._h, self._lib.psgd_lra_workspace_offsets(self._h, C.byref(self._desc), offs, cnts), "psgd_lra_workspace_offsets")
        self._views = [self._ws[offs[i]:offs[i] + 4 * cnts[i]].view(torch.float32) for i in range(4)]
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=self.dev)

    # ---- the cross-row quantities (fp32 views into the workspace) ----
    @property
    def sums(self):
        return self._views[0]      # SUM between update stages 1 and 2

    @property
    def maxima(self):
        return self._views[1]      # MAX between update stages 2 and 4

    @property
    def proj1(self):
        return self._views[2]      # SUM after apply mode 0

    @property
    def proj2(self):
        return self._views[3]      # SUM after apply mode 1

    # ---- single stages (the collective methods below, and tests that emulate the all-reduce on one GPU, drive these) ----
    def update_stage(self, stages, gh, v, lr, betaL, damping, whiten, update_U):
        rc = self._lib.psgd_lra_update_staged(self._h, C.byref(self._desc), _lib.ptr(gh), _lib.ptr(v), float(lr), float(betaL), float(damping),
                                              int(whiten), int(update_U), int(stages), _lib.ptr(self._ws), self._ws.numel(),
                                              _lib.stream_ptr(self.dev))
        _lib.check(self._h, rc, "psgd_lra_update_staged")

    def apply_stage(self, modes, g, out):
        rc = self._lib.psgd_lra_precond_grad_staged(self._h, C.byref(self._desc), _lib.ptr(g), _lib.ptr(out), _lib.ptr(self._sumsq), int(modes),
                                                    _lib.ptr(self._ws), self._ws.numel(), _lib.stream_ptr(self.dev))
        _lib.check(self._h, rc, "psgd_lra_precond_grad_staged")

    # ---- collectives ----
    def _distributed(self):
        return torch.distributed.is_available() and torch.distributed.is_initialized() and \
            torch.distributed.get_world_size(self.group) > 1

    def _all_reduce(self, t, op):
        if self._distributed():
            torch.distributed.all_reduce(t, op=op, group=self.group)

    def _agree(self, flag):
        """the CPU coin of psgd.py:1035 must be the same on every rank: rank 0's draw wins"""
        if not self._distributed():
            return flag
        t = torch.tensor([1.0 if flag else 0.0], device=self.dev)
        torch.distributed.broadcast(t, src=torch.distributed.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        return bool(t.item() > 0.5)

    def _update(self, gh, v, lr, betaL, damping, whiten, update_U):
        SUM, MAX = torch.distributed.ReduceOp.SUM, torch.distributed.ReduceOp.MAX
        gh, v = gh.contiguous(), v.contiguous()
        self.update_stage(ST_SWEEP1, gh, v, lr, betaL, damping, whiten, update_U)
        self._all_reduce(self.sums, SUM)
        self.update_stage(ST_SWEEP2, gh, v, lr, betaL, damping, whiten, update_U)
        self._all_reduce(self.maxima, MAX)
        self.update_stage(ST_FINISH, gh, v, lr, betaL, damping, whiten, update_U)

    def update_precond_lra(self, v, h, lr=0.1, betaL=0.9, update_U=None):
        """psgd.py:994-1052 on this rank's rows of (v, h); collective over the group."""
        if update_U is None:
            update_U = self._agree(bool(torch.rand([]) < 0.5))
        self._update(h, v, lr, betaL, 0.0, False, update_U)

    def update_precond_lra_whiten(self, g, lr=0.1, betaL=0.9, damping=1e-9, noise=None):
        """psgd.py:1066-1072 on this rank's rows of g; collective over the group."""
        if noise is None:
            noise = {"v": torch.randn_like(g), "update_U": self._agree(bool(torch.rand([]) < 0.5))}
        self._update(g, noise["v"], lr, betaL, damping, True, noise["update_U"])

    def precond_grad_lra(self, g, sumsq_out=None):
        """psgd.py:1055-1063: this rank's rows of d (I + V U^T)(I + U V^T) d g; `sumsq_out` receives the global sum of squares."""
        SUM = torch.distributed.ReduceOp.SUM
        g = g.contiguous()
        out = torch.empty_like(g)
        self.apply_stage(1, g, out)
        self._all_reduce(self.proj1, SUM)
        self.apply_stage(2, g, out)
        self._all_reduce(self.proj2, SUM)
        self.apply_stage(4, g, out)
        if sumsq_out is not None:
            self._all_reduce(self._sumsq, SUM)
            sumsq_out.copy_(self._sumsq.reshape(sumsq_out.shape))
        return out
