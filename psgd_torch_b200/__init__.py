"""psgd_torch_b200 -- B200-native (sm_100a) engine for the PSGD preconditioner hot path of lixilinx/psgd_torch.

    from psgd_torch_b200 import psgd          # functional API with the reference's names (psgd.py)
    from psgd_torch_b200 import KWNS4         # torch.optim.Optimizer drop-in (wrapped_as_torch_optimizer_for_ddp.py)

Host code is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic runs in hand-written CUDA
behind the C-ABI of include/psgd_b200.h (libpsgd_b200.so, built in-tree by `python -m psgd_torch_b200.build`).
"""
from . import psgd  # noqa: F401
from ._lib import EngineError, load_library, launch_count, set_fp32_tensor_cores  # noqa: F401

try:  # wrappers are pure Python on top of .psgd
    from .kwns4 import KWNS4  # noqa: F401
    from .lra_optim import LRAWhitenOptimizer  # noqa: F401
    from .closure_optim import KronNewton, KronWhiten, LRANewton, LRAWhiten  # noqa: F401
except ImportError:  # pragma: no cover - during bring-up
    pass
