"""Drop-in for /root/reference/wrapped_as_torch_optimizer_for_dtensor.py: `from psgd_torch_b200.kwns4_dtensor import KWNS4`."""
from .kwns4 import KWNS4DTensor as KWNS4  # noqa: F401
