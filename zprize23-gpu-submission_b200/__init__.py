"""B200-native PLONK prover hot path (drop-in for PNP-team/ZPrize23-gpu-submission's `gen_proof`).

Only what the path needs lives here: `csrc/` (hand-written sm_100a CUDA + the C-ABI of include/zprize_b200.h),
`build.py` (in-tree nvcc build) and `plonk.py` (ctypes mirror of the reference's Rust FFI structs and call).
The directory name contains a hyphen, so import it through `importlib` (see tests/conftest.py) or add this
directory to `sys.path` and `import plonk`.
"""
from .plonk import *  # noqa: F401,F403
from . import build as _build  # noqa: F401
