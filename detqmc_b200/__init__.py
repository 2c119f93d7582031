"""detqmc_b200 -- B200-native (sm_100a) DQMC sweep hot path behind a C ABI.

The product is detqmc_b200/libdqmc_b200.so (CUDA kernels + the C ABI declared in
include/dqmc_gpu.h).  This package is the thin host-side mirror used by the tests and by bench.py:
ctypes bindings (`lib`) and Python classes that mirror the reference's model interface for the hot
path (`DetSDWBatch`, `ReplicaExchangeLadder`, the MPI-free parallel-tempering loop `DetQMCPT`).  There is no CPU fallback: importing the bindings
fails loudly when the shared library is missing, and creating a context fails without a CUDA device.
"""
from .lib import load_library, DqmcParams, ControlData, DqmcError          # noqa: F401
from .sdw import DetSDWBatch, ReplicaExchangeLadder                          # noqa: F401
from .hubbard import DetHubbardBatch                                         # noqa: F401
from .pt import DetQMCPT                                                     # noqa: F401
