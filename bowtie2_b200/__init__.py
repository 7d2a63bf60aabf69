"""bowtie2_b200 -- B200 (sm_100a) native implementation of bowtie2's alignment hot path.

The product is ``libbt2g.so`` (hand-written CUDA behind the C ABI in ``include/bt2g.h``).
This package is the thin host-side mirror used by the tests, the benchmark and the
multi-GPU plumbing: a ctypes binding (``lib``), the policy arithmetic the caller of the hot
path needs (``policy``), and synthetic-workload generators (``synth``).  There is no CPU
fallback: if the CUDA library is missing or no GPU is present, construction raises.
"""
from .lib import Bt2Gpu, Bt2GpuError, load_library, library_path  # noqa: F401
