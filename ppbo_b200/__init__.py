"""ppbo_b200: B200-native hot path of PPBO behind a C ABI (include/ppbo_b200.h)."""
import os as _os

# The pipeline drives two concurrent chains of small kernels from two host threads (iteration.run_iteration) on more than a dozen
# streams (torch's, the library's Cholesky streams per thread, side streams).  The driver maps streams onto
# CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8); streams that share a queue serialise in launch order: the weight-space
# Hessian GEMM sat 300 us behind the other chain's small GEMMs on an otherwise idle GPU (scripts/timeline.py --api).  Must be set
# before the CUDA context exists, hence at import time; an explicit setting of the user wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
