"""Elementwise exponential-family beliefs (reference tramp/beliefs/), evaluated
by the FP64 moment kernels of tramp_b200/csrc/trb_moments.cuh."""
from . import normal, binary, sparse, positive, truncated
