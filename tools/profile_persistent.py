"""One persistent-sweep launch of BASELINE configs[1] (sign perceptron, N = 2000,
M = 4000, damping 0.5, 50 iterations) for ncu:

    ncu --set full --clock-control none --import-source on -k regex:k_sweep_persistent -c 1 \
        -o gpurun_out/r01e_persist python tools/profile_persistent.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tramp_b200 import _lib  # noqa: E402
from tramp_b200.priors import GaussianPrior  # noqa: E402
from tramp_b200.likelihoods import SgnLikelihood  # noqa: E402
from tramp_b200.channels import LinearChannel  # noqa: E402
from tramp_b200.variables import SISOVariable as V  # noqa: E402
from tramp_b200.algos import ExpectationPropagation, TrackErrors  # noqa: E402

N, M = 2000, 4000
rng = np.random.RandomState(42)
W = rng.randn(M, N) / np.sqrt(N)
x = rng.randn(N)
y = np.where(W @ x >= 0, 1.0, -1.0)
model = (GaussianPrior(size=N) @ V("x") @ LinearChannel(W) @ V("z") @ SgnLikelihood(y=y)).to_model()
_lib.load().trb_set_persistent_sweep(2)
ep = ExpectationPropagation(model)
track = TrackErrors({"x": x})
ep.iterate(max_iter=50, callback=track, damping=0.5)
print("final mse", track.errors[-1]["mse"])
