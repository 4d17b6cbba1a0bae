"""One batched State-Evolution launch (148 problems, one CTA per SM) for ncu:

    ncu --set full --clock-control none --import-source on -k regex:k_se_run -c 1 \
        -o gpurun_out/r01c_se python tools/profile_se.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tramp_b200.models import glm_state_evolution  # noqa: E402
from tramp_b200.algos import StateEvolution, PassCallback  # noqa: E402

models = [glm_state_evolution(alpha=float(a), prior_type="gauss_bernoulli", output_type="sgn",
                              prior_rho=0.5, prior_mean=0.2)
          for a in np.linspace(0.5, 3.0, 148)]
se = StateEvolution(models)
se.iterate(max_iter=20, callback=PassCallback())
print("v_x range", se.get_variable_data("x")["v"].min(), se.get_variable_data("x")["v"].max())
