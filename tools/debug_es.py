import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import tramp_oracle as orc
from tramp_b200.priors import GaussBernoulliPrior
from tramp_b200.likelihoods import GaussianLikelihood
from tramp_b200.channels import LinearChannel
from tramp_b200.variables import SISOVariable as V
from tramp_b200.algos import ExpectationPropagation
rng = np.random.RandomState(9)
B, N, M = 4, 120, 60
W = rng.randn(B, M, N) / np.sqrt(N)
x = rng.randn(B, N) * (rng.rand(B, N) < 0.1)
y = np.einsum("bmn,bn->bm", W, x) + 0.1 * rng.randn(B, M)
model = (GaussBernoulliPrior(size=N, rho=0.1, batch=B) @ V("x") @ LinearChannel(W) @ V("z")
         @ GaussianLikelihood(y=y, var=1e-2)).to_model()
ep = ExpectationPropagation(model)
ep.iterate(max_iter=200)
got = ep.get_variables_data()
print("n_iter", ep.n_iter_per_instance, "flags", ep.flags)
for b in range(B):
    ref = orc.ep_glm(dict(kind="gauss_bernoulli", rho=0.1), W[b], dict(kind="gaussian", var=1e-2, y=y[b]), 200,
                     early_stopping=dict(tol=1e-6), record_r=True)
    errs = [np.abs(got["x"]["r"][b] - r).max() for r in ref["traj"]["r_x"]]
    print(b, "ref n_iter", ref["n_iter"], "closest oracle iteration", int(np.argmin(errs)), "err", min(errs), "err vs final", errs[-1])
    print("   tol record", ep.records["tol"][:, b][-5:])
