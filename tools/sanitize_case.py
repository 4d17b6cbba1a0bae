import numpy as np, torch, json, sys
sys.path.insert(0, '.')
from tramp_b200 import ops, _lib
from tramp_b200.channels.linear_channel import thin_svd_device
lib = _lib.load()
rng = np.random.RandomState(0)
for mask in (1, 3, 0):
    lib.trb_jacobi_set_fused(mask)
    for shape, method in (((3, 96, 200), "jacobi"), ((2, 70, 75), "jacobi_direct"), ((1, 300, 900), "jacobi"), ((2, 40, 1000), "jacobi_direct")):
        W = rng.randn(*shape) / np.sqrt(shape[2])
        Ut, s, Vt = thin_svd_device(ops.to_dev(W), method)
        assert np.allclose(s.cpu().numpy(), np.linalg.svd(W, compute_uv=False), rtol=1e-10)
lib.trb_jacobi_set_fused(1)
# the device-resident adaptive schedule on a small batch
from tramp_b200.priors import GaussBernoulliPrior
from tramp_b200.likelihoods import SgnLikelihood
from tramp_b200.channels import LinearChannel
from tramp_b200.variables import SISOVariable as V
from tramp_b200.algos import ExpectationPropagation, PassCallback, EarlyStopping
B, N, M = 2, 40, 60
W = rng.randn(B, M, N) / np.sqrt(N)
x = rng.randn(B, N) * (rng.rand(B, N) < 0.3)
y = np.sign(np.einsum("bmn,bn->bm", W, x))
model = (GaussBernoulliPrior(size=N, rho=0.3, batch=B) @ V("x") @ LinearChannel(W) @ V("z") @ SgnLikelihood(y=y)).to_model()
ep = ExpectationPropagation(model)
ep.iterate(max_iter=4, callback=PassCallback(), damping="adaptive")
ep2 = ExpectationPropagation(model)
ep2.iterate(max_iter=30, callback=EarlyStopping(tol=1e-4), damping=0.3)
torch.cuda.synchronize()
print("sanitize case ok", ep.n_iter, ep2.n_iter)
