"""Small sweep that runs every kernel of the round-2c iteration (projection with the rescale inside,
chunked x / z updates over several chunks, one-CTA-per-instance fall-backs, early stopping) -- the
case compute-sanitizer is pointed at (tools/gpu_r2_ad.sh)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from tramp_b200 import ops, _lib
from tramp_b200.priors import GaussBernoulliPrior
from tramp_b200.likelihoods import GaussianLikelihood, SgnLikelihood
from tramp_b200.channels import LinearChannel
from tramp_b200.variables import SISOVariable as V
from tramp_b200.algos import ExpectationPropagation, TrackErrors, EarlyStoppingEP, JoinCallback

lib = _lib.load()
rng = np.random.RandomState(3)
B, N, M = 3, 2500, 1300                      # three chunks of x, two of z per instance
W = rng.randn(B, M, N) / np.sqrt(N)
x = rng.randn(B, N) * (rng.rand(B, N) < 0.1)
z = np.einsum("bmn,bn->bm", W, x)
U, s, Vt = np.linalg.svd(W, full_matrices=False)


def rows_padded(a):                          # [B, R, n] -> device [B, R, pad_ld(n)]
    ld = ops.pad_ld(a.shape[2])
    out = np.zeros(a.shape[:2] + (ld,))
    out[:, :, :a.shape[2]] = a
    return ops.to_dev(out)


lin = LinearChannel.from_factors(rows_padded(U.transpose(0, 2, 1)), ops.to_dev(s), rows_padded(Vt),
                                 Nx=M, Nz=N, rank=M)
out = {}
for name, lik in (("gaussian", GaussianLikelihood(y=z + 0.1 * rng.randn(B, M), var=1e-2)),
                  ("sgn", SgnLikelihood(y=np.sign(z)))):
    for fused, mask in ((1, 3), (0, 0), (1, 0), (0, 3)):
        lib.trb_set_fused_rescale(fused)
        lib.trb_set_update_kernels(mask)
        model = (GaussBernoulliPrior(size=N, rho=0.1, batch=B) @ V("x") @ lin @ V("z") @ lik).to_model()
        ep = ExpectationPropagation(model)
        ep.schedule = "general"
        track = TrackErrors({"x": x})
        ep.iterate(max_iter=8, callback=JoinCallback([track, EarlyStoppingEP(tol=1e-3)]), damping=0.2)
        out[(name, fused, mask)] = ep.get_variables_data()["x"]["r"]
    ref = out[(name, 0, 0)]
    for key, r in out.items():
        if key[0] == name:
            assert np.allclose(r, ref, rtol=1e-9, atol=1e-12), key
lib.trb_set_fused_rescale(1)
lib.trb_set_update_kernels(-1)
torch.cuda.synchronize()
print("sanitize sweep case ok")
