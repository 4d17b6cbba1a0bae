"""smoke(): one small sparse-GLM EP sweep through the public API on the GPU,
checked against the CPU oracle (which is test infrastructure, imported here
only as the checker)."""


def run_smoke(np, verbose=False):
    from tramp_b200.priors import GaussBernoulliPrior
    from tramp_b200.channels import LinearChannel
    from tramp_b200.likelihoods import GaussianLikelihood
    from tramp_b200.variables import SISOVariable as V
    from tramp_b200.algos import ExpectationPropagation, TrackErrors
    from oracle import tramp_oracle as orc

    rng = np.random.RandomState(3)
    B, N, M, n_iter = 3, 256, 128, 20
    W = rng.randn(B, M, N) / np.sqrt(N)
    x = rng.randn(B, N) * (rng.rand(B, N) < 0.1)
    y = np.einsum("bmn,bn->bm", W, x) + 0.1 * rng.randn(B, M)
    model = (GaussBernoulliPrior(size=N, rho=0.1, batch=B) @ V("x") @ LinearChannel(W) @ V("z")
             @ GaussianLikelihood(y=y, var=1e-2)).to_model()
    ep = ExpectationPropagation(model)
    track = TrackErrors({"x": x})
    ep.iterate(max_iter=n_iter, callback=track, damping=0.2)
    got = ep.get_variables_data()
    worst = 0.0
    for b in range(B):
        ref = orc.ep_glm(dict(kind="gauss_bernoulli", rho=0.1), W[b],
                         dict(kind="gaussian", var=1e-2, y=y[b]), n_iter, damping=0.2, x_true=x[b])
        for name, a, r in (("r_x", got["x"]["r"][b], ref["r_x"]), ("r_z", got["z"]["r"][b], ref["r_z"])):
            err = np.max(np.abs(a - r)) / max(1e-300, np.max(np.abs(r)))
            worst = max(worst, err)
        worst = max(worst, abs(got["x"]["v"][b] - ref["v_x"]) / ref["v_x"])
        mse = np.array([e["mse"][b] for e in track.errors])
        worst = max(worst, np.max(np.abs(mse - np.array(ref["traj"]["mse_x"])) / np.array(ref["traj"]["mse_x"])))
    if verbose:
        print(f"smoke: B={B} N={N} M={M} iters={n_iter} max rel deviation vs oracle = {worst:.3e}")
    assert worst < 1e-9, f"EP sweep deviates from the oracle: {worst:.3e}"
    return worst
