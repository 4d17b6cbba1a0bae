"""Generate tests/golden/se.npz by running the UNMODIFIED reference's State
Evolution (/root/reference, sphinxteam/tramp) on the cases of se_specs.py.

Run in the build container only:    python tests/golden/make_golden_se.py

Same veneer as make_golden.py (tests/golden/_refshim.py); no reference source is
modified or copied.  The oracle (oracle/se_oracle.py) and the CUDA path
(tramp_b200/csrc/trb_se.cu) are both checked against the output.
"""
import os
import sys
import logging
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402

_refshim.install()
logging.disable(logging.CRITICAL)

from tramp.priors import GaussBernoulliPrior, BinaryPrior, GaussianPrior  # noqa: E402
from tramp.likelihoods import GaussianLikelihood, SgnLikelihood, AbsLikelihood  # noqa: E402
from tramp.channels import LinearChannel, MarchenkoPasturChannel  # noqa: E402
from tramp.variables import SISOVariable as V  # noqa: E402
from tramp.algos import (  # noqa: E402
    StateEvolution, EarlyStopping, TrackEvolution, JoinCallback, CustomInit, ConstantInit,
)
from se_specs import (  # noqa: E402
    SE_PRIOR_SPECS, SE_PRIOR_AX, SE_LIK_SPECS, SE_LIK_POINTS, SE_ABS_POINTS, SE_MP_ALPHAS,
    SE_MP_POINTS, SE_RUNS, SE_ENTROPY_RUNS, spectrum_W,
)


def _prior(spec):
    kw = {k: v for k, v in spec.items() if k != "kind"}
    return dict(gauss_bernoulli=GaussBernoulliPrior, binary=BinaryPrior,
                gaussian=GaussianPrior)[spec["kind"]](size=None, **kw)


def _lik(spec):
    kw = {k: v for k, v in spec.items() if k != "kind"}
    return dict(gaussian=GaussianLikelihood, sgn=SgnLikelihood,
                abs=AbsLikelihood)[spec["kind"]](y=None, **kw)


def _channel(spec):
    if spec["kind"] == "marchenko":
        return MarchenkoPasturChannel(alpha=spec["alpha"])
    return LinearChannel(spectrum_W(spec))


def gen_factors(out):
    for i, spec in enumerate(SE_PRIOR_SPECS):
        p = _prior(spec)
        out[f"prior{i}_tau"] = np.float64(p.second_moment())
        out[f"prior{i}_v"] = np.array([p.compute_forward_error(ax) for ax in SE_PRIOR_AX])
        out[f"prior{i}_anew"] = np.array([p.compute_forward_state_evolution(ax) for ax in SE_PRIOR_AX])
        out[f"prior{i}_A"] = np.array([p.compute_free_energy(ax) for ax in SE_PRIOR_AX])
    for i, spec in enumerate(SE_LIK_SPECS):
        lk = _lik(spec)
        pts = SE_ABS_POINTS if spec["kind"] == "abs" else SE_LIK_POINTS
        out[f"lik{i}_v"] = np.array([lk.compute_backward_error(az, tau) for az, tau in pts])
        out[f"lik{i}_anew"] = np.array([lk.compute_backward_state_evolution(az, tau) for az, tau in pts])
        out[f"lik{i}_A"] = np.array([lk.compute_free_energy(az, tau) for az, tau in pts])
    for i, alpha in enumerate(SE_MP_ALPHAS):
        ch = MarchenkoPasturChannel(alpha=alpha)
        out[f"mp{i}_mean_spectrum"] = np.float64(ch.ensemble.mean_spectrum)
        out[f"mp{i}_vx"] = np.array([ch.compute_forward_error(az, ax, 1.0) for az, ax in SE_MP_POINTS])
        out[f"mp{i}_vz"] = np.array([ch.compute_backward_error(az, ax, 1.0) for az, ax in SE_MP_POINTS])
        ok = [(az, ax) for az, ax in SE_MP_POINTS if az > 0 and ax > 0]
        out[f"mp{i}_A"] = np.array([ch.compute_free_energy(az, ax, 0.7) for az, ax in ok])


def gen_runs(out):
    for name, case in SE_RUNS.items():
        np.random.seed(0)
        prior, lin, lik = _prior(case["prior"]), _channel(case["channel"]), _lik(case["lik"])
        model = (prior @ V(id="x") @ lin @ V(id="z") @ lik).to_model()
        se = StateEvolution(model)
        evo = TrackEvolution()
        callbacks = [evo]
        if case.get("early"):
            callbacks.append(EarlyStopping(**case["early"]))
        init = CustomInit(a_init=case["a_init"]) if case.get("a_init") else ConstantInit(a=0, b=0)
        se.iterate(max_iter=case["max_iter"], callback=JoinCallback(callbacks), initializer=init,
                   damping=case.get("damping"))
        df = evo.get_dataframe()
        out[f"{name}_vx"] = df[df.id == "x"].v.values.astype(float)
        out[f"{name}_vz"] = df[df.id == "z"].v.values.astype(float)
        out[f"{name}_n_iter"] = np.int64(se.n_iter)
        data = se.get_variables_data()
        out[f"{name}_v_final"] = np.array([data["x"]["v"], data["z"]["v"]], dtype=float)
        out[f"{name}_tau"] = np.array([data["x"]["tau"], data["z"]["tau"]], dtype=float)
        # final a of the 8 edges, in e1..e8 order
        nodes = {n.id if hasattr(n, "id") else None: n for n in se.message_dag.nodes()}
        P, X, L, Z, K = model.forward_ordering
        order = [(P, X), (X, L), (L, Z), (Z, K), (K, Z), (Z, L), (L, X), (X, P)]
        out[f"{name}_a"] = np.array([se.message_dag[s][t]["a"] for s, t in order], dtype=float)
        if name in SE_ENTROPY_RUNS:
            out[f"{name}_entropy"] = np.float64(se.entropy())
        print(name, "n_iter", se.n_iter, "v", out[f"{name}_v_final"])


def gen_errors(out):
    """Gaussian prior + sgn: az = 1/tau_z at the first backward pass -> AssertionError
    (sgn_likelihood.py:80-81)."""
    model = (GaussianPrior(size=None) @ V(id="x") @ MarchenkoPasturChannel(alpha=2.0) @ V(id="z")
             @ SgnLikelihood(y=None)).to_model()
    try:
        StateEvolution(model).iterate(max_iter=5)
        out["gauss_sgn_raises"] = np.int64(0)
    except AssertionError:
        out["gauss_sgn_raises"] = np.int64(1)


if __name__ == "__main__":
    out = {}
    gen_factors(out)
    gen_runs(out)
    gen_errors(out)
    path = os.path.join(HERE, "se.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")
