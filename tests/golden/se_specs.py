"""State-Evolution cases shared by make_golden_se.py (which needs the
reference) and the tests (which must not)."""
import numpy as np

# factor-level grids
SE_PRIOR_SPECS = [
    dict(kind="gauss_bernoulli", rho=0.1, mean=0, var=1),
    dict(kind="gauss_bernoulli", rho=0.5, mean=0.01, var=1),
    dict(kind="gauss_bernoulli", rho=0.3, mean=0.5, var=2.0),
    dict(kind="binary", p_pos=0.5),
    dict(kind="binary", p_pos=0.6),
    dict(kind="gaussian", mean=0.3, var=2.0),
]
SE_PRIOR_AX = np.array([0.0, 1e-3, 0.05, 0.3, 1.0, 2.5, 7.0, 20.0, 60.0])
SE_LIK_SPECS = [
    dict(kind="sgn"),
    dict(kind="abs"),
    dict(kind="gaussian", var=0.01),
]
# (az, tau_z) with az > 1/tau_z
SE_LIK_POINTS = np.array([(1.05, 1.0), (1.5, 1.0), (3.0, 1.0), (10.0, 1.0), (40.0, 1.0),
                          (12.0, 0.1), (25.0, 0.1), (0.7, 2.0), (2.2, 2.0)])
SE_ABS_POINTS = np.array([(1.5, 1.0), (3.0, 1.0), (12.0, 0.1)])   # dblquad is slow
SE_MP_ALPHAS = [0.3, 1.0, 2.5]
SE_MP_POINTS = np.array([(1.0, 0.0), (0.0, 1.0), (1e-3, 1.0), (1.0, 1.0), (0.3, 4.0), (10.0, 0.2),
                         (5.0, 1e4), (200.0, 3.0)])       # (az, ax)

# whole runs: name -> dict(prior, lik, channel, max_iter, damping, a_init, early)
# channel: dict(kind="marchenko", alpha) or dict(kind="spectrum", N, M, seed)
_ES = dict(tol=1e-6, min_variance=-1, wait_increase=5, max_increase=0.2)
SE_RUNS = {
    "cs_a05": dict(prior=dict(kind="gauss_bernoulli", rho=0.1, mean=0, var=1),
                   lik=dict(kind="gaussian", var=1e-2), channel=dict(kind="marchenko", alpha=0.5),
                   max_iter=200, early=_ES),
    "cs_a08_rho05_noiseless": dict(prior=dict(kind="gauss_bernoulli", rho=0.5, mean=0, var=1),
                                   lik=dict(kind="gaussian", var=1e-10),
                                   channel=dict(kind="marchenko", alpha=0.45),
                                   max_iter=200, early=_ES),
    "cs_damped": dict(prior=dict(kind="gauss_bernoulli", rho=0.25, mean=0, var=1),
                      lik=dict(kind="gaussian", var=1e-3), channel=dict(kind="marchenko", alpha=0.6),
                      max_iter=40, damping=0.5, early=None),
    "cs_informed": dict(prior=dict(kind="gauss_bernoulli", rho=0.5, mean=0, var=1),
                        lik=dict(kind="gaussian", var=1e-4), channel=dict(kind="marchenko", alpha=0.7),
                        max_iter=60, a_init=[("x", "bwd", 1e3)], early=_ES),
    "perceptron_bin": dict(prior=dict(kind="binary", p_pos=0.6), lik=dict(kind="sgn"),
                           channel=dict(kind="marchenko", alpha=1.0), max_iter=60, early=_ES),
    "perceptron_gb": dict(prior=dict(kind="gauss_bernoulli", rho=0.5, mean=0.2, var=1),
                          lik=dict(kind="sgn"), channel=dict(kind="marchenko", alpha=1.5),
                          max_iter=60, early=_ES),
    "perceptron_listdamp": dict(prior=dict(kind="binary", p_pos=0.7), lik=dict(kind="sgn"),
                                channel=dict(kind="marchenko", alpha=0.8), max_iter=25,
                                damping=[("x", "bwd", 0.3), ("z", "bwd", 0.1)], early=None),
    "phase_gb": dict(prior=dict(kind="gauss_bernoulli", rho=0.6, mean=0.01, var=1),
                     lik=dict(kind="abs"), channel=dict(kind="marchenko", alpha=0.9),
                     max_iter=6, a_init=[("x", "bwd", 0.1)], early=None),
    "phase_bin": dict(prior=dict(kind="binary", p_pos=0.6), lik=dict(kind="abs"),
                      channel=dict(kind="marchenko", alpha=1.5), max_iter=5, early=None),
    "student_cs": dict(prior=dict(kind="gauss_bernoulli", rho=0.1, mean=0, var=1),
                       lik=dict(kind="gaussian", var=1e-2),
                       channel=dict(kind="spectrum", N=120, M=60, seed=11),
                       max_iter=200, early=_ES),
    "student_perceptron": dict(prior=dict(kind="binary", p_pos=0.6), lik=dict(kind="sgn"),
                               channel=dict(kind="spectrum", N=80, M=120, seed=12),
                               max_iter=30, early=_ES),
}
# runs whose final state also gets an entropy (free-energy) golden value
SE_ENTROPY_RUNS = ["cs_a05", "cs_damped", "perceptron_gb", "perceptron_listdamp", "phase_bin",
                   "student_cs"]


def damping_dict(damping, x_id="x", z_id="z"):
    """reference message_passing.py:70-106 -> damping of e1, e3, e5, e7."""
    d = dict(e1=0., e3=0., e5=0., e7=0.)
    if not damping:
        return d
    if isinstance(damping, float):
        return dict(e1=damping, e3=damping, e5=damping, e7=damping)
    into = {(x_id, "fwd"): "e1", (x_id, "bwd"): "e7", (z_id, "fwd"): "e3", (z_id, "bwd"): "e5"}
    for id, direction, value in damping:
        d[into[(id, direction)]] = value
    return d


def a_init_dict(a_init, x_id="x", z_id="z"):
    """CustomInit(a_init=[(id, direction, a)]) -> initial a of e1..e8
    (reference initial_conditions.py:60-74 + message_passing.py:222-229)."""
    edges = dict(e1=(x_id, "fwd"), e2=(x_id, "fwd"), e3=(z_id, "fwd"), e4=(z_id, "fwd"),
                 e5=(z_id, "bwd"), e6=(z_id, "bwd"), e7=(x_id, "bwd"), e8=(x_id, "bwd"))
    table = {(id, d): a for id, d, a in (a_init or [])}
    return {name: float(table.get(key, 0.)) for name, key in edges.items()}


def spectrum_W(channel):
    """The matrix of a kind="spectrum" channel: Gaussian iid N(0, 1/N)."""
    rng = np.random.RandomState(channel["seed"])
    return rng.randn(channel["M"], channel["N"]) / np.sqrt(channel["N"])
