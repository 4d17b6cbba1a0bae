"""Critical measurement density from State Evolution (reference
tramp/experiments/critical_alpha.py:8-109).

`find_critical_alpha` bisects on alpha, one full SE run per step, exactly as
the reference does.  With `grid=G` every refinement step instead evaluates G
interior values of alpha in ONE batched SE launch (the interval shrinks by G+1
per launch instead of 2 per run), which is how the search maps onto the GPU.
"""
import logging
import numpy as np

from ..algos import StateEvolution, CustomInit

logger = logging.getLogger(__name__)


def binary_search(f, xmin, xmax, xtol):
    "Binary search on boolean f, assuming f(xmin)=0 and f(xmax)=1 (reference :8-30)"
    ymin, ymax = f(xmin), f(xmax)
    if not (ymin == 0 and ymax == 1):
        raise ValueError(f"Bad bounds: ymin={ymin} and ymax={ymax}")
    max_iter = int(np.log2((xmax - xmin) / xtol)) + 2
    for n_iter in range(1, max_iter + 1):
        xmid = (xmin + xmax) / 2
        ymid = f(xmid)
        xerr = xmax - xmin
        logger.info(f"binary search {n_iter}/{max_iter} xerr={xerr}")
        if (xerr < xtol):
            break
        if ymid == 0:
            xmin, ymin = xmid, ymid
        else:
            xmax, ymax = xmid, ymid
    assert ymin == 0 and ymax == 1
    assert (xerr < xtol)
    return dict(xmid=xmid, xmin=xmin, xmax=xmax, xerr=xerr, n_iter=n_iter)


def find_state_evolution_mse(id, a0, alpha, model_builder, **model_kwargs):
    """v of variable `id` at the SE fixed point reached from a(id -> prior) = a0
    (reference :33-57).  `alpha` may be an array: all values run in one launch."""
    initializer = CustomInit(a_init=[(id, "bwd", a0)])
    if np.ndim(alpha) == 0:
        se = StateEvolution(model_builder(alpha=alpha, **model_kwargs))
    else:
        se = StateEvolution([model_builder(alpha=float(al), **model_kwargs) for al in alpha])
    se.iterate(max_iter=200, initializer=initializer)
    return se.get_variable_data(id=id)["v"]


def grid_search(f_many, xmin, xmax, xtol, grid):
    """Like binary_search for a vectorised boolean f: `grid` interior points per
    step, the bracket moves to the first 0 -> 1 transition."""
    y_ends = f_many(np.array([xmin, xmax]))
    if not (y_ends[0] == 0 and y_ends[1] == 1):
        raise ValueError(f"Bad bounds: ymin={y_ends[0]} and ymax={y_ends[1]}")
    n_iter = 0
    while xmax - xmin >= xtol:
        n_iter += 1
        xs = np.linspace(xmin, xmax, grid + 2)[1:-1]
        ys = np.asarray(f_many(xs), dtype=bool)
        k = int(np.argmax(ys)) if ys.any() else grid       # first point where f is 1
        xmin, xmax = (xs[k - 1] if k > 0 else xmin), (xs[k] if k < grid else xmax)
        logger.info(f"grid search step {n_iter} xerr={xmax - xmin}")
    return dict(xmid=(xmin + xmax) / 2, xmin=xmin, xmax=xmax, xerr=xmax - xmin, n_iter=n_iter)


def find_critical_alpha(id, a0, mse_criterion, alpha_min, alpha_max, model_builder,
                        alpha_tol=1e-6, vtol=1e-3, grid=None, **model_kwargs):
    """Smallest alpha for which the mse criterion holds (reference :60-109).

    mse_criterion : "perfect" (v = 0 within vtol), "random" (v differs from
    tau_x by more than vtol) or a function v -> bool that is False below the
    critical alpha and True above.
    grid : None = the reference's bisection; int G = batched G-section search.
    """
    if mse_criterion == "perfect":
        def mse_criterion(v):
            return abs(v) < vtol
    elif mse_criterion == "random":
        # tau_x is assumed not to depend on alpha
        model = model_builder(alpha=0.5, **model_kwargs)
        model.init_second_moments()
        tau_x = model.get_second_moments()[id]

        def mse_criterion(v):
            return abs(v - tau_x) > vtol

    if grid:
        def f_many(alphas):
            vs = find_state_evolution_mse(id, a0, alphas, model_builder, **model_kwargs)
            return np.array([bool(mse_criterion(v)) for v in vs])
        return grid_search(f_many, alpha_min, alpha_max, alpha_tol, int(grid))["xmid"]

    def f(alpha):
        return mse_criterion(find_state_evolution_mse(id, a0, alpha, model_builder, **model_kwargs))
    return binary_search(f, alpha_min, alpha_max, alpha_tol)["xmid"]
