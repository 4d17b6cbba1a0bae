"""Callbacks (reference tramp/algos/callbacks.py).

A callback is `callback(algo, i, max_iter) -> truthy to stop`, exactly as in
the reference.  Callbacks that only need what the device-resident sweep
records itself (per-iteration MSE, variances, the EarlyStoppingEP tolerance)
declare `device_replayable()`; `iterate()` then runs the whole sweep without a
host round trip and feeds them the recorded trajectory afterwards.  Any other
callback is honoured by synchronising after every iteration (slow path).
"""
import logging
import numpy as np
import pandas as pd

from .metrics import METRICS
from ..base import ReprMixin

logger = logging.getLogger(__name__)


class Callback(ReprMixin):
    def device_replayable(self, algo):
        return False

    def device_config(self, cfg):
        """Add what this callback needs to the device sweep configuration."""

    def replay(self, algo, i, max_iter, rec):
        """Consume iteration i of a recorded trajectory; rec[key] is (n_iter, B)."""


class PassCallback(Callback):
    def __init__(self):
        self.repr_init()

    def __call__(self, algo, i, max_iter):
        return None

    def device_replayable(self, algo):
        return True


class JoinCallback(Callback):
    """Several callbacks as one; stops as soon as any of them asks to (reference
    callbacks.py:22-32).  Every member sees every iteration, also after another one voted to stop."""

    def __init__(self, callbacks):
        self.callbacks = callbacks
        self.repr_init(pad="\t")

    def __call__(self, algo, i, max_iter):
        votes = [member(algo, i, max_iter) for member in self.callbacks]
        return any(votes)

    def device_replayable(self, algo):
        return all(isinstance(c, Callback) and c.device_replayable(algo) for c in self.callbacks)

    def device_config(self, cfg):
        for c in self.callbacks:
            c.device_config(cfg)

    def replay(self, algo, i, max_iter, rec):
        for c in self.callbacks:
            c.replay(algo, i, max_iter, rec)


class _Recorder(Callback):
    """What the tracking callbacks share: a list that restarts at iteration 0 and receives, every
    `every` iterations, the rows `_rows(algo, i)` built from the live state -- or, when the sweep
    ran on the device without host round trips, the rows `_rows_recorded(algo, i, rec)` built
    from its recorded trajectory.  `store` names the attribute holding the list (the reference's
    callbacks call it `records` or `errors`); `get_dataframe()` turns it into a DataFrame."""
    store = "records"
    every = 1
    verbose = False

    def _visit(self, i, rows_of):
        if i == 0:
            setattr(self, self.store, [])
        if i % self.every:
            return
        rows = rows_of()
        if self.verbose:
            for row in rows:
                print(row)
        getattr(self, self.store).extend(rows)

    def __call__(self, algo, i, max_iter):
        self._visit(i, lambda: self._rows(algo, i))

    def replay(self, algo, i, max_iter, rec):
        self._visit(i, lambda: self._rows_recorded(algo, i, rec))

    def get_dataframe(self):
        return pd.DataFrame(getattr(self, self.store))


class LogProgress(Callback):
    """Logs the mean posterior variance of the tracked variables (reference callbacks.py:35-46)."""

    def __init__(self, ids="all", every=1):
        self.ids = ids
        self.every = every
        self.repr_init()

    def __call__(self, algo, i, max_iter):
        if i % self.every:
            return
        logger.info(f"iteration={i+1}/{max_iter}")
        for variable_id, data in algo.get_variables_data(self.ids).items():
            logger.info(f"id={variable_id} v={np.mean(data['v']):.3f}")


class TrackMessages(_Recorder):
    """One row per edge and iteration with the requested message keys (reference callbacks.py:49-60)."""

    def __init__(self, keys=["a", "n_iter", "direction"]):
        self.keys = keys
        self.records = []

    def _rows(self, algo, i):
        return algo.get_edges_data(self.keys)


class TrackObjective(Callback):
    """Per-iteration log-partitions of the edges, the nodes and the model (reference
    callbacks.py:63-85); the three lists accumulate over successive `iterate` calls, as there."""

    def __init__(self):
        self.edge_records = []
        self.node_records = []
        self.model_records = []

    def __call__(self, algo, i, max_iter):
        algo.update_objective()
        self.model_records.append(dict(A=algo.A_model, n_iter=algo.n_iter))
        self.edge_records.extend(algo.get_edges_data(["A", "n_iter", "direction"]))
        self.node_records.extend(algo.get_nodes_data(["A", "n_iter"]))

    def get_dataframe(self):
        return tuple(pd.DataFrame(rows) for rows in (self.edge_records, self.node_records, self.model_records))


class TrackOverlaps(_Recorder):
    """reference callbacks.py:165-192: m = <r, x>/N, q = <r, r>/N, Q = <x, x>/N."""

    def __init__(self, true_values, ids="all", every=1, verbose=False):
        self.ids = ids
        self.every = every
        self.repr_init()
        self.X_true = true_values
        self.records = []
        self.verbose = verbose

    def _rows(self, algo, i):
        rows = []
        for variable_id, data in algo.get_variables_data(self.ids).items():
            x, r = np.asarray(self.X_true[variable_id]), np.asarray(data["r"])
            n = x.shape[-1]
            rows.append(dict(id=variable_id, m=(r * x).sum(-1) / n, q=(r * r).sum(-1) / n,
                             Q=(x * x).sum(-1) / n, iter=i))
        return rows


def _squeeze(algo, row):
    """(B,) device record -> float for an un-batched model, array otherwise."""
    return float(row[0]) if not algo.batched else np.array(row)


class TrackEvolution(_Recorder):
    """Posterior variance of the tracked variables per iteration (reference callbacks.py:88-108)."""

    def __init__(self, ids="all", every=1, verbose=False):
        self.ids = ids
        self.every = every
        self.repr_init()
        self.records = []
        self.verbose = verbose

    def _rows(self, algo, i):
        return [dict(id=variable_id, v=data["v"], iter=i)
                for variable_id, data in algo.get_variables_data(self.ids).items()]

    def device_replayable(self, algo):
        return True

    def _rows_recorded(self, algo, i, rec):
        tracked = [v for v in algo.variable_ids if self.ids == "all" or v in self.ids]
        return [dict(id=v, v=_squeeze(algo, rec["vx" if v == algo.x_id else "vz"][i]), iter=i) for v in tracked]


class TrackEstimate(_Recorder):
    """Posterior mean of the tracked variables per iteration (reference callbacks.py:111-128);
    needs r on the host every iteration, hence the one-launch-per-iteration path."""

    def __init__(self, ids="all", every=1):
        self.ids = ids
        self.every = every
        self.repr_init()
        self.records = []

    def _rows(self, algo, i):
        return [dict(id=variable_id, r=data["r"], iter=i)
                for variable_id, data in algo.get_variables_data(self.ids).items()]


class TrackErrors(_Recorder):
    """Error metrics of the estimates against the true signals (reference callbacks.py:131-162;
    like there, the metric is called as metric(estimate, truth))."""
    store = "errors"

    def __init__(self, true_values, metrics=["mse"], every=1, verbose=False):
        self.ids = true_values.keys()
        self.metrics = metrics
        self.every = every
        self.repr_init()
        self.X_true = true_values
        self.errors = []
        self.verbose = verbose

    def _rows(self, algo, i):
        estimates = algo.get_variables_data(self.ids)
        return [dict(id=vid, iter=i, **{m: METRICS.get(m)(estimates[vid]["r"], self.X_true[vid]) for m in self.metrics})
                for vid in self.ids]

    def _visit(self, i, rows_of):
        # the reference prints the whole list of an iteration at once
        verbose, self.verbose = self.verbose, False
        try:
            before = 0 if i == 0 else len(self.errors)
            super()._visit(i, rows_of)
        finally:
            self.verbose = verbose
        if verbose and i % self.every == 0:
            print(self.errors[before:])

    def device_replayable(self, algo):
        return (list(self.ids) == [algo.x_id]
                and all(m in ("mse", "sign_mse") for m in self.metrics))

    def device_config(self, cfg):
        cfg["x_true"] = self.X_true[list(self.ids)[0]]

    def _rows_recorded(self, algo, i, rec):
        return [dict(id=algo.x_id, iter=i,
                     **{m: _squeeze(algo, rec["mse" if m == "mse" else "smse"][i]) for m in self.metrics})]


def norm(x):
    return np.sqrt(np.mean(x**2))


def _rms(x):
    """Root mean square over the components of every instance (last axis)."""
    return np.sqrt(np.mean(np.square(x), axis=-1))


class EarlyStoppingEP(Callback):
    """reference callbacks.py:250-286: stop when the relative change of every
    tracked estimate, rms(r_new - r_old) / rms(r_new), is below `tol`; if after
    `wait_increase` iterations it exceeds `max_increase`, restore the previous
    messages and stop.

    For a batched model the tolerance is taken per instance, like in the device sweep.  On the
    device-replayable path every instance stops (or is rolled back) on its own; as a plain
    per-iteration callback it can only stop the whole `iterate` call, which it does once EVERY
    instance is below `tol` -- or as soon as ONE diverges, rolling all of them back."""

    def __init__(self, ids="all", tol=1e-6, wait_increase=5, max_increase=0.2):
        self.ids = ids
        self.tol = tol
        self.wait_increase = wait_increase
        self.max_increase = max_increase
        self.repr_init()
        self.old_rs = None

    def __call__(self, algo, i, max_iter):
        if i == 0:
            self.old_rs = None
        new_rs = [np.asarray(data["r"]) for data in algo.get_variables_data(self.ids).values()]
        if self.old_rs:
            # worst tracked variable of the worst instance
            change = max(float(np.max(_rms(new - old) / _rms(new))) for old, new in zip(self.old_rs, new_rs))
            if change < self.tol:
                logger.info(f"early stopping all tolerances (on r) are below tol={self.tol:.2e}")
                return True
            if i > self.wait_increase and change > self.max_increase:
                logger.info(f"increase={change} above max_increase={self.max_increase:.2e}")
                logger.info("restoring old message dag")
                algo.reset_message_dag(self.old_message_dag)
                return True
        self.old_rs = new_rs
        self.old_message_dag = algo.snapshot()

    def _var_mask(self, algo):
        if self.ids == "all":
            return 3
        mask = 0
        for id in self.ids:
            if id == algo.x_id:
                mask |= 1
            elif id == algo.z_id:
                mask |= 2
            else:
                return None
        return mask or None

    def device_replayable(self, algo):
        return self._var_mask(algo) is not None

    def device_config(self, cfg):
        cfg["early_stopping"] = self


class EarlyStopping(Callback):
    """reference callbacks.py:195-243: stop when every tracked variance moved by
    less than `tol` (absolute), or dropped below `min_variance`; a NaN variance,
    or an increase above `max_increase` after `wait_increase` iterations,
    restores the previous messages and stops.  This is State Evolution's default
    stopper; inside `StateEvolution.iterate` the test runs in the kernel
    (`trb_se_run`), inside `ExpectationPropagation.iterate` in the sweep kernels
    (`trb_sweep.es_mode = 1`), per instance."""

    def __init__(self, ids="all", tol=1e-6, min_variance=-1, wait_increase=5, max_increase=0.2):
        self.ids = ids
        self.tol = tol
        self.min_variance = min_variance
        self.wait_increase = wait_increase
        self.max_increase = max_increase
        self.repr_init()
        self.old_vs = None

    def __call__(self, algo, i, max_iter):
        if (i == 0):
            self.old_vs = None
        variables_data = algo.get_variables_data(self.ids)
        # one float per variable (reference); a batched algorithm gives one array per
        # variable and the tests below then hold for every instance at once
        new_vs = [np.asarray(data["v"], dtype=float) for variable_id, data in variables_data.items()]
        if any(np.any(v < self.min_variance) for v in new_vs):
            logger.info(f"early stopping min variance {min(float(np.min(v)) for v in new_vs)}")
            return True
        if any(np.any(np.isnan(v)) for v in new_vs):
            logger.warning("early stopping nan values")
            logger.info("restoring old message dag")
            algo.reset_message_dag(self.old_message_dag)
            return True
        if self.old_vs:
            tol = max(float(np.max(np.abs(old_v - new_v))) for old_v, new_v in zip(self.old_vs, new_vs))
            if tol < self.tol:
                logger.info(f"early stopping all tolerances (on v) are below tol={self.tol:.2e}")
                return True
            increase = max(float(np.max(new_v - old_v)) for old_v, new_v in zip(self.old_vs, new_vs))
            if i > self.wait_increase and increase > self.max_increase:
                logger.info(f"increase={increase} above max_increase={self.max_increase:.2e}")
                logger.info("restoring old message dag")
                algo.reset_message_dag(self.old_message_dag)
                return True
        self.old_vs = new_vs
        self.old_message_dag = algo.snapshot()

    _var_mask = EarlyStoppingEP._var_mask

    def device_replayable(self, algo):
        # the State-Evolution kernel and (trb_sweep.es_mode = 1) the EP sweep kernels
        # implement the variance test
        return self._var_mask(algo) is not None

    def device_config(self, cfg):
        cfg["early_stopping"] = self
