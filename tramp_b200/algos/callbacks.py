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
        pass

    def device_replayable(self, algo):
        return True


class JoinCallback(Callback):
    """reference callbacks.py:22-32."""

    def __init__(self, callbacks):
        self.callbacks = callbacks
        self.repr_init(pad="\t")

    def __call__(self, algo, i, max_iter):
        stops = [callback(algo, i, max_iter) for callback in self.callbacks]
        return any(stops)

    def device_replayable(self, algo):
        return all(isinstance(c, Callback) and c.device_replayable(algo) for c in self.callbacks)

    def device_config(self, cfg):
        for c in self.callbacks:
            c.device_config(cfg)

    def replay(self, algo, i, max_iter, rec):
        for c in self.callbacks:
            c.replay(algo, i, max_iter, rec)


class LogProgress(Callback):
    def __init__(self, ids="all", every=1):
        self.ids = ids
        self.every = every
        self.repr_init()

    def __call__(self, algo, i, max_iter):
        if (i % self.every == 0):
            variables_data = algo.get_variables_data(self.ids)
            logger.info(f"iteration={i+1}/{max_iter}")
            for variable_id, data in variables_data.items():
                logger.info(f"id={variable_id} v={np.mean(data['v']):.3f}")


class TrackMessages(Callback):
    """reference callbacks.py:49-60."""

    def __init__(self, keys=["a", "n_iter", "direction"]):
        self.keys = keys
        self.records = []

    def __call__(self, algo, i, max_iter):
        if (i == 0):
            self.records = []
        self.records += algo.get_edges_data(self.keys)

    def get_dataframe(self):
        return pd.DataFrame(self.records)


class TrackObjective(Callback):
    """reference callbacks.py:63-85."""

    def __init__(self):
        self.edge_records = []
        self.node_records = []
        self.model_records = []

    def __call__(self, algo, i, max_iter):
        if (i == 0):
            self.records = []
        algo.update_objective()
        self.model_records.append(dict(A=algo.A_model, n_iter=algo.n_iter))
        self.edge_records += algo.get_edges_data(["A", "n_iter", "direction"])
        self.node_records += algo.get_nodes_data(["A", "n_iter"])

    def get_dataframe(self):
        return (pd.DataFrame(self.edge_records), pd.DataFrame(self.node_records),
                pd.DataFrame(self.model_records))


class TrackOverlaps(Callback):
    """reference callbacks.py:165-192: m = <r, x>/N, q = <r, r>/N, Q = <x, x>/N."""

    def __init__(self, true_values, ids="all", every=1, verbose=False):
        self.ids = ids
        self.every = every
        self.repr_init()
        self.X_true = true_values
        self.records = []
        self.verbose = verbose

    def __call__(self, algo, i, max_iter):
        if (i == 0):
            self.records = []
        if (i % self.every == 0):
            variables_data = algo.get_variables_data(self.ids)
            for variable_id, data in variables_data.items():
                x = np.asarray(self.X_true[variable_id])
                r = np.asarray(data["r"])
                n = x.shape[-1] if x.ndim > 1 else x.shape[0]
                record = dict(id=variable_id, m=(r * x).sum(-1) / n, q=(r * r).sum(-1) / n,
                              Q=(x * x).sum(-1) / n, iter=i)
                self.records.append(record)
                if self.verbose:
                    print(record)

    def get_dataframe(self):
        return pd.DataFrame(self.records)


def _squeeze(algo, row):
    """(B,) device record -> float for an un-batched model, array otherwise."""
    return float(row[0]) if not algo.batched else np.array(row)


class TrackEvolution(Callback):
    """reference callbacks.py:88-108."""

    def __init__(self, ids="all", every=1, verbose=False):
        self.ids = ids
        self.every = every
        self.repr_init()
        self.records = []
        self.verbose = verbose

    def __call__(self, algo, i, max_iter):
        if (i == 0):
            self.records = []
        if (i % self.every == 0):
            variables_data = algo.get_variables_data(self.ids)
            for variable_id, data in variables_data.items():
                record = dict(id=variable_id, v=data["v"], iter=i)
                self.records.append(record)
                if self.verbose:
                    print(record)

    def device_replayable(self, algo):
        return True

    def replay(self, algo, i, max_iter, rec):
        if (i == 0):
            self.records = []
        if (i % self.every == 0):
            for variable_id in algo.variable_ids:
                if self.ids == "all" or variable_id in self.ids:
                    key = "vx" if variable_id == algo.x_id else "vz"
                    self.records.append(dict(id=variable_id, v=_squeeze(algo, rec[key][i]), iter=i))

    def get_dataframe(self):
        return pd.DataFrame(self.records)


class TrackEstimate(Callback):
    """reference callbacks.py:111-128 (needs r every iteration: slow path)."""

    def __init__(self, ids="all", every=1):
        self.ids = ids
        self.every = every
        self.repr_init()
        self.records = []

    def __call__(self, algo, i, max_iter):
        if (i == 0):
            self.records = []
        if (i % self.every == 0):
            variables_data = algo.get_variables_data(self.ids)
            for variable_id, data in variables_data.items():
                self.records.append(dict(id=variable_id, r=data["r"], iter=i))

    def get_dataframe(self):
        return pd.DataFrame(self.records)


class TrackErrors(Callback):
    """reference callbacks.py:131-162."""

    def __init__(self, true_values, metrics=["mse"], every=1, verbose=False):
        self.ids = true_values.keys()
        self.metrics = metrics
        self.every = every
        self.repr_init()
        self.X_true = true_values
        self.errors = []
        self.verbose = verbose

    def __call__(self, algo, i, max_iter):
        if (i == 0):
            self.errors = []
        if (i % self.every == 0):
            variables_data = algo.get_variables_data(self.ids)
            X_pred = {variable_id: data["r"] for variable_id, data in variables_data.items()}
            errors = []
            for id in self.ids:
                error = dict(id=id, iter=i)
                for metric in self.metrics:
                    func = METRICS.get(metric)
                    error[metric] = func(X_pred[id], self.X_true[id])
                errors.append(error)
            if self.verbose:
                print(errors)
            self.errors += errors

    def device_replayable(self, algo):
        return (list(self.ids) == [algo.x_id]
                and all(m in ("mse", "sign_mse") for m in self.metrics))

    def device_config(self, cfg):
        cfg["x_true"] = self.X_true[list(self.ids)[0]]

    def replay(self, algo, i, max_iter, rec):
        if (i == 0):
            self.errors = []
        if (i % self.every == 0):
            error = dict(id=algo.x_id, iter=i)
            for metric in self.metrics:
                error[metric] = _squeeze(algo, rec["mse" if metric == "mse" else "smse"][i])
            self.errors.append(error)

    def get_dataframe(self):
        return pd.DataFrame(self.errors)


def norm(x):
    return np.sqrt(np.mean(x**2))


class EarlyStoppingEP(Callback):
    """reference callbacks.py:250-286: stop when the relative change of every
    tracked estimate, rms(r_new - r_old) / rms(r_new), is below `tol`; if after
    `wait_increase` iterations it exceeds `max_increase`, restore the previous
    messages and stop."""

    def __init__(self, ids="all", tol=1e-6, wait_increase=5, max_increase=0.2):
        self.ids = ids
        self.tol = tol
        self.wait_increase = wait_increase
        self.max_increase = max_increase
        self.repr_init()
        self.old_rs = None

    def __call__(self, algo, i, max_iter):
        if (i == 0):
            self.old_rs = None
        variables_data = algo.get_variables_data(self.ids)
        new_rs = [data["r"] for variable_id, data in variables_data.items()]
        if self.old_rs:
            tols = [norm(new_r - old_r) / norm(new_r) for old_r, new_r in zip(self.old_rs, new_rs)]
            if max(tols) < self.tol:
                logger.info(f"early stopping all tolerances (on r) are below tol={self.tol:.2e}")
                return True
            if i > self.wait_increase and max(tols) > self.max_increase:
                logger.info(f"increase={max(tols)} above max_increase={self.max_increase:.2e}")
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
    (`trb_se_run`), for EP it is an ordinary per-iteration callback."""

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
        # only the State-Evolution kernel implements the variance test
        return getattr(algo, "message_keys", None) == ["a"] and self._var_mask(algo) is not None

    def device_config(self, cfg):
        cfg["early_stopping"] = self
