"""Metrics (reference tramp/algos/metrics.py:5-14, 29-40).  Host-side helpers;
inside the device-resident sweep the same quantities are produced by
k_x_update (tramp_b200/csrc/trb_sweep.cu)."""
import numpy as np


def mean_squared_error(x_true, x_pred):
    return np.mean((x_true - x_pred)**2, axis=-1) if np.ndim(x_true) > 1 else np.mean((x_true - x_pred)**2)


def sign_symmetric_mse(x_true, x_pred):
    "Mean squared error up to a global sign"
    ax = -1 if np.ndim(x_true) > 1 else None
    mse_pos = np.mean((x_true - x_pred) ** 2, axis=ax)
    mse_neg = np.mean((x_true + x_pred) ** 2, axis=ax)
    return np.minimum(mse_pos, mse_neg) if ax is not None else min(mse_pos, mse_neg)


def overlap(x_true, x_pred):
    return np.mean(x_true * x_pred, axis=-1) if np.ndim(x_true) > 1 else np.mean(x_true * x_pred)


METRICS = {
    "sign_mse": sign_symmetric_mse,
    "mse": mean_squared_error,
    "overlap": overlap,
}
