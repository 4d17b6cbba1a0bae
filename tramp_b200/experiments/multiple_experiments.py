"""Grid runner (reference tramp/experiments/multiple_experiments.py:8-72).

`run_experiments(run, **kwargs)` calls `run(**point)` on every point of the
Cartesian product of the keyword values and stacks the returned records in a
DataFrame, one column per keyword.  Each `run` is an EP or SE job that already
executes on the GPU; a grid of *State Evolution* problems does not need this
loop at all -- `run_state_evolution_grid` puts the whole grid in one launch.
"""
import itertools
import logging
import numpy as np
import pandas as pd

logger = logging.getLogger(__name__)


def log_on_progress(i, total):
    logger.info(f"experiment {i}/{total}")


def as_list(x):
    """A grid axis: lists stay, arrays become lists, anything else is one value."""
    if isinstance(x, list):
        return x
    if isinstance(x, np.ndarray):
        return list(x)
    return [x]


def get_experiments_from_kwargs(**kwargs):
    """[{key: value}] over the product of the axes, last key varying fastest."""
    names = list(kwargs)
    axes = [as_list(kwargs[name]) for name in names]
    return [dict(zip(names, point)) for point in itertools.product(*axes)]


def _records_of(run, experiment):
    results = run(**experiment)
    if isinstance(results, dict):
        results = [results]
    for result in results:
        result.update(experiment)
    return results


def run_experiments(run, on_progress=None, **kwargs):
    """A failing point is logged and skipped (reference :30-49)."""
    on_progress = on_progress or log_on_progress
    experiments = get_experiments_from_kwargs(**kwargs)
    records = []
    for idx, experiment in enumerate(experiments):
        try:
            records += _records_of(run, dict(experiment))
        except Exception as e:
            logger.error(f"Experiment {experiment} failed\n{e}")
        on_progress(idx + 1, len(experiments))
    return pd.DataFrame(records)


def simple_run_experiments(run, **kwargs):
    "Same as run_experiments but raises errors and has no `on_progress` callback (reference :52-67)"
    records = []
    for experiment in get_experiments_from_kwargs(**kwargs):
        records += _records_of(run, dict(experiment))
    return pd.DataFrame(records)


def save_experiments(run, csv_file, on_progress=None, **kwargs):
    df = run_experiments(run, on_progress, **kwargs)
    df.to_csv(csv_file, index=False)
