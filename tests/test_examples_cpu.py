"""The example scripts (examples/, the reference's figure and table workflows) at toy
sizes through the emulated device: they must run end to end on the public API and
produce the reference's table layout."""
import os
import sys

import numpy as np
import pandas as pd
import pytest

from tests._emulated_device import emulated_device  # noqa: F401  (fixture)

EXAMPLES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples")


@pytest.fixture
def examples_on_path(monkeypatch):
    monkeypatch.syspath_prepend(EXAMPLES)
    yield
    for name in ("_common", "sparse_regression", "sparse_phase_retrieval", "glm_ep_vs_se"):
        sys.modules.pop(name, None)


def test_sparse_regression(emulated_device, examples_on_path, tmp_path):  # noqa: F811
    import sparse_regression
    df = sparse_regression.main(["--n", "80", "--instances", "3", "--ep-alphas", "3", "--se-alphas", "4",
                                 "--csv", str(tmp_path / "out.csv")])
    assert list(df.source) == ["EP"] * 3 + ["SE"] * 4 + ["BO"] * 4
    ep, se, bo = (df[df.source == s] for s in ("EP", "SE", "BO"))
    assert np.all(np.diff(se.v.values) < 0) and np.all(bo.v.values <= se.v.values * (1 + 1e-9))
    assert ep.v.iloc[-1] < 1e-3 < ep.v.iloc[0]          # alpha = 0.99 recovers, alpha = 0.03 does not
    assert emulated_device.calls["trb_se_run"] == 2      # each curve is one launch
    assert pd.read_csv(tmp_path / "out.csv").shape == df.shape


def test_sparse_phase_retrieval(emulated_device, examples_on_path, tmp_path):  # noqa: F811
    import sparse_phase_retrieval
    df = sparse_phase_retrieval.main(["--n", "60", "--instances", "2", "--ep-alphas", "2", "--se-alphas", "2",
                                      "--csv", str(tmp_path / "out.csv")])
    assert list(df.source) == ["EP"] * 2 + ["SE"] * 2 + ["BO"] * 2
    assert np.all(np.isfinite(df.v.values)) and np.all(df.v.values >= 0)


def test_glm_ep_vs_se(emulated_device, examples_on_path, tmp_path):  # noqa: F811
    import glm_ep_vs_se
    glm_ep_vs_se.main(["--n", "100", "--alphas", "3", "--dir", str(tmp_path)])
    cs = pd.read_csv(tmp_path / "compressed_sensing_ep_vs_se.csv")
    assert sorted(cs.columns) == ["N", "alpha", "ensemble_type", "n_iter", "prior_rho", "source", "v", "x_id"]
    assert len(cs) == 3 * 2 * 3 and set(cs.source) == {"SE", "EP", "mse"}
    pc = pd.read_csv(tmp_path / "perceptron_ep_vs_se.csv")
    assert sorted(pc.columns) == ["N", "alpha", "n_iter", "prior_p_pos", "source", "v", "x_id"]
    assert len(pc) == 2 * 4 * 3


def test_sharded_instances_single_rank(emulated_device, examples_on_path, capsys):  # noqa: F811
    import sharded_instances
    res = sharded_instances.main(["--instances", "3", "--n", "64", "--max-iter", "30"])
    assert res["r"]["x"].shape == (3, 64) and res["mse"].shape == (30, 3)
    assert "3 instances on 1 rank(s)" in capsys.readouterr().out
    sys.modules.pop("sharded_instances", None)
