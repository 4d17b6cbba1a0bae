"""Parity report: public API on the GPU vs the golden vectors of the reference.
Prints, per golden sweep and GEMV implementation, the max relative deviation of
every compared quantity.  Usage: python tools/parity_report.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.test_gpu_api import _build, _SeqInit, _configs
from tramp_b200.algos import ExpectationPropagation, TrackErrors, TrackEvolution, JoinCallback

sw = np.load("tests/golden/sweeps.npz")
report = {}
for cfg in _configs(sw):
    name = cfg["name"]
    for impl in (1, 2):
        ep = ExpectationPropagation(_build(cfg, sw, name))
        ep.gemv_impl = impl
        track, evo = TrackErrors({"x": sw[name + "_x"]}), TrackEvolution()
        init = _SeqInit(sw, name) if cfg.get("init") == "noisy" else None
        ep.iterate(max_iter=cfg["n_iter"], callback=JoinCallback([track, evo]), initializer=init,
                   damping=cfg["damping"])
        mse = np.array([e["mse"] for e in track.errors]); df = evo.get_dataframe()
        d = ep.get_variables_data()
        rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
        relmax = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
        x = sw[name + "_x"]
        r = dict(mse=rel(mse, sw[name + "_mse"]),
                 mse_cond=float(np.max(np.abs(mse - sw[name + "_mse"]) / (2 * np.sqrt(sw[name + "_mse"] * np.mean(x**2))))),
                 vx=rel(df[df.id == "x"].v.values, sw[name + "_vx"]),
                 vz=rel(df[df.id == "z"].v.values, sw[name + "_vz"]),
                 rx=relmax(d["x"]["r"], sw[name + "_rx"]), rz=relmax(d["z"]["r"], sw[name + "_rz"]))
        for k in range(1, 9):
            a, b = ep._edge(f"e{k}")
            r[f"e{k}_a"] = rel(np.asarray(a), sw[f"{name}_e{k}_a"])
            r[f"e{k}_b"] = relmax(b, sw[f"{name}_e{k}_b"])
        report[f"{name}/impl{impl}"] = r
        print(name, impl, {k: f"{v:.1e}" for k, v in r.items()}, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(report, open("gpurun_out/parity_report.json", "w"), indent=1)
