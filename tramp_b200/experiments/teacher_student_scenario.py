"""Teacher-student scenarios (reference tramp/experiments/teacher_student_scenario.py).

A teacher model generates a signal and its measurements, a student model -- the
same model in the Bayes-optimal case -- is given the measurements and estimates
the signal, by Expectation Propagation on the instance (`run_ep`) and, for the
average case, by State Evolution (`run_se`).  Both run on the GPU
(tramp_b200.algos); this module is the glue the reference's examples call.
"""
import logging
import pandas as pd

from ..algos.metrics import METRICS
from ..models import Model
from ..algos import (TrackErrors, TrackEvolution, JoinCallback, ExpectationPropagation,
                     StateEvolution)

logger = logging.getLogger(__name__)


def _with_trackers(trackers, algo_kwargs):
    """algo_kwargs with `trackers` joined in front of the caller's callback."""
    callbacks = list(trackers)
    if "callback" in algo_kwargs:
        callbacks.append(algo_kwargs["callback"])
    return dict(algo_kwargs, callback=JoinCallback(callbacks))


class TeacherStudentScenario():
    """teacher : a Model, or any object with a `.sample()` method returning
                {variable id: array}
    student : Model, the generative model the student assumes
    x_ids   : ids of the variables to infer (signals)
    y_ids   : ids of the observed variables (measurements)
    (reference :10-141)"""

    def __init__(self, teacher, student, x_ids=["x"], y_ids=["y"]):
        if not isinstance(student, Model):
            raise ValueError("student not a Model")
        if not hasattr(teacher, "sample"):
            raise ValueError("teacher does not have a .sample() method")
        # one draw, only to learn the teacher's variable ids: like the reference
        # (:27) this advances numpy's global RNG before `setup`
        known_to_teacher = teacher.sample()
        for role, ids in (("x_id", x_ids), ("y_id", y_ids)):
            for variable_id in ids:
                if variable_id not in student.variable_ids:
                    raise ValueError(f"{role} = {variable_id} not in student variable_ids")
                if variable_id not in known_to_teacher:
                    raise ValueError(f"{role} = {variable_id} not in teacher variable_ids")
        self.x_ids, self.y_ids = x_ids, y_ids
        self.teacher, self.generative_student = teacher, student

    def setup(self, seed=0):
        "The teacher draws the instance; the student gets the measurements (reference :45-52)"
        self.true_values = self.teacher.sample(seed)
        self.x_true = {x_id: self.true_values[x_id] for x_id in self.x_ids}
        self.observations = {y_id: self.true_values[y_id] for y_id in self.y_ids}
        self.student = self.generative_student.to_observed(self.observations)

    # ---------------------------------------------------------------- one run
    def _run(self, algo, algo_kwargs):
        algo.iterate(**algo_kwargs)
        x_data = algo.get_variables_data(self.x_ids)
        x_data["n_iter"] = algo.n_iter
        return x_data

    def run_se(self, **algo_kwargs):
        """State Evolution of the observed student -- its LinearChannel enters
        through its own spectrum (reference :84-89)."""
        self.se = StateEvolution(self.student)      # kept: a failed run can still be inspected
        return self._run(self.se, algo_kwargs)

    def run_ep(self, **algo_kwargs):
        "Expectation Propagation on the instance (reference :91-97)"
        self.ep = ExpectationPropagation(self.student)
        x_data = self._run(self.ep, algo_kwargs)
        self.x_pred = {x_id: x_data[x_id]["r"] for x_id in self.x_ids}
        return x_data

    def run_all(self, source="EP,SE", metrics=["mse"], **algo_kwargs):
        """Records of the variance predicted by SE, the variance EP reports and the
        error EP actually makes (reference :54-82)."""
        self.setup()
        records = []

        def variance_records(name, x_data):
            return [dict(source=name, x_id=x_id, v=x_data[x_id]["v"], n_iter=x_data["n_iter"])
                    for x_id in self.x_ids]
        if "SE" in source:
            records += variance_records("SE", self.run_se(**algo_kwargs))
        if "EP" in source:
            records += variance_records("EP", self.run_ep(**algo_kwargs))
            score = self.compute_score(self.x_pred, metrics=metrics)
            records += [dict(source=metric, x_id=x_id, v=score[x_id][metric])
                        for metric in metrics for x_id in self.x_ids]
        return records

    # ----------------------------------------------------------- trajectories
    def ep_convergence(self, metrics, **algo_kwargs):
        "Per-iteration errors and variances of EP (reference :99-115)"
        errors = TrackErrors(true_values=self.x_true, metrics=metrics)
        evolution = TrackEvolution(ids=self.x_ids)
        try:
            self.run_ep(**_with_trackers([errors, evolution], algo_kwargs))
        except Exception as e:
            logger.error(e)
        df = pd.merge(errors.get_dataframe(), evolution.get_dataframe(), on=["id", "iter"])
        if not self.ep.batched:
            for column in ["v"] + metrics:
                df[column] = df[column].clip(0, 2)
        return df

    def se_convergence(self, **algo_kwargs):
        "Per-iteration variances of State Evolution (reference :117-130)"
        evolution = TrackEvolution(ids=self.x_ids)
        try:
            self.run_se(**_with_trackers([evolution], algo_kwargs))
        except Exception as e:
            logger.error(e)
        df = evolution.get_dataframe()
        df["v"] = df["v"].clip(0, 2)
        return df

    def compute_score(self, x_pred, metrics=["mse"]):
        return {x_id: {metric: METRICS[metric](self.x_true[x_id], x_pred[x_id]) for metric in metrics}
                for x_id in self.x_ids}


class BayesOptimalScenario(TeacherStudentScenario):
    """The student knows the teacher's model (reference :143-155)."""

    def __init__(self, model, x_ids=["x"], y_ids=["y"]):
        super().__init__(teacher=model, student=model, x_ids=x_ids, y_ids=y_ids)


def run_state_evolution(x_ids, model, **algo_kwargs):
    """Records (x_id, v, n_iter) of one State Evolution run (reference :158-178)."""
    se = StateEvolution(model)
    se.iterate(**algo_kwargs)
    x_data = se.get_variables_data(ids=x_ids)
    return [dict(x_id=x_id, v=x_data[x_id]["v"], n_iter=se.n_iter) for x_id in x_ids]


def run_state_evolution_grid(x_ids, models, group=None, **algo_kwargs):
    """The same for a list of models (e.g. `glm_state_evolution` over a grid of
    alpha), all in ONE kernel launch: one CTA per model runs its whole recursion.
    Returns one list of records per model, `n_iter` being that model's own count.

    Under `torch.distributed` (one process per GPU) the grid is sharded by
    contiguous blocks of models, like EP instances: every rank runs
    its block with no data-path collective and the records are all-gathered at
    the end, so every rank returns the whole grid."""
    models = list(models)
    world, rank = 1, 0
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(group), dist.get_rank(group)
    except ImportError:
        dist = None
    if world > 1:
        from ..distributed import instance_shard
        start, stop = instance_shard(len(models), rank, world)
        local = _state_evolution_records(x_ids, models[start:stop], **algo_kwargs) if stop > start else []
        gathered = [None] * world
        dist.all_gather_object(gathered, local, group=group)
        return [records for block in gathered for records in block]
    return _state_evolution_records(x_ids, models, **algo_kwargs)


def _state_evolution_records(x_ids, models, **algo_kwargs):
    se = StateEvolution(list(models))
    se.iterate(**algo_kwargs)
    x_data = se.get_variables_data(ids=x_ids)
    return [[dict(x_id=x_id, v=float(x_data[x_id]["v"][g]), n_iter=int(se.n_iter_per_problem[g]))
             for x_id in x_ids] for g in range(se.G)]


def run_ep_sharded(build_model, n_instances, x_true=None, group=None, **algo_kwargs):
    """Expectation Propagation on `n_instances` independent instances sharded over the
    ranks of `group` (one process per GPU) by contiguous blocks -- SURVEY 8e: instances
    share nothing, so every rank runs the device-resident sweep on its own block with no
    data-path collective, and only the results are all-gathered at the end.

    build_model(start, stop) -> the batched, observed Model of instances [start, stop)
        (each rank builds, factorises and keeps only its own block);
    x_true(start, stop) -> their signals {"x": [stop - start, N]} for the per-iteration mse
        (optional);
    algo_kwargs -> `ExpectationPropagation.iterate` (max_iter, damping, callback, ...).

    Returns, identical on every rank: dict(r={id: [n_instances, N_id]}, v={id:
    [n_instances]}, n_iter=[n_instances], mse=[max_iter, n_instances] or None), an
    instance's mse being NaN after it stopped.  Without torch.distributed it is the
    plain single-GPU run."""
    import numpy as np
    from ..distributed import instance_shard, gather_records
    from .. import ops
    t = ops.torch()
    world, rank = 1, 0
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(group), dist.get_rank(group)
    except ImportError:
        pass
    start, stop = instance_shard(n_instances, rank, world)
    max_iter = int(algo_kwargs.get("max_iter", 200))
    local = {}
    if stop > start:
        ep = ExpectationPropagation(build_model(start, stop))
        if x_true is not None:
            errors = TrackErrors(true_values=x_true(start, stop), metrics=["mse"])
            if "callback" not in algo_kwargs:          # tracking must not switch the default stopper off
                algo_kwargs = dict(algo_kwargs, callback=ep.default_stopping)
            algo_kwargs = _with_trackers([errors], algo_kwargs)
        ep.iterate(**algo_kwargs)
        data = ep.get_variables_data()
        for vid, d in data.items():
            local["r_" + vid] = np.atleast_2d(d["r"]).T                       # instances last
            local["v_" + vid] = np.atleast_1d(d["v"])[None, :]
        n_iter = getattr(ep, "n_iter_per_instance", None)
        local["n_iter"] = np.atleast_1d(ep.n_iter if n_iter is None else n_iter).astype(float)[None, :]
        if x_true is not None:
            mse = np.full((max_iter, stop - start), np.nan)
            for e in errors.errors:
                mse[e["iter"]] = e["mse"]
            local["mse"] = mse
    # a rank without instances still takes part in the gathers, with the shapes of the others
    shapes = {k: v.shape[:-1] for k, v in local.items()}
    if world > 1:
        every = [None] * world
        dist.all_gather_object(every, shapes, group=group)
        shapes = next(s for s in every if s)
    out = {}
    for key in sorted(shapes):
        mine = local.get(key)
        if mine is None:
            mine = np.zeros(tuple(shapes[key]) + (0,))
        full = gather_records(t.as_tensor(mine, dtype=t.float64, device=ops.device()), group=group)
        out[key] = full.cpu().numpy()
    ids = sorted(k[2:] for k in out if k.startswith("r_"))
    return dict(r={vid: out["r_" + vid].T for vid in ids}, v={vid: out["v_" + vid][0] for vid in ids},
                n_iter=out["n_iter"][0].astype(int), mse=out.get("mse"))
