"""Teacher-student scenarios, EP and State Evolution (reference
tramp/experiments/teacher_student_scenario.py)."""
import logging
import pandas as pd

from ..algos.metrics import METRICS
from ..models import Model
from ..algos import (TrackErrors, TrackEvolution, JoinCallback, ExpectationPropagation,
                     StateEvolution)

logger = logging.getLogger(__name__)


class TeacherStudentScenario():
    """Implements teacher student scenario (reference :10-141, EP part).

    - teacher : Model instance or any object with a `.sample()` method
    - student : Model instance, generative student model
    - x_ids : ids of the variables to infer (signals)
    - y_ids : ids of the observed variables (measurements)
    """

    def __init__(self, teacher, student, x_ids=["x"], y_ids=["y"]):
        if not isinstance(student, Model):
            raise ValueError("student not a Model")
        try:
            sample = teacher.sample()      # reference :27 (advances the RNG once)
        except AttributeError:
            raise ValueError("teacher does not have a .sample() method")
        for x_id in x_ids:
            if x_id not in student.variable_ids:
                raise ValueError(f"x_id = {x_id} not in student variable_ids")
            if x_id not in sample:
                raise ValueError(f"x_id = {x_id} not in teacher variable_ids")
        for y_id in y_ids:
            if y_id not in student.variable_ids:
                raise ValueError(f"y_id = {y_id} not in  student variable_ids")
            if y_id not in sample:
                raise ValueError(f"y_id = {y_id} not in teacher variable_ids")
        self.x_ids = x_ids
        self.y_ids = y_ids
        self.teacher = teacher
        self.generative_student = student

    def setup(self, seed=0):
        sample = self.teacher.sample(seed)
        self.true_values = sample
        self.x_true = {x_id: sample[x_id] for x_id in self.x_ids}
        self.observations = {y_id: sample[y_id] for y_id in self.y_ids}
        self.student = self.generative_student.to_observed(self.observations)

    def run_all(self, source="EP,SE", metrics=["mse"], **algo_kwargs):
        "Get mse values as estimated by EP or SE (reference :54-82)"
        self.setup()
        records = []
        if "SE" in source:
            x_data = self.run_se(**algo_kwargs)
            records += [dict(source="SE", x_id=x_id, v=x_data[x_id]["v"], n_iter=x_data["n_iter"])
                        for x_id in self.x_ids]
        if "EP" in source:
            x_data = self.run_ep(**algo_kwargs)
            records += [dict(source="EP", x_id=x_id, v=x_data[x_id]["v"], n_iter=x_data["n_iter"])
                        for x_id in self.x_ids]
            x_pred = {x_id: x_data[x_id]["r"] for x_id in self.x_ids}
            score = self.compute_score(x_pred, metrics=metrics)
            records += [dict(source=metric, x_id=x_id, v=score[x_id][metric])
                        for metric in metrics for x_id in self.x_ids]
        return records

    def run_se(self, **algo_kwargs):
        """State Evolution of the observed student (its LinearChannel enters
        through its own spectrum); reference :84-89."""
        se = StateEvolution(self.student)
        se.iterate(**algo_kwargs)
        x_data = se.get_variables_data(self.x_ids)
        x_data["n_iter"] = se.n_iter
        self.se = se
        return x_data

    def run_ep(self, **algo_kwargs):
        ep = ExpectationPropagation(self.student)
        ep.iterate(**algo_kwargs)
        x_data = ep.get_variables_data(self.x_ids)
        x_data["n_iter"] = ep.n_iter
        self.x_pred = {x_id: x_data[x_id]["r"] for x_id in self.x_ids}
        self.ep = ep
        return x_data

    def ep_convergence(self, metrics, **algo_kwargs):
        track = TrackErrors(true_values=self.x_true, metrics=metrics)
        evo = TrackEvolution(ids=self.x_ids)
        callbacks = [track, evo]
        if "callback" in algo_kwargs:
            callbacks.append(algo_kwargs["callback"])
        algo_kwargs["callback"] = JoinCallback(callbacks)
        try:
            self.run_ep(**algo_kwargs)
        except Exception as e:
            logger.error(e)
        df = pd.merge(track.get_dataframe(), evo.get_dataframe(), on=["id", "iter"])
        if not self.ep.batched:
            for y in ["v"] + metrics:
                df[y] = df[y].clip(0, 2)
        return df

    def se_convergence(self, **algo_kwargs):
        "v of the x_ids along the SE iterations (reference :117-130)"
        evo = TrackEvolution(ids=self.x_ids)
        callbacks = [evo]
        if "callback" in algo_kwargs:
            callbacks.append(algo_kwargs["callback"])
        algo_kwargs["callback"] = JoinCallback(callbacks)
        try:
            self.run_se(**algo_kwargs)
        except Exception as e:
            logger.error(e)
        df = evo.get_dataframe()
        df["v"] = df["v"].clip(0, 2)
        return df

    def compute_score(self, x_pred, metrics=["mse"]):
        return {x_id: {metric: METRICS[metric](self.x_true[x_id], x_pred[x_id]) for metric in metrics}
                for x_id in self.x_ids}


class BayesOptimalScenario(TeacherStudentScenario):
    """Same generative model for teacher and student (reference :143-155)."""

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
