"""Model (reference tramp/models/base_model.py:28-109)."""
import numpy as np

from ..base import ReprMixin, Variable, Factor
from .dag_algebra import ModelDAG


def to_list(X):
    if not isinstance(X, tuple):
        X = (X,)
    return list(X)


def check_variable_ids(variables):
    for i, variable in enumerate(variables):
        if variable.id is None:
            raise ValueError(f"missing id for the i={i} {variable} ")
    if len(set(v.id for v in variables)) != len(variables):
        raise ValueError("variable ids are not unique")


class Model(ReprMixin):
    def __init__(self, model_dag):
        if not isinstance(model_dag, ModelDAG):
            raise TypeError(f"model_dag {model_dag} is not a ModelDAG")
        self.repr_init()
        self.model_dag = model_dag
        self.dag = model_dag.dag.copy()
        self.forward_ordering = self.dag.topological_sort()
        self.variables = [n for n in self.forward_ordering if isinstance(n, Variable)]
        self.variable_ids = [v.id for v in self.variables]
        check_variable_ids(self.variables)
        self.n_variables = len(self.variables)
        self.factors = [n for n in self.forward_ordering if isinstance(n, Factor)]
        for idx, factor in enumerate(self.factors):   # reference :23-25
            factor.id = f"f_{idx}"
        self.factor_ids = [f.id for f in self.factors]
        self.n_factors = len(self.factors)

    def to_observed(self, observations):
        """reference :60-69."""
        return Model(self.model_dag.to_observed(observations))

    def sample(self, seed=0):
        "Forward sampling of the model (reference :71-94; numpy global RNG, reseeded only if seed != 0)"
        if seed != 0:
            np.random.seed(seed)
        X = {}
        for factor in self.factors:
            X_prev = [X[v] for v in self.dag.predecessors(factor)]
            X_next = to_list(factor.sample(*X_prev))
            for x, variable in zip(X_next, self.dag.successors(factor)):
                X[variable] = x
        return {variable.id: X[variable] for variable in self.variables}

    def init_shapes(self):
        "Compute variable shapes in place (reference :96-109; calls every factor.sample, so it advances the RNG)"
        for factor in self.factors:
            prev_shapes = [self.dag.node[v]["shape"] for v in self.dag.predecessors(factor)]
            infer = getattr(factor, "infer_shape", None)
            shapes = infer(*prev_shapes) if infer is not None else None
            if shapes is None:
                X_next = to_list(factor.sample(*[np.ones(s) for s in prev_shapes]))
                shapes = [x.shape for x in X_next]
            for shape, variable in zip(shapes, self.dag.successors(factor)):
                self.dag.node[variable].update(shape=tuple(shape))

    def init_second_moments(self):
        "Second moment tau of every variable, in place (reference :111-124)"
        for factor in self.factors:
            if factor.n_next:
                tau_prev = [self.dag.node[v]["tau"] for v in self.dag.predecessors(factor)]
                tau_next = to_list(factor.second_moment(*tau_prev))
                for tau, variable in zip(tau_next, self.dag.successors(factor)):
                    self.dag.node[variable].update(tau=tau)

    def get_second_moments(self):
        return {v.id: self.dag.node[v]["tau"] for v in self.variables}

    def get_shapes(self):
        return {v.id: self.dag.node[v]["shape"] for v in self.variables}
