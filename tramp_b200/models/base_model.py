"""Model: a validated, topologically ordered factor graph (reference
tramp/models/base_model.py:28-148).

Everything a model computes by itself -- a joint sample, the shapes of the
variables, their second moments -- is one forward sweep over the factors in
topological order, each factor mapping what its predecessor variables carry to
what its successor variables carry; `_forward_sweep` is that sweep.
"""
import numpy as np

from ..base import ReprMixin, Variable, Factor
from .dag_algebra import ModelDAG


def to_list(X):
    """A factor with one output returns it bare, with several as a tuple."""
    return list(X) if isinstance(X, tuple) else [X]


def check_variable_ids(variables):
    ids = [variable.id for variable in variables]
    for position, variable_id in enumerate(ids):
        if variable_id is None:
            raise ValueError(f"missing id for the i={position} {variables[position]} ")
    if len(set(ids)) != len(ids):
        raise ValueError("variable ids are not unique")


class Model(ReprMixin):
    def __init__(self, model_dag):
        if not isinstance(model_dag, ModelDAG):
            raise TypeError(f"model_dag {model_dag} is not a ModelDAG")
        self.repr_init()
        self.model_dag = model_dag
        self.dag = model_dag.dag.copy()
        order = self.forward_ordering = self.dag.topological_sort()
        self.variables = [node for node in order if isinstance(node, Variable)]
        self.factors = [node for node in order if isinstance(node, Factor)]
        check_variable_ids(self.variables)
        self.variable_ids = [variable.id for variable in self.variables]
        # factors are numbered in topological order (reference :23-25, :49)
        self.factor_ids = []
        for number, factor in enumerate(self.factors):
            factor.id = f"f_{number}"
            self.factor_ids.append(factor.id)
        self.n_variables, self.n_factors = len(self.variables), len(self.factors)

    def to_observed(self, observations):
        """The model with the given leaves observed, their channels turned into
        likelihoods (reference :60-69)."""
        return Model(self.model_dag.to_observed(observations))

    def _forward_sweep(self, emit, start=None):
        """carried[variable] for every variable: emit(factor, [carried of the
        predecessors]) -> one value per successor variable (or None to skip the
        factor)."""
        carried = dict(start or {})
        for factor in self.factors:
            inputs = [carried[v] for v in self.dag.predecessors(factor)]
            outputs = emit(factor, inputs)
            if outputs is None:
                continue
            for value, variable in zip(to_list(outputs), self.dag.successors(factor)):
                carried[variable] = value
        return carried

    def sample(self, seed=0):
        """Forward (ancestral) sample {variable id: array} (reference :71-94).  Draws
        from numpy's global RNG, re-seeded only when seed != 0."""
        if seed != 0:
            np.random.seed(seed)
        drawn = self._forward_sweep(lambda factor, X_prev: factor.sample(*X_prev))
        return {variable.id: drawn[variable] for variable in self.variables}

    def init_shapes(self):
        """Shape of every variable, stored on the graph (reference :96-109).  A factor
        that cannot tell its output shape (`infer_shape`) is sampled on arrays of
        ones, which advances the RNG exactly as the reference does."""
        def shape_of(factor, prev_shapes):
            infer = getattr(factor, "infer_shape", None)
            shapes = infer(*prev_shapes) if infer is not None else None
            if shapes is None:
                outputs = to_list(factor.sample(*[np.ones(shape) for shape in prev_shapes]))
                shapes = [x.shape for x in outputs]
            return tuple(tuple(shape) for shape in shapes)
        for variable, shape in self._forward_sweep(shape_of).items():
            self.dag.node[variable].update(shape=shape)

    def init_second_moments(self):
        """Second moment tau of every variable, stored on the graph (reference
        :111-124); leaf factors (likelihoods) emit nothing."""
        def tau_of(factor, tau_prev):
            if not factor.n_next:
                return None
            tau_next = factor.second_moment(*tau_prev)
            return tuple(tau_next) if isinstance(tau_next, (tuple, list)) else tau_next
        for variable, tau in self._forward_sweep(tau_of).items():
            self.dag.node[variable].update(tau=tau)

    def _variable_attribute(self, key):
        return {variable.id: self.dag.node[variable][key] for variable in self.variables}

    def get_second_moments(self):
        return self._variable_attribute("tau")

    def get_shapes(self):
        return self._variable_attribute("shape")
