"""Initial values of the messages (reference tramp/algos/initial_conditions.py).

The driver asks `initializer.init(key, shape, variable_id, direction)` for every
edge, `key` being "a" (precision) or "b" (natural mean); `shape` is the
variable's shape -- (N,) for one instance, (B, N) for a batch -- and is None
under State Evolution, which only asks for "a".  The eight edges are visited in
the reference's order -- `message_dag.edges()`, node by node: e1, e2, e8, e3, e7, e4, e6,
e5 in the numbering of SURVEY 3.3 -- "a" before "b", so that `NoisyInit` consumes numpy's
global random stream exactly as the reference does (same seed, same initial messages).
"""
import numpy as np

from ..base import ReprMixin


def _filled(shape, value):
    if shape is None:
        raise AssertionError("the shape of the variable is unknown (call model.init_shapes())")
    return np.full(shape, value, dtype=float)


class InitialConditions(ReprMixin):
    """Dispatches `init("a" | "b", ...)` to `init_a` / `init_b` (reference :5-10)."""

    def init(self, message_key, shape, id, direction):
        handler = {"a": self.init_a, "b": self.init_b}.get(message_key)
        return None if handler is None else handler(shape, id, direction)


class ConstantInit(InitialConditions):
    """Every edge starts at the same (a, b) (reference :13-24)."""

    def __init__(self, a=0, b=0):
        self.a, self.b = a, b
        self.repr_init()

    def init_a(self, shape, id, direction):
        return self.a

    def init_b(self, shape, id, direction):
        return _filled(shape, self.b)


class NoisyInit(InitialConditions):
    """Gaussian initial messages, a ~ N(a_mean, a_var) (one draw per edge) and
    b ~ N(b_mean, b_var) elementwise, from numpy's global RNG (reference :27-42)."""

    def __init__(self, a_mean=0, a_var=0, b_mean=0, b_var=1):
        self.a_mean, self.a_var = a_mean, a_var
        self.b_mean, self.b_var = b_mean, b_var
        self.repr_init()
        self.a_sigma, self.b_sigma = np.sqrt(a_var), np.sqrt(b_var)

    def init_a(self, shape, id, direction):
        noise = np.random.standard_normal()
        return self.a_mean + self.a_sigma * noise

    def init_b(self, shape, id, direction):
        if shape is None:
            raise AssertionError("the shape of the variable is unknown (call model.init_shapes())")
        noise = np.random.standard_normal(shape)
        return self.b_mean + self.b_sigma * noise


class CustomInit(InitialConditions):
    """Chosen values on chosen edges, constants elsewhere (reference :45-85).

    a_init / b_init: lists of (variable id, direction, value); an entry applies to
    both edges of that variable with that direction (factor -> variable and
    variable -> factor).  a, b: the constants used for every other edge.

    Like the reference (`{id: {direction: a} for id, direction, a in a_init}`, :63-66), ONE entry
    per variable id is kept: a later entry for the same id replaces an earlier one, also when its
    direction differs."""

    def __init__(self, a_init=None, b_init=None, a=0, b=0):
        self.a_init = self._table(a_init)
        self.b_init = self._table(b_init)
        self.a, self.b = a, b
        self.repr_init()

    @staticmethod
    def _table(entries):
        return {variable_id: {direction: value} for variable_id, direction, value in (entries or [])}

    def init_a(self, shape, id, direction):
        return self.a_init.get(id, {}).get(direction, self.a)

    def init_b(self, shape, id, direction):
        chosen = self.b_init.get(id, {}).get(direction)
        if chosen is None:
            return _filled(shape, self.b)
        if shape is None or chosen.shape != shape:
            raise AssertionError(f"b_init of {id!r} ({direction}) has shape {chosen.shape}, expected {shape}")
        return chosen
