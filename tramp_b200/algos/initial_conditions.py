"""Initial conditions of the messages (reference tramp/algos/initial_conditions.py).

`shape` is the variable's shape: (N,) for one instance, (B, N) for a batch (then
`a` may be a scalar shared by all instances or an array of B values)."""
import numpy as np
from ..base import ReprMixin


class InitialConditions(ReprMixin):
    def init(self, message_key, shape, id, direction):
        if message_key == "a":
            return self.init_a(shape, id, direction)
        if message_key == "b":
            return self.init_b(shape, id, direction)


class ConstantInit(InitialConditions):
    """reference initial_conditions.py:13-24."""

    def __init__(self, a=0, b=0):
        self.a = a
        self.b = b
        self.repr_init()

    def init_a(self, shape, id, direction):
        return self.a

    def init_b(self, shape, id, direction):
        assert shape is not None
        return self.b * np.ones(shape)


class NoisyInit(InitialConditions):
    """reference initial_conditions.py:27-42.  Draws use numpy's global RNG in
    the edge order e1..e8, `a` then `b` for each edge."""

    def __init__(self, a_mean=0, a_var=0, b_mean=0, b_var=1):
        self.a_mean = a_mean
        self.a_var = a_var
        self.b_mean = b_mean
        self.b_var = b_var
        self.repr_init()
        self.a_sigma = np.sqrt(a_var)
        self.b_sigma = np.sqrt(b_var)

    def init_a(self, shape, id, direction):
        return self.a_mean + self.a_sigma * np.random.standard_normal()

    def init_b(self, shape, id, direction):
        assert shape is not None
        return self.b_mean + self.b_sigma * np.random.standard_normal(shape)


class CustomInit(InitialConditions):
    """Custom init on variables (reference initial_conditions.py:45-85).

    - a_init: list of (variable.id, direction, a) tuples
    - b_init: list of (variable.id, direction, b) tuples
    - a, b : default constants
    """

    def __init__(self, a_init=None, b_init=None, a=0, b=0):
        a_init = a_init or []
        self.a_init = {}
        for id, direction, a_ in a_init:
            self.a_init.setdefault(id, {})[direction] = a_
        b_init = b_init or []
        self.b_init = {}
        for id, direction, b_ in b_init:
            self.b_init.setdefault(id, {})[direction] = b_
        self.a = a
        self.b = b
        self.repr_init()

    def init_a(self, shape, id, direction):
        try:
            return self.a_init[id][direction]
        except KeyError:
            return self.a

    def init_b(self, shape, id, direction):
        assert shape is not None
        try:
            b = self.b_init[id][direction]
            assert b.shape == shape
        except KeyError:
            b = self.b * np.ones(shape)
        return b
