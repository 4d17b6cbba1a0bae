"""EP and State-Evolution drivers (reference tramp/algos/)."""
from .expectation_propagation import ExpectationPropagation
from .message_passing import MessagePassing
from .state_evolution import StateEvolution
from .callbacks import (
    Callback, PassCallback, JoinCallback, LogProgress, TrackEvolution,
    TrackEstimate, TrackErrors, EarlyStoppingEP, EarlyStopping,
    TrackMessages, TrackObjective, TrackOverlaps,
)
from .initial_conditions import ConstantInit, NoisyInit, CustomInit
from .metrics import METRICS, mean_squared_error, sign_symmetric_mse, overlap
