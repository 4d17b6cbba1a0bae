"""Drivers of the two message-passing algorithms and what is passed to them
(reference tramp/algos/):

* `ExpectationPropagation` -- EP on one instance or a batch of instances, the
  sweep running on the device (message_passing.py, tramp_b200/csrc/trb_sweep.cu,
  trb_persist.cu);
* `StateEvolution` -- its scalar average-case twin (state_evolution.py,
  tramp_b200/csrc/trb_se.cu);
* callbacks (`callback(algo, i, max_iter) -> stop?`), initial conditions and
  error metrics, under the reference's names.
"""
from . import callbacks as _callbacks
from . import initial_conditions as _initial_conditions
from . import metrics as _metrics
from .message_passing import MessagePassing
from .expectation_propagation import ExpectationPropagation
from .state_evolution import StateEvolution

_EXPORTS = {
    _callbacks: ("Callback", "PassCallback", "JoinCallback", "LogProgress", "TrackMessages",
                 "TrackObjective", "TrackOverlaps", "TrackEvolution", "TrackEstimate", "TrackErrors",
                 "EarlyStoppingEP", "EarlyStopping"),
    _initial_conditions: ("InitialConditions", "ConstantInit", "NoisyInit", "CustomInit"),
    _metrics: ("METRICS", "mean_squared_error", "sign_symmetric_mse", "overlap"),
}
__all__ = ["MessagePassing", "ExpectationPropagation", "StateEvolution"]
for _module, _names in _EXPORTS.items():
    for _name in _names:
        globals()[_name] = getattr(_module, _name)
        __all__.append(_name)
del _module, _names, _name
