"""What the reference's example scripts call (reference tramp/experiments/):
teacher-student scenarios running EP and State Evolution, the grid runner that
turns a `run(**point)` function into a DataFrame over a Cartesian grid, and the
critical-alpha search on top of State Evolution.  Plotting helpers (`qplot`,
`plot_compare`, ...) are out of scope.

Additions over the reference: `run_state_evolution_grid` (a list of models in one
kernel launch, sharded over ranks under torch.distributed), `run_ep_sharded` (a batch
of independent EP instances sharded over the GPUs of a box) and the `grid=` option of
`find_critical_alpha`.
"""
from . import critical_alpha as _critical_alpha
from . import multiple_experiments as _multiple_experiments
from . import teacher_student_scenario as _scenario

_EXPORTS = {
    _scenario: ("TeacherStudentScenario", "BayesOptimalScenario", "run_state_evolution",
                "run_state_evolution_grid", "run_ep_sharded"),
    _multiple_experiments: ("run_experiments", "simple_run_experiments", "save_experiments",
                            "log_on_progress", "get_experiments_from_kwargs"),
    _critical_alpha: ("binary_search", "grid_search", "find_state_evolution_mse", "find_critical_alpha"),
}
__all__ = []
for _module, _names in _EXPORTS.items():
    for _name in _names:
        globals()[_name] = getattr(_module, _name)
        __all__.append(_name)
del _module, _names, _name
