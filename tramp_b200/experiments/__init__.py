"""Experiment helpers (reference tramp/experiments/): teacher-student
scenarios, grid runner, critical-alpha search.  Plotting helpers are out of
scope."""
from .teacher_student_scenario import (
    TeacherStudentScenario, BayesOptimalScenario, run_state_evolution,
    run_state_evolution_grid,
)
from .multiple_experiments import (
    run_experiments, simple_run_experiments, save_experiments, log_on_progress,
    get_experiments_from_kwargs,
)
from .critical_alpha import (
    binary_search, find_state_evolution_mse, find_critical_alpha,
)
