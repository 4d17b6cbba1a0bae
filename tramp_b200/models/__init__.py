"""Model construction (reference tramp/models/): DAG algebra with `@`, Model,
glm_generative.  Chain DAGs only -- the EP hot path is
prior @ V @ LinearChannel @ V @ likelihood."""
from .dag_algebra import DAG, ModelDAG, channel2likelihood
from .base_model import Model
from .generalized_linear_model import glm_generative, glm_state_evolution
