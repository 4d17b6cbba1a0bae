"""Generalized linear models y = output(F x) (reference
tramp/models/generalized_linear_model.py:8-55): the generative model with an
explicit matrix, for EP, and its State-Evolution twin in which the matrix is
replaced by its spectral law.

Keyword arguments are routed by prefix: `prior_rho=0.1` reaches the prior as
`rho=0.1`, `output_var=...` the output channel, `ensemble_...` the matrix
ensemble.
"""
from ..channels import get_channel
from ..priors import get_prior
from ..ensembles import get_ensemble
from ..likelihoods import get_likelihood
from ..variables import SISOVariable, SILeafVariable


def get_kwargs(target, kwargs):
    """The entries of kwargs addressed to `target`, prefix `target_` removed."""
    prefix = target + "_"
    return {key[len(prefix):]: value for key, value in kwargs.items() if key.startswith(prefix)}


def _chain(prior, linear, output, observed_leaf):
    dag = prior @ SISOVariable(id="x") @ linear @ SISOVariable(id="z") @ output
    if observed_leaf:
        dag = dag @ SILeafVariable(id="y")
    return dag.to_model()


def glm_generative(N, alpha, ensemble_type, prior_type, output_type, **kwargs):
    """Generative GLM with M = int(alpha N) measurements.  The matrix is drawn
    FIRST, from numpy's global RNG (seed parity with the reference, :20-23)."""
    ensemble = get_ensemble(ensemble_type, M=int(alpha * N), N=N, **get_kwargs("ensemble", kwargs))
    F = ensemble.generate()
    return _chain(prior=get_prior(size=N, prior_type=prior_type, **get_kwargs("prior", kwargs)),
                  linear=get_channel("linear", W=F, name="F"),
                  output=get_channel(channel_type=output_type, **get_kwargs("output", kwargs)),
                  observed_leaf=True)


def glm_state_evolution(alpha, prior_type, output_type, **kwargs):
    """The same GLM for State Evolution only (reference :37-55): sizes are
    irrelevant (`size=None`, `y=None`), the linear channel is known through the
    Marchenko-Pastur law at ratio alpha and the output is already a likelihood."""
    return _chain(prior=get_prior(size=None, prior_type=prior_type, **get_kwargs("prior", kwargs)),
                  linear=get_channel("marchenko", alpha=alpha, name="F"),
                  output=get_likelihood(y=None, y_name="y", likelihood_type=output_type,
                                        **get_kwargs("output", kwargs)),
                  observed_leaf=False)
