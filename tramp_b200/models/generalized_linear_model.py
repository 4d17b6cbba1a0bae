"""glm_generative, glm_state_evolution (reference tramp/models/generalized_linear_model.py:8-55)."""
from ..channels import get_channel
from ..priors import get_prior
from ..ensembles import get_ensemble
from ..likelihoods import get_likelihood
from ..variables import SISOVariable as V, SILeafVariable as O


def get_kwargs(target, kwargs):
    n_char = len(target) + 1
    return {key[n_char:]: val for key, val in kwargs.items() if key.startswith(target)}


def glm_generative(N, alpha, ensemble_type, prior_type, output_type, **kwargs):
    "Build a generative Generalized Linear Model (W is drawn first, from the global RNG)"
    M = int(alpha * N)
    ensemble = get_ensemble(ensemble_type, M=M, N=N, **get_kwargs("ensemble", kwargs))
    F = ensemble.generate()
    prior = get_prior(size=N, prior_type=prior_type, **get_kwargs("prior", kwargs))
    linear = get_channel("linear", W=F, name="F")
    output = get_channel(channel_type=output_type, **get_kwargs("output", kwargs))
    return (prior @ V(id="x") @ linear @ V(id="z") @ output @ O(id="y")).to_model()


def glm_state_evolution(alpha, prior_type, output_type, **kwargs):
    """GLM used only for State Evolution: the linear channel is known through the
    Marchenko-Pastur law, the likelihood carries no data (reference :37-55)."""
    prior = get_prior(size=None, prior_type=prior_type, **get_kwargs("prior", kwargs))
    linear = get_channel("marchenko", alpha=alpha, name="F")
    output = get_likelihood(y=None, y_name="y", likelihood_type=output_type,
                            **get_kwargs("output", kwargs))
    return (prior @ V(id="x") @ linear @ V(id="z") @ output).to_model()
