"""tramp_b200 -- B200-native expectation propagation behind Tree-AMP's model/plugin API.

Same names as the reference package (`tramp.priors`, `tramp.likelihoods`,
`tramp.channels`, `tramp.variables`, `tramp.models`, `tramp.algos`,
`tramp.ensembles`, `tramp.experiments`), for the EP sweep of
`prior @ V @ LinearChannel @ V @ likelihood`.  All EP arithmetic runs in
hand-written sm_100a CUDA kernels (tramp_b200/csrc) reached through the C ABI
of include/tramp_b200.h; there is no CPU fallback.
"""
__version__ = "0.1.0"
