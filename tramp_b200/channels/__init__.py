"""Channels on the EP hot path (reference tramp/channels/)."""
from ..base import Registry
from .base_channel import Channel
from .linear_channel import LinearChannel
from .gaussian_channel import GaussianChannel
from .activation import SgnChannel, AbsChannel
from .analytical_linear_channel import AnalyticalLinearChannel, MarchenkoPasturChannel

CHANNEL_CLASSES = Registry("channel", {
    "linear": LinearChannel,
    "marchenko": MarchenkoPasturChannel,
    "gaussian": GaussianChannel,
    "sgn": SgnChannel,
    "abs": AbsChannel,
})


def get_channel(channel_type, **kwargs):
    """reference channels/__init__.py:68-70."""
    return CHANNEL_CLASSES[channel_type](**kwargs)
