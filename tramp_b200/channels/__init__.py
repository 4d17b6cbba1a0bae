"""Channels on the EP hot path (reference tramp/channels/)."""
from .base_channel import Channel
from .linear_channel import LinearChannel
from .gaussian_channel import GaussianChannel
from .activation import SgnChannel, AbsChannel

CHANNEL_CLASSES = {
    "linear": LinearChannel,
    "gaussian": GaussianChannel,
    "sgn": SgnChannel,
    "abs": AbsChannel,
}


def get_channel(channel_type, **kwargs):
    """reference channels/__init__.py:68-70."""
    return CHANNEL_CLASSES[channel_type](**kwargs)
