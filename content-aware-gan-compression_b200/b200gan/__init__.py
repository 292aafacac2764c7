"""Host side of the B200-native StyleGAN2 generator hot path."""
from . import config  # noqa: F401
