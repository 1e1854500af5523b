from .discriminator import Discriminator  # noqa: F401
