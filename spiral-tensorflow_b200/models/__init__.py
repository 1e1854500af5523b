from .discriminator import Discriminator  # noqa: F401
from .policy import Policy  # noqa: F401
