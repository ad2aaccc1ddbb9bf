from .certificate import Certificate, L2Certificate  # noqa: F401
from .smooth import Smooth  # noqa: F401
