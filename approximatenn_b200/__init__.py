"""approximatenn_b200 — B200-native randomized all-points kNN behind approximateNN's C API."""
from .api import Backend, Result, Save, SaveT, srandom  # noqa: F401
