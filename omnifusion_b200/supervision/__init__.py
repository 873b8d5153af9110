"""Training-side losses of the reference's ``supervision`` package."""
from . import direct  # noqa: F401
