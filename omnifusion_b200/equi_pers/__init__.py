"""Drop-in replacements for the reference's ``equi_pers`` resamplers."""
from .equi2pers_v3 import equi2pers  # noqa: F401
from .pers2equi_v3 import pers2equi  # noqa: F401
