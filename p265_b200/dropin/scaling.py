"""Drop-in for the reference's decoder/scaling.py: put `p265_b200.dropin_path()` ahead
of the decoder directory on sys.path and `import scaling` (intra.py:4) resolves here."""
from p265_b200.residual_api import inverse_scaling  # noqa: F401
