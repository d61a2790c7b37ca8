"""Drop-in for the reference's decoder/reconstruction.py (intra.py:6)."""
from p265_b200.residual_api import reconstruction  # noqa: F401
