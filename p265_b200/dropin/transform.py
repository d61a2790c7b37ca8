"""Drop-in for the reference's decoder/transform.py (intra.py:5)."""
from p265_b200.residual_api import (inverse_transform, inverse_transform_1d,  # noqa: F401
                                    set_mode)
from p265_b200.tables import trans_matrix_type0, trans_matrix_type1  # noqa: F401
