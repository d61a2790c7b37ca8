"""Drop-in for the reference's decoder/sld.py (sps.py:3,18,90; pps.py:3,11,135): scaling_list_data()
syntax + ScalingFactor derivation that the reference's own module cannot run."""
from p265_b200.scaling_list import ScalingListData  # noqa: F401
