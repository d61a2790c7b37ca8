"""Drop-in for the reference's decoder/sao.py (ctu.py:2,22): the per-CTU `Sao` parameter
object, plus `filter_picture`, the SAO filter the reference never had."""
from p265_b200.sao_api import Sao, filter_picture  # noqa: F401
