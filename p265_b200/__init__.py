"""p265_b200 -- B200 (sm_100a) residual + SAO path of the p265 H.265 decoder.

Public surface:
  p265_b200.engine.Engine          C-ABI context (residual / SAO launches)
  p265_b200.residual_api           inverse_scaling / inverse_transform (+ flush_picture)
  p265_b200.sao_api                Sao parameters, filter_picture
  p265_b200.dropin_path()          directory of the `scaling`, `transform`, `sao` modules
                                   that replace the reference's decoder/*.py by name
"""
import os

__version__ = "0.1.0"


def dropin_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")
