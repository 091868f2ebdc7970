"""dc_sand_b200 -- B200-native (sm_100a) drop-in for the feng/ddc digital down-converter of ska-sa/dc_sand.

Public surface (mirrors /root/reference/feng/ddc/src):
    dc_sand_b200.ddc.DigitalDownConverter   (ddc.py:10-188)  -- run() executes on the GPU through libddcb200.so
    dc_sand_b200.cwg.generate_carrier_wave  (cwg.py:6-44)    -- host test-vector generator
    dc_sand_b200.stream.DDCStream           (extension)      -- the same operator on an endless stream, pushed in pieces
There is no CPU fallback: importing the operator without the CUDA library raises.
"""
from . import cwg, ddc  # noqa: F401
from .ddc import DigitalDownConverter  # noqa: F401
from .stream import DDCStream  # noqa: F401

__version__ = "0.1.0"
