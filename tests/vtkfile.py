"""Fluid.<t>.vti helpers for the tests: the format restatement lives in oracle/fluidfiles.py (the oracle of the device-fed files)."""
from oracle.fluidfiles import read_vti as read_fluid  # noqa: F401
from oracle.fluidfiles import vti_bytes as fluid_bytes  # noqa: F401
from oracle.fluidfiles import vti_frame as frame  # noqa: F401
