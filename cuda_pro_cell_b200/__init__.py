"""cuda_pro_cell_b200 - B200-native ProCell proliferation simulator (host-side mirror of the reference interface).

The compute path is libprocell_b200.so (hand-written sm_100a CUDA behind a C ABI, include/procell_b200.h).
This package only binds it; there is no Python or CPU implementation of the simulation, and importing
`cuda_pro_cell_b200.api` fails loudly when the shared library has not been built.
"""
from . import synth  # noqa: F401

__all__ = ["synth", "api"]
