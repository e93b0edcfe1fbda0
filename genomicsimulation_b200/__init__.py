"""genomicsimulation_b200 — B200-native meiosis + genomic-evaluation engine behind
genomicSimulation's own C API for that path.

The product is two in-tree shared objects (built by `make -C genomicsimulation_b200`):
  libgsb200.so      hand-written sm_100a CUDA kernels behind the C-ABI in include/gsb200.h
  libgsb200_gsc.so  host-side C mirror of the reference's gsc_* hot-path API (include/gsb200_gsc.h)
This Python package is only the ctypes binding used by tests/ and bench.py.  There is no CPU
fallback: importing works anywhere, creating a SimData without a CUDA device raises.
"""
from ._lib import engine, host, build, LIB_ENGINE, LIB_HOST, EngineError   # noqa: F401
from .simdata import SimData, GenOptions                                    # noqa: F401
