"""latticednaorigami_b200 — B200-native replica-batched Monte Carlo engine for the LatticeDNAOrigami
model. The product is the CUDA library ``libldo_b200.so`` (C-ABI: include/ldo_b200.h, include/ldo_host.h)
and the ``latticeDNAOrigami_b200`` command-line driver; this package is a thin ctypes binding used by
the tests, the benchmark and Python tooling. There is no CPU fallback: importing :mod:`.binding` and
calling :func:`load` raises if the CUDA library has not been built (``make`` / ``__graft_entry__.build()``).
"""
from .binding import (  # noqa: F401
    LIB_PATH,
    Engine,
    Simulation,
    LdoError,
    load,
    DRAW_DTYPE,
)
