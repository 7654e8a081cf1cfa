"""voxel_cone_tracing_b200 -- B200-native (sm_100a) voxel cone tracing hot path.

The product is the C-ABI shared library ``libvct_cuda.so`` (include/vct/vct_c.h) plus the C++
host layer that mirrors the reference's Renderer / Device / texture_3d API.  This Python
package is only the ctypes harness used by tests and bench.py.  It never falls back to a CPU
implementation: importing :mod:`voxel_cone_tracing_b200.capi` fails loudly if the CUDA
library has not been built (run ``make`` or ``python -c "import __graft_entry__ as g; g.build()"``).
"""
from . import scene  # noqa: F401

__all__ = ["scene", "capi"]
