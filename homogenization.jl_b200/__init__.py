"""homogenization.jl_b200 -- B200-native hot path of haampie/Homogenization.jl.

The matrix-free operator on the implicit fine grid and the geometric multigrid V-cycle as
hand-written sm_100a CUDA kernels behind a C ABI (include/hmg.h, csrc/), plus this thin host-side
mirror of the reference's Julia functions.  There is no CPU fallback: importing works without a
GPU (so that the ABI can be inspected), creating a context does not.

The directory name contains a dot, so import it through the repository's ``hmgb200`` shim
(``import hmgb200 as hmg``).
"""
from . import _lib                                            # noqa: F401
from ._lib import HmgError, LIB_PATH, PROTOTYPES, load        # noqa: F401
from .api import *                                            # noqa: F401,F403
from .api import Mesh, DeviceMatrix, LevelState, ImplicitFineGrid, BaseLevel  # noqa: F401
from . import inputs                                          # noqa: F401
from . import vtk                                             # noqa: F401
from . import driver                                          # noqa: F401
