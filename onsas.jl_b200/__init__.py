"""onsas.jl_b200 -- B200-native Newton-Raphson hot path behind ONSAS.jl's API.

The directory name is the one the project layout fixes; import it as `onsas_jl_b200`
(root-level alias module).  Compute happens only in libonsas_cuda.so (hand-written sm_100a CUDA
behind the C ABI of include/onsas_cuda.h); this package is the host-side mirror of the reference's
interface for the path and the ctypes binding.  There is no CPU fallback.
"""
from . import _lib, meshgen  # noqa: F401
from ._lib import (FAMILY_TET, FAMILY_TRUSS, MAT_ISOLINEAR, MAT_NEOHOOKEAN, MAT_SVK, PRECOND_JACOBI,  # noqa: F401
                   PRECOND_NONE, PRECOND_TWO_LEVEL, STRAIN_GREEN, STRAIN_ROTATED_ENGINEERING)
from .device import DeviceContext, NativePartition, NegativeVolumeError, OnsasError, context_from_flat  # noqa: F401
from .model import (SVK, Circle, FixedField, GenericCrossSection, GlobalLoad, GreenStrain,  # noqa: F401
                    IsotropicLinearElastic, Mesh, NeoHookean, Node, Pressure, Rectangle, RotatedEngineeringStrain,
                    Square, StructuralBoundaryCondition, StructuralMaterial, Structure, Tetrahedron, TriangularFace,
                    Truss, set_dofs)
from .solve import (ConvergenceSettings, DeltaUCriterion, LinearStaticAnalysis, MaxIterCriterion,  # noqa: F401
                    NewtonRaphson, NewtonRaphsonCUDA, NonLinearStaticAnalysis, NotConvergedYet,
                    ResidualForceCriterion, ResidualsIterationStep, Solution, isconverged, solve, solve_)
from .vtk import write_vtk, write_vtu  # noqa: F401


def build(verbose: bool = False) -> str:
    """Compile the CUDA extension in-tree (nvcc, sm_100a)."""
    return _lib.build(verbose)
