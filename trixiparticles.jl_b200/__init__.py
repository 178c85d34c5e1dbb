"""trixiparticles.jl_b200 -- B200-native WCSPH right-hand side behind TrixiParticles.jl's
`Semidiscretization` / `semidiscretize` / `kick!` / `drift!` interface.

Only what the hot path needs lives here: the CUDA library (csrc/, built to libtpb200.so),
its ctypes binding (_lib.py) and the host-side mirror of the reference interface.
Import it as `import trixiparticles.jl_b200` (the `trixiparticles/` shim next to this
directory maps the dotted name onto this folder).
"""
from .model import (AdamiPressureExtrapolation, ArtificialViscosityMonaghan, BernoulliPressureExtrapolation,
                    BoundaryModelDummyParticles, ContinuityDensity,
                    DensityDiffusionMolteniColagrossi, SchoenbergCubicSplineKernel,
                    SchoenbergQuarticSplineKernel, SchoenbergQuinticSplineKernel,
                    SourceTermDamping, StateEquationAdaptiveCole, StateEquationCole, SummationDensity,
                    ViscosityAdami, ViscosityMorris,
                    WallBoundarySystem, TotalLagrangianSPHSystem, BoundaryModelMonaghanKajtar,
                    PenaltyForceGanzenmueller, PrescribedMotion, OscillatingMotion2D,
                    WeaklyCompressibleSPHSystem, WendlandC2Kernel, WendlandC4Kernel, WendlandC6Kernel,
                    compact_support)
from .semidiscretization import (B200Backend, DynamicalODEProblem, FullGridCellList,
                                 GridNeighborhoodSearch, Semidiscretization, drift_, kick_,
                                 semidiscretize)
from .interpolation import interpolate_line, interpolate_points
from .setups import InitialCondition, RectangularShape, RectangularTank, SphereShape, reset_wall_, union

__all__ = [
    "AdamiPressureExtrapolation", "ArtificialViscosityMonaghan", "BernoulliPressureExtrapolation",
    "BoundaryModelDummyParticles",
    "ContinuityDensity", "DensityDiffusionMolteniColagrossi", "SchoenbergCubicSplineKernel",
    "SchoenbergQuarticSplineKernel", "SchoenbergQuinticSplineKernel",
    "SourceTermDamping", "StateEquationAdaptiveCole", "StateEquationCole", "SummationDensity",
    "ViscosityAdami", "ViscosityMorris",
    "WallBoundarySystem", "TotalLagrangianSPHSystem", "BoundaryModelMonaghanKajtar",
    "PenaltyForceGanzenmueller", "PrescribedMotion", "OscillatingMotion2D",
    "WeaklyCompressibleSPHSystem", "WendlandC2Kernel", "WendlandC4Kernel", "WendlandC6Kernel",
    "compact_support", "B200Backend",
    "DynamicalODEProblem", "FullGridCellList", "GridNeighborhoodSearch", "Semidiscretization",
    "drift_", "kick_", "semidiscretize", "InitialCondition", "RectangularShape",
    "RectangularTank", "SphereShape", "reset_wall_", "union", "interpolate_line", "interpolate_points",
]
