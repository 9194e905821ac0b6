"""breeze.jl_b200 — B200-native drop-in for Breeze.jl's anelastic SSP-RK3 tendency + pressure-correction path.

Import as `breeze_b200` (repo-root shim; a directory name containing a dot cannot be imported directly).
Everything numerical runs in csrc/libbreeze_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/breeze_b200.h). There is no CPU fallback: constructing a model without the built library raises.
"""
from . import cases
from .abi import BreezeError, Context, Library, bz_config, bz_forcing, load_cuda_library, cuda_library_path, FIELD_IDS
from .model import (B200, DragFluxBoundaryCondition, FluxBoundaryCondition, Forcing, FPlane, GeostrophicForcing, SubsidenceForcing,
                    geostrophic_forcings, AnelasticDynamics, AtmosphereModel, Bounded, Flat, Periodic, RectilinearGrid, ReferenceState,
                    SaturationAdjustment, Simulation, ThermodynamicConstants, TimeStepWizard, NaNChecker, WENO,
                    conjure_time_step_wizard_, enable_peer_memory, many_time_steps_, run_, set_, time_step_)

from .compressible import (CompressibleAtmosphereModel, CompressibleContext, CompressibleDynamics, ConstantSubstepSize, MonolithicFirstStage,
                           NoDivergenceDamping, ProportionalSubsteps, SplitExplicitTimeDiscretization, ThermalDivergenceDamping,
                           UpperSponge, LinearRamp, CubicRamp, Sin2Ramp,
                           bzc_config, compressible_library)

__all__ = [n for n in dir() if not n.startswith("_")]
