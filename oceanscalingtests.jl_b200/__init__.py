"""ocean-b200: B200-native `time_step!` of the OceanScalingTests.jl hydrostatic model.

Only what the hot path needs: csrc/ (CUDA kernels + the C ABI, libocean_b200.so) and a thin
host-side mirror of the reference's Julia interface (same names and argument meaning).
The directory name contains a dot, so import it through `__graft_entry__.package()` (or
tests/conftest.py), which registers it as the module `ocean_b200`.
"""
from . import _abi, synthetic
from .bathymetry_and_forcings import (T_reference, cubic_profile, double_drake_bathymetry, exponential_profile,
                                      exponential_z_faces, linear_z_faces, salinity_flux,
                                      salinity_flux_coefficients, wind_stress, wind_stress_coefficients)
from .boundary_and_initial_conditions import (flux_filenum, initialize_model, load_fluxes, load_restoring,
                                              partition_global_array, restoring_filenum, set_boundary_conditions,
                                              update_fluxes, update_restoring)
from .grid_load_balance import (LatitudeLongitudeGrid, Partition, assess_x_load, assess_x_load_from_bottom, assess_y_load, calculate_local_N,
                                kernel_cost_per_column, load_balanced_grid, redistribute_size_to_fulfill_memory_limitation)
from .model import (B200Backend, BoundaryConditions, HydrostaticFreeSurfaceModel, QGLeith,
                    RiBasedVerticalDiffusivity, SplitExplicitFreeSurface, VerticalScalarDiffusivity,
                    barotropic_substeps, build_config, initialize_free_surface, set_model, substeps, time_step, weights_from_substeps)
from .near_global_simulation import (Callback, Comm, Simulation, aligned_time_step, equation_of_state, experiment_depth,
                                     initialize_simulation, profiled_time_steps, run, scaling_test_simulation,
                                     time_step_wizard)
from .simulation_outputs import (Checkpointer, GroupFile, IterationInterval, SurfaceWriter, TimeInterval, set_from_checkpoint,
                                 set_outputs, set_time_stepper_tendencies, load_free_surface,
                                 vertical_vorticity_surface)

__all__ = [n for n in dir() if not n.startswith("_")]
