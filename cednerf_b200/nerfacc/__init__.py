"""Drop-in for the subset of `nerfacc` the reference imports (cednerf/utils.py:12-18, cednerf/render.py:6,
train_real.py:27): same names, argument meaning and return shapes, backed by libcednerf_b200.so."""
from .grid import RayIntervals, RaySamples, ray_aabb_intersect, traverse_grids
from .volrend import (accumulate_along_rays, accumulate_along_rays_, render_transmittance_from_density,
                      render_visibility_from_density, render_weight_from_density)
from .estimators.occ_grid import OccGridEstimator
from . import estimators, grid, volrend

__all__ = ["RayIntervals", "RaySamples", "ray_aabb_intersect", "traverse_grids", "accumulate_along_rays",
           "accumulate_along_rays_", "render_transmittance_from_density", "render_visibility_from_density",
           "render_weight_from_density", "OccGridEstimator"]
