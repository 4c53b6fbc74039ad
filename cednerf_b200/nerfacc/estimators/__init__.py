from .occ_grid import OccGridEstimator

__all__ = ["OccGridEstimator"]
