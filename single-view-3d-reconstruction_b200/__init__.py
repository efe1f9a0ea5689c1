"""B200-native IF-Net implicit query path (depth map -> voxel occupancy -> sampled MLP decoder).

Importable as ``svr_b200`` (see the shim at the repository root) because this directory's name is
not a Python identifier.  Public surface mirrors the reference's ``model.projection`` and
``model.ifnet`` modules."""
from . import _abi, ops  # noqa: F401
from .model.ifnet import (IFNet, IFNetFeatureExtractor, IFNetFeatureExtractor128, configure, evaluate_network_on_grid,  # noqa: F401
                          implicit_to_mesh, make_3d_grid)
from .model.projection import project  # noqa: F401
from .prefetch import HostPrefetcher  # noqa: F401
from .graph import GraphedStep  # noqa: F401
from .mesh import export_obj, marching_cubes  # noqa: F401
from . import data  # noqa: F401

__all__ = ["IFNet", "IFNetFeatureExtractor", "IFNetFeatureExtractor128", "configure", "evaluate_network_on_grid",
           "implicit_to_mesh", "make_3d_grid", "project", "ops", "HostPrefetcher", "GraphedStep", "marching_cubes", "export_obj"]
