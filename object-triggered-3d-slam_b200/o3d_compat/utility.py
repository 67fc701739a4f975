"""o3d.utility: vector containers are plain NumPy arrays (inputs are COPIED in, like Open3D)."""
import numpy as np


def Vector3dVector(a=()):
    a = np.array(a, dtype=np.float64, copy=True)
    if a.size == 0:
        return np.zeros((0, 3), np.float64)
    if a.ndim != 2 or a.shape[1] != 3:
        raise RuntimeError("Vector3dVector expects an (N, 3) array")
    return np.ascontiguousarray(a)


def Vector3iVector(a=()):
    a = np.array(a, dtype=np.int32, copy=True)
    if a.size == 0:
        return np.zeros((0, 3), np.int32)
    if a.ndim != 2 or a.shape[1] != 3:
        raise RuntimeError("Vector3iVector expects an (N, 3) array")
    return np.ascontiguousarray(a)


def random_seed(seed):
    """o3d.utility.random.seed equivalent: fixes the sampler seed used by sample_points_uniformly."""
    from . import geometry
    geometry._GLOBAL_SEED[0] = int(seed)
