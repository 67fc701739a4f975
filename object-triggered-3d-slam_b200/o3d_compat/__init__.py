"""Drop-in for the subset of the `open3d` Python API that the reference's hot-path scripts use
(call sites /root/reference/3d_model/reconstruct_rgbd.py:4,25,79-118,
reconstruct_rgbd_filter.py:112-140, multi_reconstruct_rgbd_filter.py:57-137,
fusion/hybrid_map.py:57-59,73-91,115-129, 3d_model/check_one_frame.py:20-30):

    import otslam_b200.o3d_compat as o3d

Every heavy operation goes to the sm_100a CUDA library through the C ABI (otslam_b200._lib);
there is no CPU fallback.
"""
from . import camera, geometry, io, pipelines, utility, visualization  # noqa: F401

__version__ = "otslam_b200-compat"
