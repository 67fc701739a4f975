"""otslam_b200 -- B200-native RGB-D object-reconstruction hot path of
TakiRyo/object-triggered-3D-SLAM (the frame loop of 3d_model/reconstruct_rgbd*.py and the cloud
merge of fusion/hybrid_map.py) behind the subset of the open3d Python API those scripts use.

    import otslam_b200.o3d_compat as o3d      # drop-in for `import open3d as o3d`

Hand-written sm_100a CUDA kernels behind a C ABI (include/otslam_b200.h); no CPU fallback.
"""
__version__ = "0.1.0"
