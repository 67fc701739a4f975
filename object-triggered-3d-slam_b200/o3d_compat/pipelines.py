"""o3d.pipelines.integration: ScalableTSDFVolume on the GPU (reconstruct_rgbd.py:79-83,107,112)."""
import enum
import types

import numpy as np

from .. import _lib
from ..volume import TSDFVolume
from . import geometry

_FMT = "[ScalableTSDFVolume::Integrate] Unsupported image format."


class TSDFVolumeColorType(enum.Enum):
    NoColor = 0
    RGB8 = 1
    Gray32 = 2


class ScalableTSDFVolume:
    def __init__(self, voxel_length, sdf_trunc, color_type, volume_unit_resolution=16, depth_sampling_stride=4,
                 device=0, slab=None):
        if volume_unit_resolution != 16 or depth_sampling_stride != 4:
            raise RuntimeError("otslam_b200 supports volume_unit_resolution=16, depth_sampling_stride=4 (the Open3D "
                               "defaults the reference relies on)")
        if color_type == TSDFVolumeColorType.Gray32:
            raise RuntimeError("TSDFVolumeColorType.Gray32 is not supported (the reference uses RGB8)")
        self.voxel_length, self.sdf_trunc, self.color_type = float(voxel_length), float(sdf_trunc), color_type
        self._vol = TSDFVolume(voxel_length, sdf_trunc, color=(color_type == TSDFVolumeColorType.RGB8), device=device,
                               slab=slab)

    # -- reference API -------------------------------------------------------------------------
    def reset(self):
        self._vol.reset()

    def integrate(self, image, intrinsic, extrinsic):
        """volume.integrate(rgbd, intrinsics, extrinsic) -- reconstruct_rgbd.py:107."""
        depth_raw, color = image._depth_raw, image._color_arr
        if depth_raw is None or depth_raw.ndim != 2:
            raise RuntimeError(_FMT)
        H, W = depth_raw.shape
        if (W, H) != (intrinsic.width, intrinsic.height):
            raise RuntimeError(_FMT)
        rgb = None
        if self.color_type == TSDFVolumeColorType.RGB8:
            if color is None or color.dtype != np.uint8 or color.shape != (H, W, 3):
                raise RuntimeError(_FMT)
            rgb = color
        ext = np.asarray(extrinsic, np.float64)
        if ext.shape != (4, 4):
            raise RuntimeError("extrinsic must be a 4x4 matrix")
        if image._depth_is_raw_u16:
            self._vol.integrate_u16(depth_raw, rgb, intrinsic.fxfycxcy(), ext, image._depth_scale, image._depth_trunc)
        elif depth_raw.dtype == np.float32:
            self._vol.integrate_f32(depth_raw, rgb, intrinsic.fxfycxcy(), ext)
        else:
            raise RuntimeError(_FMT)

    def extract_triangle_mesh(self):
        """reconstruct_rgbd.py:112.  The mesh (often ~100 MB) stays in HBM: the returned TriangleMesh
        downloads its arrays only when they are read (np.asarray(mesh.vertices), write_triangle_mesh);
        len(mesh.vertices), compute_vertex_normals() and sample_points_uniformly() work on the device."""
        m = geometry.TriangleMesh()
        nv, nf = self._vol.extract_mesh_resident(owner=m)
        m._attach_resident(self._vol, nv, nf, colors=self.color_type != TSDFVolumeColorType.NoColor)
        return m

    def extract_point_cloud(self):
        """ScalableTSDFVolume.extract_point_cloud(): zero crossings along +x/+y/+z with colours and TSDF-gradient normals."""
        p, c, _, n = self._vol.extract_point_cloud(normals=True)
        pc = geometry.PointCloud()
        pc.points = p
        pc.normals = n
        if self.color_type != TSDFVolumeColorType.NoColor:
            pc.colors = c
        return pc

    # -- batched extension: the whole frame loop in one call (frame order preserved) ------------
    def integrate_sequence(self, depths, colors, intrinsic, extrinsics, depth_scale=1000.0, depth_trunc=3.0):
        n = len(extrinsics)
        if n == 0:
            return
        if tuple(depths.shape[1:]) != (intrinsic.height, intrinsic.width):
            raise RuntimeError(_FMT)
        self._vol.integrate_batch(depths, colors if self.color_type == TSDFVolumeColorType.RGB8 else None,
                                  intrinsic.fxfycxcy(), extrinsics, depth_scale, depth_trunc)


integration = types.SimpleNamespace(ScalableTSDFVolume=ScalableTSDFVolume, TSDFVolumeColorType=TSDFVolumeColorType)
