"""o3d.camera.PinholeCameraIntrinsic (reconstruct_rgbd.py:22-25)."""
import numpy as np


class PinholeCameraIntrinsic:
    def __init__(self, width=-1, height=-1, fx=0.0, fy=0.0, cx=0.0, cy=0.0):
        self.width, self.height = int(width), int(height)
        self.intrinsic_matrix = np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], np.float64)

    def set_intrinsics(self, width, height, fx, fy, cx, cy):
        self.__init__(width, height, fx, fy, cx, cy)

    def get_focal_length(self):
        return float(self.intrinsic_matrix[0, 0]), float(self.intrinsic_matrix[1, 1])

    def get_principal_point(self):
        return float(self.intrinsic_matrix[0, 2]), float(self.intrinsic_matrix[1, 2])

    def fxfycxcy(self):
        m = self.intrinsic_matrix
        return float(m[0, 0]), float(m[1, 1]), float(m[0, 2]), float(m[1, 2])

    def is_valid(self):
        return self.width > 0 and self.height > 0

    def __repr__(self):
        return f"PinholeCameraIntrinsic with width = {self.width} and height = {self.height}."
