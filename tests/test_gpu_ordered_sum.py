"""The parallel evaluation of the reference's left-to-right FP64 sums (surface area, area CDF of
sample_points_uniformly -- SURVEY A.10 -- and the mean / sigma of remove_statistical_outlier -- A.8)
must reproduce the scalar chain bit for bit: totals of the three sum modes and every element of the
in-place prefix (csrc/ordered_sum.cu vs its own scalar-loop kernel)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pattern", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("n", [2048, 2049, 100_000, 1_500_001])
def test_ordered_sum_equals_scalar_chain(pattern, n):
    from otslam_b200 import _lib
    bad = C.c_uint64(99)
    _lib.check(_lib.lib.otslam_selftest_ordered_sum(n, 17 * pattern + n, pattern, C.byref(bad), 0))
    assert bad.value == 0


def test_cdf_of_real_mesh_matches_oracle_counts():
    """End to end: per-triangle sample counts (exact by contract) on a mesh large enough to take
    the parallel path, against the oracle's scalar loops."""
    from oracle import oracle
    import otslam_b200.o3d_compat as o3d
    rng = np.random.default_rng(3)
    nv, nf = 20000, 60000
    verts = rng.normal(size=(nv, 3))
    faces = rng.integers(0, nv, size=(nf, 3)).astype(np.int32)
    m = o3d.geometry.TriangleMesh()
    m.vertices, m.triangles = verts, faces
    got = np.asarray(m.sample_points_uniformly(50000, seed=5).points)
    want, _, _, _ = oracle.sample_uniform(verts, None, None, faces, 50000, 5)
    assert got.shape == want.shape and (got == want).all()
