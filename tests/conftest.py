import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if not has_gpu:
        skip = pytest.mark.skip(reason="no CUDA device in this container (the product has no CPU fallback)")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the oracle and the CUDA library exist (compiles on a CPU box too)."""
    from oracle import oracle
    oracle.build()
    so = os.path.join(ROOT, "object-triggered-3d-slam_b200", "libotslam_b200.so")
    if not os.path.exists(so):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def table_seq():
    """6 frames of the config-1 table trajectory (every 50th of 300), 640x480."""
    from otslam_b200 import synth
    seq = synth.make_sequence("table", 300, subsample=(0, 50))
    d, c = seq.numpy()
    return seq, d, c


@pytest.fixture(scope="session")
def small_seq():
    """4 frames, 160x120 (intrinsics / 4): fast enough for pure-CPU oracle tests."""
    from otslam_b200 import synth
    intr = (160, 120, 565.6009 / 4, 565.6009 / 4, 80.5, 60.5)
    seq = synth.make_sequence("chair_table", 40, intr=intr, subsample=(0, 10))
    d, c = seq.numpy()
    return seq, d, c


def lexorder(k):
    k = np.asarray(k)
    return np.lexsort(tuple(k[:, i] for i in range(k.shape[1] - 1, -1, -1)))


def canon_mesh(verts, cols, faces, ek):
    """Canonical mesh: vertices sorted by edge key, faces rotated to start at their smallest index and sorted."""
    order = lexorder(ek)
    inv = np.empty_like(order)
    inv[order] = np.arange(len(order))
    f = inv[np.asarray(faces, np.int64)]
    if len(f):
        r = np.argmin(f, axis=1)
        f = np.stack([np.take_along_axis(f, ((r + k) % 3)[:, None], 1)[:, 0] for k in range(3)], 1)
        f = f[np.lexsort((f[:, 2], f[:, 1], f[:, 0]))]
    return verts[order], cols[order], f, ek[order]
