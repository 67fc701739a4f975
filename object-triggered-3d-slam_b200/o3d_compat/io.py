"""o3d.io: image reading and binary PLY reading/writing in Open3D's on-disk layout
(SURVEY Appendix C).  Point-cloud vertex records (27 B: 3 x f64 + 3 x u8) are packed on the GPU."""
import ctypes as C
import os

import numpy as np

from .. import _lib
from . import geometry

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "<i2", "int16": "<i2",
              "ushort": "<u2", "uint16": "<u2", "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4",
              "float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8"}


def read_image(filename):
    """o3d.io.read_image (reconstruct_rgbd.py:88-89): RGB order, native bit depth.  A missing or
    unreadable file yields an EMPTY image plus a warning -- the failure surfaces at integrate()."""
    import cv2
    a = cv2.imread(filename, cv2.IMREAD_UNCHANGED) if os.path.exists(filename) else None
    if a is None:
        print(f"[Open3D WARNING] Read image failed: unable to open file: {filename}")
        return geometry.Image()
    if a.ndim == 3:             # OpenCV decodes to BGR(A); Open3D hands out RGB.  cvtColor is ~10x a strided NumPy copy
        a = cv2.cvtColor(a, cv2.COLOR_BGRA2RGB if a.shape[2] == 4 else cv2.COLOR_BGR2RGB) if a.shape[2] in (3, 4) else a[..., :3][..., ::-1]
    return geometry.Image(np.ascontiguousarray(a))


def _header(kind_lines):
    return ("ply\nformat binary_little_endian 1.0\ncomment Created by Open3D\n" + "".join(kind_lines) + "end_header\n").encode()


def write_point_cloud(filename, pointcloud, write_ascii=False, compressed=False, print_progress=False):
    """o3d.io.write_point_cloud (reconstruct_rgbd_filter.py:140, fusion/hybrid_map.py:121)."""
    if str(filename).lower().endswith(".pcd"):
        return _write_pcd(filename, pointcloud, write_ascii)
    if write_ascii:
        raise RuntimeError("ASCII PLY output is not supported (the reference writes binary)")
    if not str(filename).lower().endswith(".ply"):
        raise RuntimeError("Write geometry::PointCloud failed: unknown file extension (.ply and .pcd are supported)")
    p = pointcloud.points
    n = len(p)
    if n == 0:
        print("[Open3D WARNING] Write PLY failed: point cloud has 0 points.")
        return False
    lines = [f"element vertex {n}\n", "property double x\nproperty double y\nproperty double z\n"]
    if pointcloud.has_normals():
        lines.append("property double nx\nproperty double ny\nproperty double nz\n")
    if pointcloud.has_colors():
        lines.append("property uchar red\nproperty uchar green\nproperty uchar blue\n")
    if pointcloud.has_colors() and not pointcloud.has_normals():
        rec = pack_cloud_records([p], [pointcloud.colors])           # GPU: 27-byte records
    else:
        dt = [("p", "<f8", 3)] + ([("n", "<f8", 3)] if pointcloud.has_normals() else []) + \
             ([("c", "u1", 3)] if pointcloud.has_colors() else [])
        rec = np.zeros(n, np.dtype(dt))
        rec["p"] = p
        if pointcloud.has_normals():
            rec["n"] = pointcloud.normals
        if pointcloud.has_colors():
            rec["c"] = _color_bytes(pointcloud.colors)
    with open(filename, "wb") as f:
        f.write(_header(lines))
        f.write(rec.tobytes())
    return True


def _write_pcd(filename, pointcloud, write_ascii=False):
    """PCD v0.7 as Open3D's WritePointCloudToPCD lays it out (north_star names .pcd beside .ply; the reference itself
    only writes .ply): float32 x y z, optional float32 normal_x normal_y normal_z, optional `rgb` = one float32 whose BITS
    are r << 16 | g << 8 | b, DATA binary (or ascii with write_ascii=True, rgb printed as the same float)."""
    p = np.asarray(pointcloud.points, np.float64)
    n = len(p)
    if n == 0:
        print("[Open3D WARNING] Write PCD failed: point cloud has 0 points.")
        return False
    fields, dt = ["x", "y", "z"], [("p", "<f4", 3)]
    if pointcloud.has_normals():
        fields += ["normal_x", "normal_y", "normal_z"]
        dt.append(("n", "<f4", 3))
    if pointcloud.has_colors():
        fields.append("rgb")
        dt.append(("c", "<u4"))
    k = len(fields)
    header = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n"
              f"FIELDS {' '.join(fields)}\nSIZE {' '.join(['4'] * k)}\nTYPE {' '.join(['F'] * k)}\nCOUNT {' '.join(['1'] * k)}\n"
              f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {'ascii' if write_ascii else 'binary'}\n")
    rec = np.empty(n, np.dtype(dt))
    rec["p"] = p
    if pointcloud.has_normals():
        rec["n"] = pointcloud.normals
    if pointcloud.has_colors():
        c = _color_bytes(pointcloud.colors).astype(np.uint32)
        rec["c"] = (c[:, 0] << 16) | (c[:, 1] << 8) | c[:, 2]
    with open(filename, "wb") as f:
        f.write(header.encode())
        if write_ascii:
            cols = [rec["p"]] + ([rec["n"]] if pointcloud.has_normals() else []) + \
                   ([rec["c"].view(np.float32)[:, None]] if pointcloud.has_colors() else [])
            np.savetxt(f, np.concatenate(cols, axis=1), fmt="%.10g")
        else:
            rec.tofile(f)
    return True


def _read_pcd(filename):
    with open(filename, "rb") as f:
        data = f.read()
    hdr, off = {}, 0
    while True:
        nl = data.find(b"\n", off)
        if nl < 0:
            raise RuntimeError("Read PCD failed: no DATA line")
        line = data[off:nl].decode("ascii", "replace").strip()
        off = nl + 1
        if not line or line.startswith("#"):
            continue
        t = line.split()
        hdr[t[0].upper()] = t[1:]
        if t[0].upper() == "DATA":
            break
    fields, sizes, types = hdr["FIELDS"], [int(x) for x in hdr["SIZE"]], hdr["TYPE"]
    counts = [int(x) for x in hdr.get("COUNT", ["1"] * len(fields))]
    n = int(hdr["POINTS"][0])
    np_t = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 1): "u1", ("U", 2): "<u2", ("U", 4): "<u4", ("I", 1): "i1", ("I", 2): "<i2", ("I", 4): "<i4"}
    dt = np.dtype([(f, np_t[(t, s)], c) if c > 1 else (f, np_t[(t, s)]) for f, s, t, c in zip(fields, sizes, types, counts)])
    kind = hdr["DATA"][0].lower()
    if kind == "binary":
        a = np.frombuffer(data, dt, n, off)
    elif kind == "ascii":
        tok = np.array(data[off:].split(), dtype=np.float64).reshape(n, -1)
        a = np.empty(n, dt)
        for i, f in enumerate(fields):
            a[f] = tok[:, i].astype(np.float32).view(np.uint32) if (f in ("rgb", "rgba") and dt[f] == np.dtype("<u4")) else tok[:, i]
    else:
        raise RuntimeError(f"Read PCD failed: DATA {kind} is not supported")
    pts = np.stack([a["x"], a["y"], a["z"]], 1).astype(np.float64)
    nrm = np.stack([a["normal_x"], a["normal_y"], a["normal_z"]], 1).astype(np.float64) if "normal_x" in fields else None
    col = None
    if "rgb" in fields or "rgba" in fields:
        bits = np.ascontiguousarray(a["rgb" if "rgb" in fields else "rgba"]).view(np.uint32)
        col = np.stack([(bits >> 16) & 255, (bits >> 8) & 255, bits & 255], 1).astype(np.float64) / 255.0
    return pts, nrm, col


def _color_bytes(c):
    return np.floor(np.clip(c, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)


def pack_cloud_records(points_list, colors_list=None, paint=None):
    """Concatenate clouds and pack 27-byte PLY vertex records on the GPU
    (paint_uniform_color + `+=` + write of fusion/hybrid_map.py:59,88-91,115,121)."""
    pts = [np.ascontiguousarray(p, np.float64) for p in points_list]
    counts = np.array([len(p) for p in pts], np.int64)
    total = int(counts.sum())
    out = np.empty((total, 27), np.uint8)
    if total == 0:
        return out
    k = len(pts)
    pp = (C.c_void_p * k)(*[p.ctypes.data for p in pts])
    cp, cols, pa = None, None, None
    if paint is not None:
        pa = np.ascontiguousarray(paint, np.float64).reshape(k, 3)
    else:
        cols = [np.ascontiguousarray(c, np.float64) for c in colors_list]
        cp = (C.c_void_p * k)(*[c.ctypes.data for c in cols])
    _lib.check(_lib.lib.otslam_cloud_merge_pack(k, pp, cp, _lib.ptr(counts), _lib.ptr(pa), _lib.ptr(out), 0))
    return out


def write_cloud_records(filename, records):
    """Write pre-packed 27-byte records (xyz f64 + rgb u8) as an Open3D-layout PLY."""
    n = len(records)
    lines = [f"element vertex {n}\n", "property double x\nproperty double y\nproperty double z\n",
             "property uchar red\nproperty uchar green\nproperty uchar blue\n"]
    with open(filename, "wb") as f:
        f.write(_header(lines))
        f.write(np.ascontiguousarray(records).tobytes())
    return True


def write_triangle_mesh(filename, mesh, write_ascii=False, compressed=False, write_vertex_normals=True,
                        write_vertex_colors=True, write_triangle_uvs=True, print_progress=False):
    """o3d.io.write_triangle_mesh (reconstruct_rgbd.py:118): 51-byte vertices + 13-byte faces."""
    if write_ascii:
        raise RuntimeError("ASCII PLY output is not supported (the reference writes binary)")
    v = mesh.vertices
    n, nf = len(v), len(mesh.triangles)
    if n == 0:
        print("[Open3D WARNING] Write PLY failed: mesh has 0 vertices.")
        return False
    wn = write_vertex_normals and mesh.has_vertex_normals()
    wc = write_vertex_colors and mesh.has_vertex_colors()
    lines = [f"element vertex {n}\n", "property double x\nproperty double y\nproperty double z\n"]
    dt = [("p", "<f8", 3)]
    if wn:
        lines.append("property double nx\nproperty double ny\nproperty double nz\n")
        dt.append(("n", "<f8", 3))
    if wc:
        lines.append("property uchar red\nproperty uchar green\nproperty uchar blue\n")
        dt.append(("c", "u1", 3))
    lines.append(f"element face {nf}\nproperty list uchar uint vertex_indices\n")
    rec = np.empty(n, np.dtype(dt))              # packed records: every byte is assigned below
    rec["p"] = v
    if wn:
        rec["n"] = mesh.vertex_normals
    if wc:
        rec["c"] = _color_bytes(mesh.vertex_colors)
    fr = np.empty(nf, np.dtype([("k", "u1"), ("i", "<u4", 3)]))
    fr["k"] = 3
    fr["i"] = mesh.triangles
    with open(filename, "wb") as f:
        f.write(_header(lines))
        rec.tofile(f)                            # straight from the array: no tobytes() copy of a ~50 MB buffer
        fr.tofile(f)
    return True


def _read_ply(filename):
    with open(filename, "rb") as f:
        data = f.read()
    end = data.find(b"end_header")
    if not data.startswith(b"ply") or end < 0:
        raise RuntimeError(f"Read PLY failed: {filename} is not a PLY file")
    nl = data.index(b"\n", end) + 1
    header = data[:nl].decode("ascii", "replace").splitlines()
    fmt, elements = None, []
    for line in header:
        t = line.split()
        if not t:
            continue
        if t[0] == "format":
            fmt = t[1]
        elif t[0] == "element":
            elements.append((t[1], int(t[2]), []))
        elif t[0] == "property":
            elements[-1][2].append(t[1:])
    out, off = {}, nl
    if fmt == "ascii":
        tokens = data[nl:].split()
        pos = 0
        for name, count, props in elements:
            if any(p[0] == "list" for p in props):
                rows = []
                for _ in range(count):
                    k = int(tokens[pos]); rows.append([int(x) for x in tokens[pos + 1:pos + 1 + k]]); pos += 1 + k
                out[name] = {"list": rows}
            else:
                arr = np.array(tokens[pos:pos + count * len(props)], dtype=np.float64).reshape(count, len(props))
                pos += count * len(props)
                out[name] = {p[1]: arr[:, i] for i, p in enumerate(props)}
        return out
    if fmt != "binary_little_endian":
        raise RuntimeError(f"Read PLY failed: unsupported format {fmt}")
    for name, count, props in elements:
        if any(p[0] == "list" for p in props):
            if len(props) != 1:
                raise RuntimeError("Read PLY failed: unsupported face layout")
            ct, it = _PLY_TYPES[props[0][1]], _PLY_TYPES[props[0][2]]
            if count == 0:
                out[name] = {"list": np.zeros((0, 3), np.int32)}
                continue
            k = int(np.frombuffer(data, ct, 1, off)[0])
            dt = np.dtype([("k", ct), ("i", it, k)])
            a = np.frombuffer(data, dt, count, off)
            if not (a["k"] == k).all():
                raise RuntimeError("Read PLY failed: only fixed-size faces are supported")
            out[name] = {"list": a["i"].astype(np.int32)}
            off += dt.itemsize * count
        else:
            dt = np.dtype([(p[1], _PLY_TYPES[p[0]]) for p in props])
            a = np.frombuffer(data, dt, count, off)
            out[name] = {p[1]: a[p[1]] for p in props}
            off += dt.itemsize * count
    return out


def _vertex_arrays(ply):
    v = ply.get("vertex", {})
    if not all(k in v for k in ("x", "y", "z")):
        return np.zeros((0, 3)), None, None
    pts = np.stack([v["x"], v["y"], v["z"]], 1).astype(np.float64)
    nrm = np.stack([v["nx"], v["ny"], v["nz"]], 1).astype(np.float64) if all(k in v for k in ("nx", "ny", "nz")) else None
    col = None
    if all(k in v for k in ("red", "green", "blue")):
        col = np.stack([v["red"], v["green"], v["blue"]], 1).astype(np.float64) / 255.0
    return pts, nrm, col


def read_point_cloud(filename, format="auto", remove_nan_points=False, remove_infinite_points=False, print_progress=False):
    """o3d.io.read_point_cloud (fusion/hybrid_map.py:79): a mesh PLY yields its vertices; failures
    give an empty cloud plus a warning."""
    pc = geometry.PointCloud()
    try:
        pts, nrm, col = _read_pcd(filename) if str(filename).lower().endswith(".pcd") else _vertex_arrays(_read_ply(filename))
    except Exception as e:  # noqa: BLE001
        print(f"[Open3D WARNING] Read {'PCD' if str(filename).lower().endswith('.pcd') else 'PLY'} failed: {e}")
        return pc
    pc.points = pts
    if nrm is not None:
        pc.normals = nrm
    if col is not None:
        pc.colors = col
    return pc


def read_triangle_mesh(filename, enable_post_processing=False, print_progress=False):
    """o3d.io.read_triangle_mesh (fusion/hybrid_map.py:83)."""
    m = geometry.TriangleMesh()
    try:
        ply = _read_ply(filename)
    except Exception as e:  # noqa: BLE001
        print(f"[Open3D WARNING] Read PLY failed: {e}")
        return m
    pts, nrm, col = _vertex_arrays(ply)
    m.vertices = pts
    if nrm is not None:
        m.vertex_normals = nrm
    if col is not None:
        m.vertex_colors = col
    if "face" in ply:
        m.triangles = np.asarray(ply["face"]["list"], np.int32).reshape(-1, 3)
    return m
