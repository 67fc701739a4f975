#!/usr/bin/env python
"""bench.py -- RGB-D frames/s integrated (640x480, 5 mm TSDF) on B200.

Default workload = BASELINE.json configs[1]: the frame loop of reconstruct_rgbd_filter.py
(/root/reference/3d_model/reconstruct_rgbd_filter.py:86-111) over a 1000-frame synthetic chair+table
sequence (two camera rings), 640x480, 5 mm voxels.  One "step" = that loop from a reset volume:
    reset -> [depth convert + depth-range mask, block allocation, TSDF+colour integration] x n_frames.
The script's post stage (extract mesh -> sample 100 000 -> z mask -> voxel_down_sample ->
remove_statistical_outlier) is timed once after the loop and reported under "post_stage" with the
filter kernels' own HBM figures; it is not part of `value` (the metric is frames integrated per s).
  value  : frames/s with the sequence already resident in HBM (kernel path only)
  e2e    : same metric through the public C-ABI call with HOST (pinned) buffers, H2D copies and a
           D2H read of the result statistics inside the timed region
  e2e_full: the whole job, host frames -> final cloud in host memory (frame loop, extraction, sampling / gather, download)
  roofline: integrate_kernel, algorithmic bytes (40 B x N_upd + 5 B x W x H per frame)
           / CUDA-event duration of that kernel, against MEASURED_PEAKS.json; bound "issue" 
  cpu_baseline: the oracle (CPU port of the Open3D algorithm the reference calls; real open3d when it imports) on a
           bounded sample of the same sequence, all host threads; full_loop = incl. JPEG / PNG decode, loadtxt, inv
  hd     : the same measurement on BASELINE configs[3]'s geometry (1280x720 x --hd-frames frames, 2 mm voxels), so that the
           1/2/4/8-GPU runs record north_star's scaling config in the same JSON line

Multi-GPU (torchrun, one rank per GPU): every rank receives every frame and integrates only its
own spatial slab -- total work fixed => "scaling": "strong"; no per-frame collective.
`--impl reference` times the reference's CPU path (the oracle port; Open3D itself is not
installable here) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "RGB-D frames/s integrated (640x480, 5mm TSDF)"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--scene", default="chair_table", help="chair_table = configs[1]; table with --frames 300 = configs[0]")
    ap.add_argument("--voxel", type=float, default=0.005)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--hd", action="store_true", help="1280x720 / 2 mm large-room config (SURVEY 8d config 4)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="oracle sample size (0 = auto)")
    ap.add_argument("--slab-thickness", type=int, default=1, help="multi-GPU: blocks per slab (cyclic over ranks)")
    ap.add_argument("--slab-halo", type=int, default=0, help="multi-GPU: 1 = replicate +1 halo blocks, 0 = exchange planes")
    ap.add_argument("--emulate-world", type=int, default=0, help="dev: single GPU integrating only rank 0's slabs of an N-rank run")
    ap.add_argument("--emulate-rank", type=int, default=0, help="dev: which rank's slabs --emulate-world integrates")
    ap.add_argument("--slab-axis", type=int, default=3, help="multi-GPU: 0/1/2 = x/y/z slabs, 3 = diagonal slabs (ownership by kx + ky; halo 0 only)")
    ap.add_argument("--zsplit", type=int, default=0, help="dev: CTAs per block along z in the integration kernel")
    ap.add_argument("--hd-frames", type=int, default=2000, help="frames of the 1280x720 / 2 mm sub-run reported under \"hd\" (0 = skip)")
    ap.add_argument("--hd-steps", type=int, default=2)
    ap.add_argument("--hd-voxel", type=float, default=0.002)
    ap.add_argument("--ingest-chunk", type=int, default=128, help="multi-GPU host ingest: frames per all-gathered chunk")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-post", action="store_true")
    return ap.parse_args()


def workload_name(a):
    if a.hd:
        return f"large-room synthetic sequence 1280x720 x {a.frames} frames, voxel {a.voxel*1000:g} mm / trunc {4*a.voxel*1000:g} mm"
    script = "reconstruct_rgbd_filter.py (configs[1])" if a.scene == "chair_table" else "reconstruct_rgbd.py (configs[0])"
    return (f"{script} frame loop: {a.frames}-frame synthetic '{a.scene}' sequence 640x480, depth-range mask "
            f"depth_trunc 3 m, voxel {a.voxel*1000:g} mm / trunc {4*a.voxel*1000:g} mm")


def config_of(a, world):
    """The `config` object: a function of the command line only, so the reference arm and ours print the same one."""
    W, H = (1280, 720) if a.hd else (640, 480)
    par = "single GPU" if world == 1 else (
        f"{['x-axis', 'y-axis', 'z-axis', 'diagonal (kx+ky)'][a.slab_axis]} slabs of {a.slab_thickness} block(s), cyclic over {world} ranks, " +
        ("+1 block halo integrated redundantly" if a.slab_halo else "owned blocks only; boundary planes exchanged once before extraction"))
    return {"workload": workload_name(a), "frames_per_step": a.frames, "frames_per_launch": a.batch, "width": W, "height": H,
            "voxel_mm": a.voxel * 1000, "trunc_mm": 4 * a.voxel * 1000, "parallelism": par,
            "l2": "inputs (%.0f MB per step) exceed the 126 MB L2; no flush needed" % (a.frames * W * H * 5 / 1e6)}


def host_cores():
    """Host threads this process may use (affinity-aware); torchrun's OMP_NUM_THREADS=1 is deliberately ignored."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:  # noqa: BLE001
        return max(1, os.cpu_count() or 1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(a, seq_np, budget_s=12.0):
    """Time the oracle on a bounded sample of the same sequence: frames k, k+8, k+16, ... (each
    sub-pass covers the whole trajectory), then k+1, k+9, ... into the same volume, until ~budget_s
    of CPU work (or --cpu-frames frames) is done; wraps around for sequences shorter than the budget."""
    from oracle import open3d_ref, oracle
    depth, rgb, extr, fxfycxcy = seq_np
    n = len(depth)
    cores = oracle.set_num_threads(host_cores())
    stride = 8 if n >= 64 else 1
    order = [k for off in range(stride) for k in range(off, n, stride)]
    sample = (f"frames of the {n}-frame sequence (every {stride}th frame, then the next offset, ...) into one volume "
              f"(depth convert + allocate + integrate per frame)")
    if open3d_ref.available() and not os.environ.get("OTSLAM_BENCH_NO_OPEN3D"):
        # the real backend of the reference (reconstruct_rgbd.py:99-107), same frames, same order
        H, W = depth.shape[1:]
        done, dt = open3d_ref.time_frame_loop(depth, rgb, (W, H) + tuple(fxfycxcy), extr, a.voxel, 4 * a.voxel, budget_s, order)
        return {"value": done / dt, "unit": UNIT, "cores": cores, "kind": "reference", "backend": "open3d " + open3d_ref.version(),
                "sample": f"{done} {sample}, {dt:.1f} s of CPU work, Open3D's own OpenMP threading on {cores} host threads"}, dt, done
    vol = oracle.Volume(a.voxel, 4 * a.voxel)
    done = 0
    t0 = time.perf_counter()
    while True:
        k = order[done % n]
        vol.integrate(oracle.depth_convert(depth[k], 1000.0, 3.0), rgb[k], fxfycxcy, extr[k])
        done += 1
        dt = time.perf_counter() - t0
        if (a.cpu_frames and done >= a.cpu_frames) or (not a.cpu_frames and dt >= budget_s) or done >= 64 * n:
            break
    return {"value": done / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{done} {sample}, {dt:.1f} s of CPU work on {cores} threads"}, dt, done


def cpu_full_loop(a, seq, n_files=48, budget_s=6.0):
    """The reference's WHOLE per-frame loop on the host (reconstruct_rgbd.py:86-109): read_image x2 (JPEG + 16-bit PNG
    decode), np.loadtxt, pose @ T_fix, np.linalg.inv, create_from_color_and_depth, integrate -- sequential, as the script
    runs it, from a capture tree written in scanner_node.cpp's format.  Integration uses all host threads (oracle port, or
    open3d when importable); decoding is single-threaded like the reference's."""
    import shutil
    import tempfile
    import cv2
    import numpy as np
    from oracle import open3d_ref, oracle
    from otslam_b200 import capture, synth
    depth, rgb = seq.numpy()
    n = min(n_files, len(seq))
    idx = [int(i) for i in range(0, len(seq), max(1, len(seq) // n))][:n]
    base = tempfile.mkdtemp(prefix="otslam_cpu_loop_")
    cores = oracle.set_num_threads(host_cores())
    use_o3d = open3d_ref.available() and not os.environ.get("OTSLAM_BENCH_NO_OPEN3D")
    try:
        for j, k in enumerate(idx):
            capture.save_frame(base, "Object_0", j + 1, rgb[k], depth[k], seq.pose_ros[k])
        H, W = depth.shape[1:]
        vol = open3d_ref.make_volume(a.voxel, 4 * a.voxel) if use_o3d else oracle.Volume(a.voxel, 4 * a.voxel)
        intr = open3d_ref.intrinsic(W, H, *seq.fxfycxcy) if use_o3d else None
        done, t0 = 0, time.perf_counter()
        while True:
            j = done % n + 1
            col = cv2.cvtColor(cv2.imread(os.path.join(base, "color", f"Object_0_{j}.jpg"), cv2.IMREAD_UNCHANGED), cv2.COLOR_BGR2RGB)
            dep = cv2.imread(os.path.join(base, "depth", f"Object_0_{j}.png"), cv2.IMREAD_UNCHANGED)
            pose_ros = np.loadtxt(os.path.join(base, "poses", f"Object_0_{j}.txt"))
            extrinsic = np.linalg.inv(pose_ros @ synth.T_FIX)
            if use_o3d:
                open3d_ref.integrate(vol, dep, col, intr, extrinsic)
            else:
                vol.integrate(oracle.depth_convert(dep, 1000.0, 3.0), col, seq.fxfycxcy, extrinsic)
            done += 1
            dt = time.perf_counter() - t0
            if dt >= budget_s or done >= 16 * n:
                break
        return {"value": done / dt, "unit": UNIT, "cores": cores, "kind": "reference" if use_o3d else "port",
                "sample": f"{done} frames ({n} distinct capture triples) decoded + integrated sequentially in {dt:.1f} s"}
    finally:
        shutil.rmtree(base, ignore_errors=True)


def post_stage(a, vol, n_blocks, peak):
    """The rest of reconstruct_rgbd_filter.py after the frame loop (:113-134: extract_triangle_mesh,
    compute_vertex_normals, sample_points_uniformly(100000), z >= 0.03 mask) plus the north_star
    filters (voxel_down_sample, remove_statistical_outlier), each once through the C ABI with host
    buffers.  wall_ms includes the host<->device copies; device_ms is the kernel section alone (CUDA
    events inside the library); GB/s = SURVEY 8(d) algorithmic bytes / device_ms.  The filters are
    also run on a 1 M-point cloud (the config-5 object size): 100 000 points are launch-latency
    bound and say nothing about bandwidth."""
    import ctypes as C
    import numpy as np
    import otslam_b200.o3d_compat as o3d
    from otslam_b200 import _lib
    out = {}

    def run(name, fn, alg_bytes=None, **info):
        fn()                                                    # warm (first-use allocations, tables)
        t0 = time.perf_counter()
        r = fn()
        wall = time.perf_counter() - t0
        dev = _lib.last_op_device_ms()
        e = {"wall_ms": 1e3 * wall, "device_ms": dev}
        if alg_bytes is not None:
            b = float(alg_bytes(r))
            e["algorithmic_MB"] = b / 1e6
            if dev > 0:
                e["device_GBps"] = b / (dev * 1e-3) / 1e9
                e["frac_of_hbm_peak"] = e["device_GBps"] / peak
        e.update(info)
        out[name] = e
        return r

    # reconstruct_rgbd_filter.py:113-123 as the drop-in runs it: the mesh stays in HBM (lazy TriangleMesh)
    def extract():
        m = o3d.geometry.TriangleMesh()
        nv, nf = vol.extract_mesh_resident(owner=m)
        m._attach_resident(vol, nv, nf, True)
        m.compute_vertex_normals()
        return m

    mesh = run("extract_mesh+normals", extract, lambda m: 8 * 4096 * n_blocks + 48 * len(m.vertices) + 12 * len(m.triangles))
    out["extract_mesh+normals"].update(vertices=int(len(mesh.vertices)), faces=int(len(mesh.triangles)))

    def zfilter(pc, zmin=0.03):
        n = len(pc.points)
        op, oc, m = np.empty((n, 3)), np.empty((n, 3)), C.c_int64(0)
        _lib.check(_lib.lib.otslam_cloud_zfilter(_lib.ptr(pc._points), _lib.ptr(pc._colors), n, zmin, _lib.ptr(op), _lib.ptr(oc),
                                                 C.byref(m), 0))
        r = o3d.geometry.PointCloud()
        r._points, r._colors = np.ascontiguousarray(op[:m.value]), np.ascontiguousarray(oc[:m.value])
        return r

    for tag, n_samples in (("", 100000), ("_1M", 1000000)):
        pcd = run("sample_points_uniformly" + tag, lambda: mesh.sample_points_uniformly(n_samples, seed=0), points=n_samples)
        flt = run("z_mask" + tag, lambda: zfilter(pcd), lambda r: 48 * n_samples + 48 * len(r.points))
        out["z_mask" + tag]["kept"] = int(len(flt.points))
        # the script's own chain filters what survives the z mask; the "_1M" legs feed the filters the full 1 M-point cloud
        # (the config-5 object size) so that their GB/s figures are measured at the stated size
        src = flt if not tag else pcd
        n = len(src.points)
        ds = run("voxel_down_sample" + tag, lambda: src.voxel_down_sample(0.01), lambda r: 36 * (n + len(r.points)), voxel_m=0.01)
        out["voxel_down_sample" + tag].update(points_in=n, points_out=int(len(ds.points)))
        sel = run("remove_statistical_outlier" + tag, lambda: src.remove_statistical_outlier(20, 2.0),
                  lambda r: 24 * n * 2 + 8 * n + 36 * len(r[1]), nb_neighbors=20, std_ratio=2.0)
        out["remove_statistical_outlier" + tag].update(points_in=n, kept=int(len(sel[1])))

    # the same chain with the cloud resident in HBM between the operators (otslam_b200.cloud.DeviceCloud: the C-ABI operators
    # take device pointers); one download at the end.  wall_ms covers the whole chain incl. that download.
    from otslam_b200.cloud import DeviceCloud
    import torch

    def chain(n_samples, zmin):
        dc = DeviceCloud.sample_mesh(vol, n_samples, 0, colors=True)
        if zmin is not None:
            dc = dc.zfilter(zmin)
        dv = dc.voxel_down_sample(0.01)
        dsor, _ = dv.remove_statistical_outlier(20, 2.0)
        return dsor.to_pointcloud(), len(dc), len(dv)

    def host_chain(n_samples, zmin):
        pc = mesh.sample_points_uniformly(n_samples, seed=0)
        if zmin is not None:
            pc = zfilter(pc, zmin)
        hv = pc.voxel_down_sample(0.01)
        hs, _ = hv.remove_statistical_outlier(20, 2.0)
        return hs, len(pc.points), len(hv.points)

    for tag, n_samples, zmin in (("", 100000, 0.03), ("_1M", 1000000, None)):
        for name, fn in (("device_resident", chain), ("host_arrays", host_chain)):
            fn(n_samples, zmin)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res, n_in, n_vds = fn(n_samples, zmin)
            torch.cuda.synchronize()
            out[f"chain_sample_mask_vds_sor{tag}:{name}"] = {"wall_ms": 1e3 * (time.perf_counter() - t0), "points_sampled": n_samples,
                                                            "points_after_mask": n_in, "after_vds": n_vds, "kept": int(len(res.points))}
    # reconstruct_rgbd.py writes the mesh instead: the one-off download of the resident arrays
    t0 = time.perf_counter()
    mesh._materialize()
    out["mesh_download_for_write_triangle_mesh"] = {"wall_ms": 1e3 * (time.perf_counter() - t0),
                                                    "bytes": int(len(mesh.vertices)) * 72 + int(len(mesh.triangles)) * 12}
    out["total_wall_ms_config1_stage"] = sum(out[k]["wall_ms"] for k in (
        "extract_mesh+normals", "sample_points_uniformly", "z_mask", "voxel_down_sample", "remove_statistical_outlier"))
    return out


def config3_stage(a, device, frames_per_object=150, repeats=3):
    """BASELINE configs[2] (multi_reconstruct_rgbd_filter.py:139-145): four objects (table, chair, cone, cardboard), each its
    own volume in the reference, one after the other.  Frames resident in HBM, 640x480, the bench's voxel size.
    sequential = four volumes integrated one after the other (the reference's structure on the GPU); arena = ONE multi-object
    arena (object id in the block key) fed with the four objects' frames interleaved as pipeline.interleave_plan orders them:
    one work list and one integration launch per 32-frame batch over the union of the objects' blocks.  A single small
    object's batch touches a few hundred blocks and cannot fill 148 SMs; the union can.  Wall clock around the synchronous
    calls, best of `repeats`; every object's statistics must be equal both ways."""
    import numpy as np
    import torch
    from otslam_b200 import pipeline, synth
    from otslam_b200.volume import ArenaView, TSDFVolume
    scenes = ("table", "chair", "cone", "cardboard")
    seqs = [synth.make_sequence(s, frames_per_object, device=device) for s in scenes]
    counts = [len(s) for s in seqs]
    order = pipeline.interleave_plan(counts)
    base = np.cumsum([0] + counts[:-1])
    idx = torch.tensor([int(base[o]) + k for o, k in order], device=device)
    # (indexing through an int16 view: advanced indexing of uint16 tensors is not available on every torch build)
    depth = torch.cat([s.depth for s in seqs]).view(torch.int16)[idx].view(seqs[0].depth.dtype).contiguous()
    rgb = torch.cat([s.rgb for s in seqs])[idx].contiguous()
    ext = np.concatenate([s.extrinsic for s in seqs])[idx.cpu().numpy()]
    ids = np.array([o for o, _ in order], np.int32)
    dev_index = torch.device(device).index or 0
    vols = [TSDFVolume(a.voxel, 4 * a.voxel, device=dev_index) for _ in scenes]
    arena = TSDFVolume(a.voxel, 4 * a.voxel, device=dev_index)
    arena.set_objects(len(scenes))
    try:
        t_seq = t_arena = 1e30
        for _ in range(repeats + 1):                              # first pass warms pools and staging
            for v in vols:
                v.reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for v, s in zip(vols, seqs):
                v.integrate_batch(s.depth.contiguous(), s.rgb.contiguous(), s.fxfycxcy, s.extrinsic)
            t_seq = min(t_seq, time.perf_counter() - t0)
            arena.reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            arena.integrate_batch(depth, rgb, seqs[0].fxfycxcy, ext, object_ids=ids)
            t_arena = min(t_arena, time.perf_counter() - t0)
        per_obj = [v.stats() for v in vols]
        same = all(ArenaView(arena, o).stats() == per_obj[o] for o in range(len(scenes)))
        n = sum(counts)
        return {"objects": list(scenes), "frames": n, "frames_per_object": frames_per_object, "voxel": a.voxel,
                "sequential_frames_per_s": n / t_seq, "arena_frames_per_s": n / t_arena, "arena_speedup": t_seq / t_arena,
                "identical_per_object": bool(same), "blocks_per_object": [st["n_blocks"] for st in per_obj],
                "note": "frames resident in HBM; wall clock of the synchronous C-ABI calls, best of %d" % repeats}
    finally:
        for v in vols:
            v.close()
        arena.close()


def hybrid_merge_stage(peak, n_obj=20, per_obj=1_000_000, map_px=2000):
    """BASELINE.json configs[4]: fusion/hybrid_map.py -- 2-D occupancy grid -> Z = 0 points (reference :45-55, a Python
    per-pixel loop) and the paint + concatenate + PLY-record packing of `n_obj` object clouds of `per_obj` points
    (reference :59,88-91,115,121), through the C ABI with host buffers.  Algorithmic bytes per SURVEY 8(d):
    1 B x w*h + 24 B x sum(N_obj) read, 27 B x (P + sum(N_obj)) written."""
    import ctypes as C
    import numpy as np
    import otslam_b200.o3d_compat as o3d
    from otslam_b200 import _lib, synth
    out = {}
    img = synth.occupancy_map(map_px, map_px, 0.02, 0)
    mp = np.empty((img.size, 3))
    nmap = C.c_int64(0)

    def grid():
        _lib.check(_lib.lib.otslam_grid_to_points(_lib.ptr(img), img.shape[1], img.shape[0], 0.05, -50.0, -50.0, 100, _lib.ptr(mp),
                                                  C.byref(nmap), 0))
    grid()
    t0 = time.perf_counter()
    grid()
    out["grid_to_points"] = {"wall_ms": 1e3 * (time.perf_counter() - t0), "device_ms": _lib.last_op_device_ms(), "pixels": int(img.size),
                             "points": int(nmap.value)}
    rng = np.random.default_rng(0)
    objs = [rng.random((per_obj, 3)) for _ in range(n_obj)]
    clouds = [mp[:nmap.value]] + objs
    paint = [[0.2, 0.2, 0.2]] + [[1.0, 0.0, 0.0]] * n_obj
    o3d.io.pack_cloud_records(clouds, paint=paint)
    t0 = time.perf_counter()
    rec = o3d.io.pack_cloud_records(clouds, paint=paint)
    wall, dev = time.perf_counter() - t0, _lib.last_op_device_ms()
    b = img.size + 24.0 * n_obj * per_obj + 27.0 * len(rec)
    out["merge_paint_pack"] = {"wall_ms": 1e3 * wall, "device_ms": dev, "points": int(len(rec)), "algorithmic_MB": b / 1e6,
                               "device_GBps": b / (dev * 1e-3) / 1e9 if dev > 0 else None,
                               "frac_of_hbm_peak": b / (dev * 1e-3) / 1e9 / peak if dev > 0 else None,
                               "note": "wall = pageable H2D of 480 MB + kernel + D2H of 540 MB (PCIe-bound); device = the kernel"}
    return out


def files_e2e(a, seq, n_files=768):
    """The drop-in script's loop as a user runs it: a capture tree on disk (color/*.jpg, depth/*.png, poses/*.txt as
    scanner_node.cpp writes them) -> pipeline.integrate_files (thread-pool decode one chunk ahead of the GPU) into
    a fresh ScalableTSDFVolume.  File decoding, not the GPU, bounds this number; the sequential decode rate is what
    the reference's own per-frame loop (reconstruct_rgbd.py:86-109) pays on top of its CPU integration."""
    import shutil
    import tempfile
    import numpy as np
    import otslam_b200.o3d_compat as o3d
    from otslam_b200 import pipeline, synth
    depth, rgb = seq.numpy()
    n = min(n_files, len(seq))
    idx = [int(i) for i in range(0, len(seq), max(1, len(seq) // n))][:n]
    base = tempfile.mkdtemp(prefix="otslam_bench_")
    try:
        from otslam_b200 import capture
        for j, k in enumerate(idx):
            capture.save_frame(base, "Object_0", j + 1, rgb[k], depth[k], seq.pose_ros[k])
        triples = [(os.path.join(base, "color", f"Object_0_{j}.jpg"), os.path.join(base, "depth", f"Object_0_{j}.png"),
                    os.path.join(base, "poses", f"Object_0_{j}.txt"), j) for j in range(1, len(idx) + 1)]
        intr = o3d.camera.PinholeCameraIntrinsic(*seq.intr)
        t0 = time.perf_counter()
        for t in triples[:32]:
            pipeline.load_frame(t[0], t[1], t[2], intr, synth.T_FIX)
        seq_decode_fps = 32 / (time.perf_counter() - t0)
        vol = o3d.pipelines.integration.ScalableTSDFVolume(voxel_length=a.voxel, sdf_trunc=4 * a.voxel,
                                                           color_type=o3d.pipelines.integration.TSDFVolumeColorType.RGB8)

        def timed_pass():
            """(frames, best wall time of three passes, all pass times in ms) after one warm pass (file cache, staging /
            decoder buffers, block-pool growth); wall clock, one process, so the best pass is the one nothing else disturbed"""
            pipeline.integrate_files(vol, triples, intr, synth.T_FIX)
            times = []
            for _ in range(3):
                vol.reset()
                t0 = time.perf_counter()
                m = pipeline.integrate_files(vol, triples, intr, synth.T_FIX)
                times.append(time.perf_counter() - t0)
            return m, min(times), [round(1e3 * t, 1) for t in times]

        # (1) JPEG / PNG decoded by OpenCV threads on the host (round 1's loop)
        os.environ["OTSLAM_GPU_DECODE"] = "0"
        try:
            done_h, dt_h, passes_h = timed_pass()
            ref = vol._vol.stats()
        finally:
            os.environ.pop("OTSLAM_GPU_DECODE", None)
        # (2) the default: compressed bytes uploaded, inflate / filters / Huffman / IDCT / colour on the GPU (csrc/imgcodec.cu)
        vol.reset()
        done, dt, passes = timed_pass()
        chunks = list(pipeline.last_decode_profile)
        keys = ("inflate_ms", "png_filter_emit_ms", "jpeg_huffman_ms", "jpeg_idct_ms", "jpeg_color_ms")
        failed = pipeline.gpu_decode_selfcheck["failed"]
        out = {"frames": done, "frames_per_s": done / dt, "pass_ms": passes, "decode": "gpu" if not failed else "host (GPU self-check failed: %s)" % failed,
               "identical_volume": vol._vol.stats() == ref,
               "chunks_in_preparation": pipeline.DECODE_AHEAD,
               "host_threads": pipeline._decode_workers(),
               "decoder_device_ms_per_chunk": {k: float(np.mean([c[k] for c in chunks])) for k in keys} if chunks else None,
               "chunk_frames": pipeline.CHUNK_FRAMES,
               "compressed_bytes_per_frame": int(sum(c["compressed_bytes"] for c in chunks) / max(1, done)),
               "raw_bytes_per_frame": int(seq.intr[0] * seq.intr[1] * 5),
               "frames_passed_to_host_decoders": int(sum(c["passed_on"] for c in chunks)),
               "host_decode": {"frames": done_h, "frames_per_s": done_h / dt_h, "pass_ms": passes_h, "decode_threads": pipeline._decode_workers()},
               "sequential_decode_frames_per_s": seq_decode_fps,
               "note": "capture tree on disk -> frames in the volume; default = GPU decoders (host threads only read and frame the "
                       "files), host_decode = OpenCV threads (OTSLAM_GPU_DECODE=0), sequential = the reference's one-core loop"}
        # (3) the decoders alone: one chunk of files -> decoded frames in HBM (file reads from the page cache, upload of the
        # compressed bytes and all six kernels inside the timed call)
        try:
            from otslam_b200.decoder import FrameDecoder
            m = min(pipeline.CHUNK_FRAMES, len(triples))
            dec = FrameDecoder(int(seq.intr[1]), int(seq.intr[0]), m)
            cps, dps = [t[0] for t in triples[:m]], [t[1] for t in triples[:m]]
            dec.decode_files(cps, dps)
            t0 = time.perf_counter()
            for _ in range(3):
                cs, ds = dec.decode_files(cps, dps)
            dt3 = (time.perf_counter() - t0) / 3
            prof = dec.profile()
            out["decode_only"] = {"frames": m, "frames_per_s": m / dt3, "wall_ms": 1e3 * dt3, "all_decoded": bool((cs == 0).all() and (ds == 0).all()),
                                  "device_ms": {k: prof[k] for k in keys},
                                  "decoded_GBps": m * seq.intr[0] * seq.intr[1] * 5 / dt3 / 1e9}
            dec.close()
        except Exception as e:  # noqa: BLE001
            out["decode_only"] = {"error": repr(e)}
        # the same loop served from the raw side-car (OTSLAM_SIDECAR=1: decoded frames cached next to the tree on the first
        # pass): identical volumes, no JPEG / PNG decode on later passes
        os.environ["OTSLAM_SIDECAR"] = "1"
        try:
            vol.reset()
            pipeline.integrate_files(vol, triples, intr, synth.T_FIX)            # writes the side-car
            vol.reset()
            pipeline.integrate_files(vol, triples, intr, synth.T_FIX)            # warm: page cache + staging
            vol.reset()
            t0 = time.perf_counter()
            done2 = pipeline.integrate_files(vol, triples, intr, synth.T_FIX)
            dt2 = time.perf_counter() - t0
            out["sidecar"] = {"frames": done2, "frames_per_s": done2 / dt2, "identical_volume": vol._vol.stats() == ref,
                              "bytes_per_frame": int(seq.intr[0] * seq.intr[1] * 5),
                              "note": "raw depth + rgb read straight into the staging buffers by the same thread pool"}
        finally:
            os.environ.pop("OTSLAM_SIDECAR", None)
        return out
    finally:
        shutil.rmtree(base, ignore_errors=True)


def make_sequence(a, device):
    from otslam_b200 import synth
    intr = synth.HD_INTRINSICS if a.hd else synth.REF_INTRINSICS
    scene = "room" if a.hd else a.scene
    return synth.make_sequence(scene, a.frames, intr=intr, device=device)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    seq = make_sequence(a, "cpu")
    d, c = seq.numpy()
    seq_np = (d, c, seq.extrinsic, seq.fxfycxcy)
    # each step = a bounded sample of the workload
    res = None
    vals = []
    budget = max(2.0, min(12.0, 150.0 / max(1, a.warmup + a.steps)))      # whole run stays within a few minutes
    for s in range(a.warmup + a.steps):
        res, dt, nf = cpu_sample(a, seq_np, budget)
        if s >= a.warmup:
            vals.append((nf, dt))
    tot_f = sum(v[0] for v in vals); tot_t = sum(v[1] for v in vals)
    value = tot_f / tot_t
    res["value"] = value
    try:
        res["full_loop"] = cpu_full_loop(a, seq)
    except Exception as e:  # noqa: BLE001 -- informational
        res["full_loop"] = {"error": repr(e)}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * tot_t / max(1, a.steps), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(a, max(1, a.gpus)),
            "reference_note": ("reference CPU path = " + (res.get("backend") or "oracle port of Open3D legacy ScalableTSDFVolume "
                               "(open3d itself is not installable offline)") + "; each step is a bounded sample of the workload"),
            "cpu_baseline": res,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def measure(a, seq, world, rank, local, steps, warmup, with_e2e=True, hd=False):
    """Time `steps` passes of the frame loop over `seq` (resident in HBM), then the same from pinned host memory (e2e),
    then -- multi-GPU -- the halo exchange and the NCCL gather of the extracted points, and the whole job end to end
    (e2e_full).  Returns a dict; rank 0's is printed."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from otslam_b200 import _lib
    from otslam_b200 import slab as slabmod
    from otslam_b200.volume import TSDFVolume
    dev = f"cuda:{local}"
    W, H = seq.intr[0], seq.intr[1]
    n = len(seq)
    depth_dev, rgb_dev = seq.depth.contiguous(), seq.rgb.contiguous()
    slab = None if world == 1 else (a.slab_axis, a.slab_thickness, world, rank, a.slab_halo)
    if world == 1 and a.emulate_world > 1:
        slab = (a.slab_axis, a.slab_thickness, a.emulate_world, a.emulate_rank, a.slab_halo)
    voxel = a.hd_voxel if hd else a.voxel
    vol = TSDFVolume(voxel, 4 * voxel, device=local, slab=slab)
    vol.set_batch(a.batch)
    if a.zsplit:
        vol.set_zsplit(a.zsplit)
    stream = torch.cuda.Stream()
    vol.set_stream(stream.cuda_stream)
    torch.cuda.synchronize()

    def step():
        vol.reset()
        vol.integrate_batch(depth_dev, rgb_dev, seq.fxfycxcy, seq.extrinsic)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.cuda.stream(stream):
        for _ in range(warmup):
            step()
        stats = vol.stats()
        n_upd = stats["weight_sum"]
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        vol.profile(True)
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count() - l0
        prof = vol.profile(False)
        clocks = sampler.stop() if rank == 0 else None
    ms_max = max_over_ranks(ms)
    value = n * steps / (ms_max * 1e-3)
    # totals over the ranks (each rank integrates its own slabs): algorithmic bytes are a property of the whole job
    tot = torch.tensor([float(n_upd), float(stats["n_blocks"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    n_upd_all, n_blocks_all = int(tot[0].item()), int(tot[1].item())

    # ---- roofline of the dominant kernel (this rank's share)
    peak, peak_src = peaks()
    k4_ms, k4_launches = prof["integrate"]
    bytes_step = 40.0 * n_upd + 5.0 * W * H * n
    bytes_per_launch = bytes_step * steps / max(1, k4_launches)
    k4_avg_ms = k4_ms / max(1, k4_launches)
    achieved = bytes_per_launch / (k4_avg_ms * 1e-3) / 1e9 if k4_avg_ms > 0 else 0.0
    traffic, issue_active = None, None
    tp = os.path.join(ROOT, "profiles", "integrate_kernel_traffic.json")
    if os.path.exists(tp) and world == 1 and not hd and not a.emulate_world and a.scene == "chair_table" and a.frames == 1000:
        try:                                   # ncu capture of exactly this workload / launch shape; not valid for other runs
            j = json.load(open(tp))
            traffic, issue_active = j.get("dram_bytes_per_launch"), j.get("issue_active_frac")
        except Exception:  # noqa: BLE001
            pass
    roofline = {"bound": "issue", "kernel": "integrate_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bound_note": "frac = ALGORITHMIC bytes (40 B x N_upd + 5 B x W x H per frame, SURVEY 8d) / kernel time / measured HBM peak; "
                              "32-frame fusion makes the physical DRAM traffic several times smaller than the algorithmic bytes, so the "
                              "kernel's real limiter is instruction issue (issue_active_frac, from the committed ncu capture), not HBM",
                "issue_active_frac": issue_active,
                "algorithmic_bytes_per_launch": bytes_per_launch, "launch_ms": k4_avg_ms,
                "launches_per_step": k4_launches / max(1, steps),
                "n_upd_per_frame": n_upd / n, "kernel_share_of_step": k4_ms / ms if ms > 0 else None,
                "other_kernels_ms_per_step": {"pack": prof["pack"][0] / steps, "alloc": prof["alloc"][0] / steps,
                                              "note": "pack/alloc of batch b+1 run on a second stream UNDER integrate(b); "
                                                      "their event spans include that contention and are not additive"}}

    out = {"value": value, "ms_per_step": ms_max / steps, "frames": n, "n_blocks": n_blocks_all, "n_blocks_this_rank": stats["n_blocks"],
           "n_upd_per_frame_all_ranks": n_upd_all / n, "roofline": roofline, "gpu_launches": int(launches), "clocks": clocks,
           "achieved_hbm_gbs_whole_step": (40.0 * n_upd_all + 5.0 * W * H * n) * steps / (ms_max * 1e-3) / 1e9,
           "_vol": vol, "_stream": stream, "_stats": stats}

    # ---- end to end through the C ABI with host buffers (pinned copy of the sequence on every rank)
    if with_e2e:
        # pinned host copy of the frames THIS rank uploads: everything on one GPU, its 1/world share of every chunk otherwise
        hd_, hc_ = slabmod.rank_shards(depth_dev, rgb_dev, rank, world, a.ingest_chunk)
        torch.cuda.synchronize()

        def ingest():
            vol.reset()
            if world > 1:       # frames cross PCIe once per box: 1/world per rank, all-gather over NVLink (slab.py)
                slabmod.integrate_host_sharded(vol, hd_, hc_, seq.fxfycxcy, seq.extrinsic, rank, world, dev, stream=stream,
                                               chunk_frames=a.ingest_chunk, shards=True)
            else:
                vol.integrate_batch(hd_, hc_, seq.fxfycxcy, seq.extrinsic)

        def e2e_step():
            ingest()
            return vol.stats()            # D2H read of the step's result

        with torch.cuda.stream(stream):
            for _ in range(max(1, min(2, warmup))):
                e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                st = e2e_step()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        assert st["weight_sum"] == n_upd, "host-path result differs from the resident-path result"
        out["e2e"] = {"value": n * steps / max_over_ranks(dt), "unit": UNIT,
                      "h2d_bytes_per_step": int(n * W * H * 5 + n * 16 * 8), "d2h_bytes_per_step": 16 + 16,
                      "api": "otslam_volume_reset + otslam_volume_integrate_batch(OTSLAM_MEM_HOST, pinned) + otslam_volume_stats" if world == 1 else
                             f"otslam_volume_reset + slab.integrate_host_sharded (H2D of 1/world of each {a.ingest_chunk}-frame chunk per rank, "
                             "ncclAllGather over NVLink, otslam_volume_integrate_batch on the resident chunk) + otslam_volume_stats"}

        # ---- the whole job as a user runs it, wall clock: frames in host memory -> final cloud in host memory.
        #      1 GPU = reconstruct_rgbd_filter.py:86-134 (frame loop, extract_triangle_mesh, compute_vertex_normals,
        #      sample_points_uniformly(100000), z >= 0.03 mask); N GPUs = frame loop on slabs, halo exchange, per-slab
        #      extract_point_cloud, NCCL gather to rank 0, download there.
        def full_step():
            t = [time.perf_counter()]
            ingest()
            torch.cuda.synchronize(); t.append(time.perf_counter())
            if world == 1:
                nv, nf = vol.extract_mesh_resident()
                t.append(time.perf_counter())
                if nv:      # as the drop-in script does it: sample and z mask stay in HBM, one (pinned) download of the survivors
                    from otslam_b200.cloud import DeviceCloud
                    res = DeviceCloud.sample_mesh(vol, 100000, 0, colors=True).zfilter(0.03).to_numpy()[:2]
                else:
                    res = (np.zeros((0, 3)), np.zeros((0, 3)))
                t.append(time.perf_counter())
                names = ("ingest+integrate", "extract_mesh+normals", "sample+z_mask+download")
            else:
                if not a.slab_halo:
                    slabmod.exchange_halo(vol, rank, world, device=dev)
                torch.cuda.synchronize(); t.append(time.perf_counter())
                res = slabmod.extract_and_gather_points(vol, rank, world, device=dev, as_numpy=False)
                torch.cuda.synchronize(); t.append(time.perf_counter())
                if rank == 0:
                    res = slabmod.to_host(res[:2], copy=False)
                t.append(time.perf_counter())
                names = ("ingest+integrate", "halo_exchange", "extract+nccl_gather", "download_rank0")
            return res, dict(zip(names, (1e3 * (y - x) for x, y in zip(t, t[1:]))))

        with torch.cuda.stream(stream):
            full_step()
            barrier()
            t0 = time.perf_counter()
            res, parts = full_step()
            torch.cuda.synchronize()
            dtf = time.perf_counter() - t0
        dtf = max_over_ranks(dtf)
        out["e2e_full"] = {"value": n / dtf, "unit": UNIT, "wall_ms": 1e3 * dtf, "stages_ms_rank0": parts,
                           "h2d_bytes": int(n * W * H * 5 + n * 16 * 8),
                           "d2h_bytes": int(sum(x.nbytes for x in res)) if (rank == 0 and res is not None) else 0,
                           "result_points": int(len(res[0])) if (rank == 0 and res is not None) else 0,
                           "note": "one pass, wall clock, max over ranks; host frames -> final cloud in host memory"}
        del hd_, hc_

    # ---- multi-GPU: halo exchange + extraction + NCCL gather of the extracted points, timed on their own
    if world > 1:
        with torch.cuda.stream(stream):
            t_halo, t_gather = float("inf"), float("inf")
            for it in range(3):                     # pass 0 warms NCCL's p2p channels, the scratch cache and the hash capacity;
                step()                              # the figure is the better of passes 1 and 2 (max over ranks each)
                barrier()
                t0 = time.perf_counter()
                got = 0 if a.slab_halo else slabmod.exchange_halo(vol, rank, world, device=dev)
                torch.cuda.synchronize()
                th = max_over_ranks(time.perf_counter() - t0)
                barrier()
                t0 = time.perf_counter()
                pts = slabmod.extract_and_gather_points(vol, rank, world, device=dev, as_numpy=False)
                torch.cuda.synchronize()
                tg = max_over_ranks(time.perf_counter() - t0)
                if it:
                    t_halo, t_gather = min(t_halo, th), min(t_gather, tg)
        out["halo_exchange_ms"] = 1e3 * t_halo
        out["halo_pieces_received_rank0"] = int(got)
        out["extract_gather_ms"] = 1e3 * t_gather
        out["gathered_points"] = int(pts[0].shape[0]) if rank == 0 else 0
        out["gather_note"] = "device-resident: pack kernel -> ncclSend/Recv on HBM buffers -> import kernel; extracted points HBM -> NVLink -> rank 0 HBM"
        del pts
    return out


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    seq = make_sequence(a, f"cuda:{local}")
    W, H = seq.intr[0], seq.intr[1]
    n = len(seq)
    m = measure(a, seq, world, rank, local, a.steps, a.warmup, with_e2e=not a.no_e2e, hd=a.hd)
    vol, stats = m.pop("_vol"), m.pop("_stats")
    m.pop("_stream")
    peak, _ = peaks()

    post, cpu = None, None
    if rank == 0 and world == 1:
        if not a.no_cpu:
            d, c = seq.numpy()
            cpu, _, _ = cpu_sample(a, (d, c, seq.extrinsic, seq.fxfycxcy))
            try:
                cpu["full_loop"] = cpu_full_loop(a, seq)
            except Exception as e:  # noqa: BLE001 -- informational
                cpu["full_loop"] = {"error": repr(e)}
        if not a.no_post and not a.emulate_world:
            vol.reset()
            vol.integrate_batch(seq.depth.contiguous(), seq.rgb.contiguous(), seq.fxfycxcy, seq.extrinsic)
            post = post_stage(a, vol, stats["n_blocks"], peak)
            try:
                post["hybrid_map_config5"] = hybrid_merge_stage(peak)
            except Exception as e:  # noqa: BLE001
                post["hybrid_map_config5"] = {"error": repr(e)}
            try:
                post["config3_four_objects"] = config3_stage(a, f"cuda:{local}")
            except Exception as e:  # noqa: BLE001 -- informational
                post["config3_four_objects"] = {"error": repr(e)}
            try:
                post["files_e2e"] = files_e2e(a, seq)
            except Exception as e:  # noqa: BLE001 -- informational; never let it take the bench line down
                post["files_e2e"] = {"error": repr(e)}
    vol.close()
    del vol, seq
    torch.cuda.empty_cache()

    # ---- configs[3] (north_star's scaling config): 1280x720 / 2 mm large room, same slabs, in the same JSON line so that
    #      the driver's 1/2/4/8-GPU runs record it (efficiency = hd.value(N) / (N x hd.value(1)))
    hd = None
    if not a.hd and a.hd_frames > 0 and not a.emulate_world:
        import copy
        from otslam_b200 import synth
        b = copy.copy(a)
        b.hd, b.frames, b.voxel = True, a.hd_frames, a.hd_voxel
        t0 = time.perf_counter()
        hseq = synth.make_sequence("room", a.hd_frames, intr=synth.HD_INTRINSICS, device=f"cuda:{local}")
        synth_s = time.perf_counter() - t0
        h = measure(b, hseq, world, rank, local, max(1, a.hd_steps), 1, with_e2e=not a.no_e2e, hd=True)
        h.pop("_vol").close(); h.pop("_stats"); h.pop("_stream")
        r = h["roofline"]
        hd = {"workload": workload_name(b), "frames": a.hd_frames, "steps": max(1, a.hd_steps), "warmup": 1, "value": h["value"], "unit": UNIT,
              "ms_per_step": h["ms_per_step"], "e2e": (h.get("e2e") or {}).get("value"), "e2e_full": h.get("e2e_full"),
              "frac": r["frac"], "achieved_gbs": r["achieved"], "launch_ms": r["launch_ms"], "kernel_share_of_step": r["kernel_share_of_step"],
              "n_blocks": h["n_blocks"], "n_blocks_this_rank": h["n_blocks_this_rank"], "volume_gb": h["n_blocks"] * 65536 / 1e9,
              "inputs_gb": a.hd_frames * 1280 * 720 * 5 / 1e9, "n_upd_per_frame": h["n_upd_per_frame_all_ranks"],
              "achieved_hbm_gbs_whole_step": h["achieved_hbm_gbs_whole_step"], "synthesis_s": synth_s}
        for k in ("halo_exchange_ms", "halo_pieces_received_rank0", "extract_gather_ms", "gathered_points"):
            if k in h:
                hd[k] = h[k]
        del hseq

    if rank == 0:
        line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_of(a, world), "n_blocks": m["n_blocks"],
                "volume_mb": m["n_blocks"] * 65536 / 1e6,
                "roofline": m["roofline"], "cpu_baseline": cpu, "e2e": m.get("e2e"), "e2e_full": m.get("e2e_full"), "post_stage": post,
                "gpu_launches": m["gpu_launches"], "clocks": m["clocks"], "achieved_hbm_gbs_whole_step": m["achieved_hbm_gbs_whole_step"],
                "hd": hd}
        for k in ("halo_exchange_ms", "halo_pieces_received_rank0", "extract_gather_ms", "gathered_points", "gather_note"):
            if k in m:
                line[k] = m[k]
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    if a.slab_halo and a.slab_axis == 3:
        a.slab_axis = 0             # the replicated-halo mode exists for axis slabs only
    if a.hd and a.voxel == 0.005:
        a.voxel = 0.002
    if a.hd:
        a.hd_voxel = a.voxel
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
