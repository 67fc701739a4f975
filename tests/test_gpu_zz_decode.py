"""GPU decoders of the capture files (csrc/imgcodec.cu) through the C ABI: pixels equal the stock decoders'
(OpenCV = libpng / libjpeg-turbo, what the reference's o3d.io.read_image wraps) byte for byte, and the frame loop fed by them
builds the same volume as the host-decoded loop, with the same per-frame error semantics."""
import io
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frames():
    from otslam_b200 import synth
    seq = synth.make_sequence("chair_table", 40, subsample=(0, 4))          # 10 frames 640x480
    d, c = seq.numpy()
    return seq, d, c


def test_depth_png_and_colour_jpeg_equal_opencv(frames):
    from otslam_b200.decoder import FrameDecoder
    seq, dep, rgb = frames
    rng = np.random.default_rng(0)
    n = len(dep)
    png_params = [[], [cv2.IMWRITE_PNG_COMPRESSION, 0], [cv2.IMWRITE_PNG_COMPRESSION, 9],
                  [cv2.IMWRITE_PNG_COMPRESSION, 6, cv2.IMWRITE_PNG_STRATEGY, cv2.IMWRITE_PNG_STRATEGY_DEFAULT],
                  [cv2.IMWRITE_PNG_STRATEGY, cv2.IMWRITE_PNG_STRATEGY_FIXED]]
    jpg_params = [[], [cv2.IMWRITE_JPEG_QUALITY, 50], [cv2.IMWRITE_JPEG_QUALITY, 100], [cv2.IMWRITE_JPEG_OPTIMIZE, 1],
                  [cv2.IMWRITE_JPEG_RST_INTERVAL, 7], [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444],
                  [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422]]
    dfiles, cfiles, dref, cref = [], [], [], []
    for k in range(n):
        d = dep[k] if k % 2 == 0 else (dep[k].astype(np.int32) + rng.integers(-3, 4, dep[k].shape)).clip(0, 65535).astype(np.uint16)
        c = rgb[k] if k % 3 else rng.integers(0, 256, rgb[k].shape).astype(np.uint8)
        ok, e = cv2.imencode(".png", d, png_params[k % len(png_params)])
        dfiles.append(e.tobytes()); dref.append(d)
        ok, e = cv2.imencode(".jpg", c[..., ::-1], jpg_params[k % len(jpg_params)])
        cfiles.append(e.tobytes()); cref.append(cv2.imdecode(e, cv2.IMREAD_UNCHANGED)[..., ::-1])
    dec = FrameDecoder(480, 640, n)
    cs, ds = dec.decode_bytes(cfiles, dfiles)
    assert (cs == 0).all() and (ds == 0).all()
    gd, gc = dec.fetch(0, n)
    for k in range(n):
        assert (gd[k] == dref[k]).all(), f"depth {k}"
        assert (gc[k] == cref[k]).all(), f"colour {k}: max diff {np.abs(gc[k].astype(int) - cref[k]).max()}"
    prof = dec.profile()
    assert prof["compressed_bytes"] > 0 and prof["inflate_ms"] > 0 and prof["jpeg_huffman_ms"] > 0
    # a second, shorter batch through the same object; depth only
    cs, ds = dec.decode_bytes(None, dfiles[3:6])
    assert cs is None and (ds == 0).all()
    gd2, _ = dec.fetch(0, 3, rgb=False)
    assert (gd2 == np.stack(dref[3:6])).all()
    dec.close()


def test_png_colour_adaptive_filters_and_other_sizes():
    """gt_ / plain capture tools write PNG colour; Pillow's encoder uses all five filters (Paeth lines force one band to wait
    for the previous line).  Sizes that are not multiples of the MCU, 4:2:0."""
    from PIL import Image
    from otslam_b200.decoder import FrameDecoder
    rng = np.random.default_rng(1)
    H, W = 45, 83
    cols = [rng.integers(0, 256, (H, W, 3)).astype(np.uint8), np.kron(rng.integers(0, 256, (6, 11, 3)), np.ones((8, 8, 1))).astype(np.uint8)[:H, :W]]
    deps = [rng.integers(0, 65536, (H, W)).astype(np.uint16), (np.arange(H * W).reshape(H, W) % 5000).astype(np.uint16)]
    cfiles, dfiles = [], []
    for c, d in zip(cols, deps):
        bio = io.BytesIO(); Image.fromarray(c).save(bio, "PNG"); cfiles.append(bio.getvalue())
        bio = io.BytesIO(); Image.fromarray(d).save(bio, "PNG"); dfiles.append(bio.getvalue())
    ok, e = cv2.imencode(".jpg", cols[0][..., ::-1]); cfiles.append(e.tobytes())
    cols.append(cv2.imdecode(e, cv2.IMREAD_UNCHANGED)[..., ::-1])
    ok, e = cv2.imencode(".png", deps[0]); dfiles.append(e.tobytes()); deps.append(deps[0])
    rgba = np.dstack([cols[1], rng.integers(0, 256, (H, W)).astype(np.uint8)])
    bio = io.BytesIO(); Image.fromarray(rgba).save(bio, "PNG"); cfiles.append(bio.getvalue()); cols.append(cols[1])
    dfiles.append(dfiles[1]); deps.append(deps[1])
    dec = FrameDecoder(H, W, 8)
    cs, ds = dec.decode_bytes(cfiles, dfiles)
    assert (cs == 0).all() and (ds == 0).all(), (cs, ds)
    gd, gc = dec.fetch(0, 4)
    for k in range(4):
        assert (gd[k] == deps[k]).all() and (gc[k] == cols[k]).all(), k
    dec.close()


def test_status_codes_and_put(frames):
    from otslam_b200.decoder import CORRUPT, OK, UNSUPPORTED, FrameDecoder
    seq, dep, rgb = frames
    ok, png = cv2.imencode(".png", dep[0]); png = png.tobytes()
    ok, jpg = cv2.imencode(".jpg", rgb[0][..., ::-1]); jpg = jpg.tobytes()
    ok, prog = cv2.imencode(".jpg", rgb[0][..., ::-1], [cv2.IMWRITE_JPEG_PROGRESSIVE, 1]); prog = prog.tobytes()
    bad = bytearray(png); bad[300] ^= 0x10
    ok, small = cv2.imencode(".png", dep[0][:100, :100]); small = small.tobytes()
    cfiles = [jpg, prog, jpg, None, jpg[:200], b"not an image at all", jpg]
    dfiles = [png, png, bytes(bad), png, png, png[:4000], small]
    dec = FrameDecoder(480, 640, 8)
    cs, ds = dec.decode_bytes(cfiles, dfiles)
    assert list(cs) == [OK, UNSUPPORTED, OK, CORRUPT, CORRUPT, UNSUPPORTED, OK]
    assert list(ds) == [OK, OK, CORRUPT, OK, OK, CORRUPT, UNSUPPORTED]
    gd, gc = dec.fetch(0, 1)
    assert (gd[0] == dep[0]).all()
    # host-decoded arrays into a slot, and back
    dec.put(1, dep[1], rgb[1])
    gd, gc = dec.fetch(1, 1)
    assert (gd[0] == dep[1]).all() and (gc[0] == rgb[1]).all()
    dec.close()


def _volume(o3d, voxel=0.01):
    return o3d.pipelines.integration.ScalableTSDFVolume(voxel_length=voxel, sdf_trunc=4 * voxel,
                                                        color_type=o3d.pipelines.integration.TSDFVolumeColorType.RGB8)


def test_frame_loop_from_files_gpu_decode_equals_host_decode(frames, tmp_path, monkeypatch):
    """pipeline.integrate_files with the GPU decoders == the same loop with OpenCV on the host: identical blocks, weights,
    TSDF and colours; a damaged depth file, a progressive JPEG (stock decoder), a missing colour file and a broken pose in
    the middle keep the reference's print-and-skip / abort semantics (reconstruct_rgbd_filter.py:108-109)."""
    import otslam_b200.o3d_compat as o3d
    from otslam_b200 import capture, pipeline, synth
    seq, dep, rgb = frames
    base = str(tmp_path)
    n = len(dep)
    for k in range(n):
        capture.save_frame(base, "Object_0", k + 1, rgb[k], dep[k], seq.pose_ros[k])
    P = lambda sub, k, ext: os.path.join(base, sub, f"Object_0_{k}.{ext}")
    # frame 3: progressive JPEG (valid, not covered by the GPU decoder); frame 5: damaged PNG; frame 7: colour missing;
    # frame 8: pose with 15 numbers
    cv2.imwrite(P("color", 3, "jpg"), rgb[2][..., ::-1], [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    b = bytearray(open(P("depth", 5, "png"), "rb").read()); b[len(b) // 2] ^= 0xFF
    open(P("depth", 5, "png"), "wb").write(bytes(b))
    os.remove(P("color", 7, "jpg"))
    open(P("poses", 8, "txt"), "w").write(" ".join(["1.0"] * 15) + "\n")
    triples = [(P("color", k, "jpg"), P("depth", k, "png"), P("poses", k, "txt"), f"Object_0_{k}") for k in range(1, n + 1)]
    intr = o3d.camera.PinholeCameraIntrinsic(*seq.intr)
    monkeypatch.setattr(pipeline, "CHUNK_FRAMES", 4)              # three chunks: double-buffered decoders, holes in two of them

    def run(gpu):
        monkeypatch.setenv("OTSLAM_GPU_DECODE", "1" if gpu else "0")
        vol = _volume(o3d)
        errs, seen = [], []
        m = pipeline.integrate_files(vol, triples, intr, synth.T_FIX, skip_errors=True, on_error=lambda l, e: errs.append((l, str(e))),
                                     progress=lambda l, i, t: seen.append(l))
        return vol, m, errs, seen

    vh, mh, eh, sh = run(False)
    vg, mg, eg, sg = run(True)
    assert mh == mg == n - 3 and sh == sg
    assert [l for l, _ in eh] == [l for l, _ in eg] == ["Object_0_5", "Object_0_7", "Object_0_8"]
    assert [m for _, m in eh] == [m for _, m in eg]
    assert len(pipeline.last_decode_profile) == 3 and sum(c["passed_on"] for c in pipeline.last_decode_profile) == 3
    kh, th, wh, ch = vh._vol.export_blocks()
    kg, tg, wg, cg = vg._vol.export_blocks()
    assert (kh == kg).all() and (wh == wg).all() and (th == tg).all() and (ch == cg).all()
    # abort semantics (reconstruct_rgbd.py has no try): the first bad frame raises, nothing after it is integrated
    vol = _volume(o3d)
    with pytest.raises(Exception):
        pipeline.integrate_files(vol, triples, intr, synth.T_FIX)
    assert vol._vol.stats()["weight_sum"] > 0                               # chunk 0 (frames 1-4) went in before frame 5 failed


def test_hd_frames_gpu_decode(tmp_path):
    """configs[3] geometry: 1280x720 frames of the room through the GPU decoders == OpenCV."""
    from otslam_b200 import synth
    from otslam_b200.decoder import FrameDecoder
    seq = synth.make_sequence("room", 60, intr=synth.HD_INTRINSICS, subsample=(0, 20))
    dep, rgb = seq.numpy()
    dfiles = [cv2.imencode(".png", d)[1].tobytes() for d in dep]
    enc = [cv2.imencode(".jpg", c[..., ::-1])[1] for c in rgb]
    dec = FrameDecoder(720, 1280, len(dep))
    cs, ds = dec.decode_bytes([e.tobytes() for e in enc], dfiles)
    assert (cs == 0).all() and (ds == 0).all()
    gd, gc = dec.fetch(0, len(dep))
    assert (gd == dep).all()
    for k, e in enumerate(enc):
        assert (gc[k] == cv2.imdecode(e, cv2.IMREAD_UNCHANGED)[..., ::-1]).all()
    dec.close()
