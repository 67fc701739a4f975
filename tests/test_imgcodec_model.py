"""CPU model test of the capture-file decoders: csrc/imgcodec_core.h holds every arithmetic step of the
GPU decoders as __host__ __device__ functions; tests/models/imgcodec_model.cpp composes them the way the kernels do and is
compiled here with g++.  Oracle = the stock decoders the reference's o3d.io.read_image wraps (libpng, libjpeg with its
default islow IDCT + fancy upsampling), reached through OpenCV / Pillow.  Bar: every byte equal."""
import ctypes
import io
import os
import subprocess
import sys

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("imgcodec") / "imgcodec_model.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", so,
                           os.path.join(ROOT, "tests", "models", "imgcodec_model.cpp")])
    lib = ctypes.CDLL(so)
    lib.model_decode_png.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    lib.model_decode_jpeg.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return lib


def dec_png(lib, b, H, W, band=4, crc=1, team=1):
    """team > 1: the inflate step runs as a lock-step team of that many threads (the warp of the GPU kernel)."""
    buf = np.frombuffer(b, np.uint8)
    out = np.zeros(H * W * 4, np.uint8)
    ch = ctypes.c_int(0)
    st = lib.model_decode_png(buf.ctypes.data, len(b), H, W, band, crc, out.ctypes.data, ctypes.byref(ch), team)
    if st:
        return st, None
    if ch.value == 1:
        return 0, out[:H * W * 2].view(np.uint16).reshape(H, W)
    return 0, out[:H * W * 3].reshape(H, W, 3)


def dec_jpg(lib, b, H, W):
    buf = np.frombuffer(b, np.uint8)
    out = np.zeros((H, W, 3), np.uint8)
    return lib.model_decode_jpeg(buf.ctypes.data, len(b), H, W, out.ctypes.data), out


@pytest.fixture(scope="module")
def frames():
    from otslam_b200 import synth
    seq = synth.make_sequence("chair_table", 3)
    return seq.numpy()


def depth_cases(frames):
    rng = np.random.default_rng(1)
    dep = frames[0]
    noisy = (dep[2].astype(np.int32) + rng.integers(-3, 4, dep[2].shape)).clip(0, 65535).astype(np.uint16)   # sensor-like
    return [dep[0], noisy, rng.integers(0, 65536, (37, 53)).astype(np.uint16), np.zeros((16, 16), np.uint16)]


PNG_PARAMS = [[], [cv2.IMWRITE_PNG_COMPRESSION, 0], [cv2.IMWRITE_PNG_COMPRESSION, 9],
              [cv2.IMWRITE_PNG_COMPRESSION, 6, cv2.IMWRITE_PNG_STRATEGY, cv2.IMWRITE_PNG_STRATEGY_DEFAULT],
              [cv2.IMWRITE_PNG_STRATEGY, cv2.IMWRITE_PNG_STRATEGY_FIXED],
              [cv2.IMWRITE_PNG_STRATEGY, cv2.IMWRITE_PNG_STRATEGY_HUFFMAN_ONLY]]


def test_depth_png_equals_libpng(model, frames):
    """scanner_node.cpp:281 `cv::imwrite(depth_path, depth_u16)`: stored / fixed / dynamic deflate blocks, every zlib
    strategy OpenCV offers, and Pillow's adaptive filters (Up / Paeth lines); bands of 1, 4 and all scan lines."""
    from PIL import Image
    for d in depth_cases(frames):
        H, W = d.shape
        for params in PNG_PARAMS:
            ok, enc = cv2.imencode(".png", d, params)
            ref = cv2.imdecode(enc, cv2.IMREAD_UNCHANGED)
            for band, team in ((1, 1), (4, 3), (100000, 4)):
                st, o = dec_png(model, enc.tobytes(), H, W, band, team=team)
                assert st == 0 and o.dtype == ref.dtype and (o == ref).all() and (ref == d).all(), (params, band, team)
        bio = io.BytesIO()
        Image.fromarray(d).save(bio, "PNG")
        st, o = dec_png(model, bio.getvalue(), H, W, 4)
        assert st == 0 and (o == d).all()


def test_colour_png_equals_libpng(model, frames):
    """gt_color_%04d.png / color_%04d.png of the older capture tools (rgbd_capture_node_gt.cpp:126): RGB8 and RGBA8."""
    from PIL import Image
    rng = np.random.default_rng(2)
    for c in [frames[1][0], rng.integers(0, 256, (33, 47, 3)).astype(np.uint8)]:
        H, W, _ = c.shape
        for opt in (False, True):
            bio = io.BytesIO()
            Image.fromarray(c).save(bio, "PNG", optimize=opt)
            for band in (1, 4, 100000):
                st, o = dec_png(model, bio.getvalue(), H, W, band)
                assert st == 0 and (o == c).all()
        ok, enc = cv2.imencode(".png", c[..., ::-1])
        st, o = dec_png(model, enc.tobytes(), H, W)
        assert st == 0 and (o == c).all()
        rgba = np.dstack([c, rng.integers(0, 256, (H, W)).astype(np.uint8)])
        bio = io.BytesIO()
        Image.fromarray(rgba).save(bio, "PNG")
        st, o = dec_png(model, bio.getvalue(), H, W)
        assert st == 0 and (o == c).all()          # alpha dropped, as the frame loop's cvtColor(BGRA2RGB) does


def test_png_status_codes(model, frames):
    d = frames[0][0]
    ok, enc = cv2.imencode(".png", d)
    b = enc.tobytes()
    bad = bytearray(b)
    bad[200] ^= 0x40
    assert dec_png(model, bytes(bad), 480, 640)[0] == 2            # CRC of a critical chunk
    assert dec_png(model, b[:5000], 480, 640)[0] == 2              # truncated
    assert dec_png(model, b"\x89PNG\r\n\x1a\n" + b"\0" * 40, 480, 640)[0] == 2
    assert dec_png(model, b, 481, 640)[0] == 1                     # another size: the stock path raises the format error
    ok, enc8 = cv2.imencode(".png", (d >> 8).astype(np.uint8))
    assert dec_png(model, enc8.tobytes(), 480, 640)[0] == 1        # 8-bit grey: not a depth file of this path
    from PIL import Image
    bio = io.BytesIO()
    Image.fromarray(frames[1][0]).convert("P").save(bio, "PNG")
    assert dec_png(model, bio.getvalue(), 480, 640)[0] == 1        # palette


JPEG_PARAMS = [[], [cv2.IMWRITE_JPEG_QUALITY, 50], [cv2.IMWRITE_JPEG_QUALITY, 100], [cv2.IMWRITE_JPEG_QUALITY, 5],
               [cv2.IMWRITE_JPEG_OPTIMIZE, 1], [cv2.IMWRITE_JPEG_RST_INTERVAL, 3],
               [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444],
               [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422],
               [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_RST_INTERVAL, 1,
                cv2.IMWRITE_JPEG_OPTIMIZE, 1, cv2.IMWRITE_JPEG_QUALITY, 97]]


def colour_cases(frames):
    rng = np.random.default_rng(3)
    blocks = np.kron(rng.integers(0, 256, (9, 11, 3)), np.ones((8, 8, 1))).astype(np.uint8)[:67, :85]
    return [frames[1][0], frames[1][1], rng.integers(0, 256, (480, 640, 3)).astype(np.uint8),
            rng.integers(0, 256, (45, 83, 3)).astype(np.uint8), rng.integers(0, 256, (16, 16, 3)).astype(np.uint8), blocks]


def test_colour_jpeg_equals_libjpeg(model, frames):
    """scanner_node.cpp:275 `cv::imwrite(color_path, rgb)` (baseline, 4:2:0, quality 95, standard tables) and the other
    baseline variants OpenCV can write: qualities 5..100, optimised Huffman tables, restart intervals, 4:2:2, 4:4:4, sizes
    that are not multiples of the MCU.  Every RGB byte equals libjpeg-turbo's."""
    for c in colour_cases(frames):
        H, W, _ = c.shape
        for params in JPEG_PARAMS:
            ok, enc = cv2.imencode(".jpg", c[..., ::-1], params)
            ref = cv2.imdecode(enc, cv2.IMREAD_UNCHANGED)[..., ::-1]
            st, o = dec_jpg(model, enc.tobytes(), H, W)
            assert st == 0, (params, c.shape)
            assert (o == ref).all(), (params, c.shape, int(np.abs(o.astype(int) - ref).max()))


def test_jpeg_status_codes(model, frames):
    c = frames[1][0]
    ok, enc = cv2.imencode(".jpg", c[..., ::-1], [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    assert dec_jpg(model, enc.tobytes(), 480, 640)[0] == 1          # progressive: the stock decoder's business
    ok, enc = cv2.imencode(".jpg", c[..., 0])
    assert dec_jpg(model, enc.tobytes(), 480, 640)[0] == 1          # grey
    ok, enc = cv2.imencode(".jpg", c[..., ::-1])
    b = enc.tobytes()
    assert dec_jpg(model, b, 480, 641)[0] == 1
    assert dec_jpg(model, b[:300], 480, 640)[0] == 2                # cut inside the tables
    assert dec_jpg(model, b"\xff\xd8\xff\xd9", 480, 640)[0] == 2
    # a scan cut short still decodes (missing bits read as zeros, like libjpeg's warning path) and never reads past the data
    st, o = dec_jpg(model, b[:len(b) // 2], 480, 640)
    assert st == 0
    ref = cv2.imdecode(np.frombuffer(b[:len(b) // 2], np.uint8), cv2.IMREAD_UNCHANGED)
    if ref is not None:                                              # rows decoded before the cut agree
        assert (o[:200] == ref[:200, :, ::-1]).all()


def test_decoders_survive_garbage(model):
    """Bounded loops on hostile input: random bytes after valid headers must return, not hang or crash."""
    rng = np.random.default_rng(4)
    c = rng.integers(0, 256, (64, 64, 3)).astype(np.uint8)
    ok, enc = cv2.imencode(".jpg", c)
    b = bytearray(enc.tobytes())
    for trial in range(50):
        g = bytearray(b)
        for _ in range(20):
            g[int(rng.integers(2, len(g)))] = int(rng.integers(0, 256))
        dec_jpg(model, bytes(g), 64, 64)
    ok, enc = cv2.imencode(".png", rng.integers(0, 65536, (64, 64)).astype(np.uint16))
    b = bytearray(enc.tobytes())
    for trial in range(50):
        g = bytearray(b)
        for _ in range(5):
            g[int(rng.integers(40, len(g) - 12))] = int(rng.integers(0, 256))
        assert dec_png(model, bytes(g), 64, 64, crc=0, team=1 + trial % 3)[0] in (0, 2)       # (77 = the lanes of a team disagreed)


def test_colour_jpeg_second_encoder_and_second_decoder(model):
    """Files written by Pillow's encoder (its own libjpeg build; subsampling 4:4:4 / 4:2:2 / 4:2:0, optimised tables) and
    decoded by BOTH stock decoders in this image -- OpenCV's libjpeg-turbo and Pillow's -- agree with the model byte for
    byte: the pin does not hang on one library build."""
    from PIL import Image
    rng = np.random.default_rng(5)
    for H, W in [(48, 64), (45, 83), (100, 17), (33, 200)]:
        base = np.kron(rng.integers(0, 256, (H // 8 + 1, W // 8 + 1, 3)), np.ones((8, 8, 1)))[:H, :W]
        img = (base * 0.7 + rng.integers(0, 80, (H, W, 3))).clip(0, 255).astype(np.uint8)
        for sub in (0, 1, 2):
            for q, opt in ((30, False), (75, True), (95, False)):
                bio = io.BytesIO()
                Image.fromarray(img).save(bio, "JPEG", quality=q, subsampling=sub, optimize=opt)
                b = bio.getvalue()
                st, o = dec_jpg(model, b, H, W)
                ref = cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_UNCHANGED)[..., ::-1]
                pil = np.asarray(Image.open(io.BytesIO(b)).convert("RGB"))
                assert st == 0 and (o == ref).all() and (o == pil).all(), (H, W, sub, q, opt)


def _png_bytes(img, filters, level=6):
    """A PNG written here: scan line r filtered with filters[r] (PNG spec 9.2), zlib level `level`, one IDAT per 1000 bytes."""
    import struct
    import zlib
    if img.dtype == np.uint16:
        raw, ctype, depth, bpp = img.astype(">u2").view(np.uint8).reshape(img.shape[0], -1), 0, 16, 2
    else:
        raw, ctype, depth, bpp = img.reshape(img.shape[0], -1), (2 if img.shape[2] == 3 else 6), 8, img.shape[2]
    H, rb = raw.shape
    lines = []
    prev = np.zeros(rb, np.int32)
    for r in range(H):
        cur = raw[r].astype(np.int32)
        a = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        b = prev
        c = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        ft = int(filters[r])
        if ft == 0:
            pred = 0
        elif ft == 1:
            pred = a
        elif ft == 2:
            pred = b
        elif ft == 3:
            pred = (a + b) >> 1
        else:
            p = a + b - c
            pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))
        lines.append(bytes([ft]) + ((cur - pred) & 255).astype(np.uint8).tobytes())
        prev = cur
    z = zlib.compress(b"".join(lines), level)

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", img.shape[1], H, depth, ctype, 0, 0, 0))
    for i in range(0, len(z), 1000):
        out += chunk(b"IDAT", z[i:i + 1000])
    return out + chunk(b"IEND", b"")


def test_png_every_filter_in_every_position(model):
    """Scan lines filtered with random filter types (first line included: Up / Average / Paeth against an all-zero line
    above), long runs of Paeth (bands that must wait), zlib levels 0 / 1 / 9: the band-parallel reconstruction equals
    libpng's for every band height and team size."""
    rng = np.random.default_rng(7)
    cases = [(rng.integers(0, 65536, (41, 29)).astype(np.uint16), rng.integers(0, 5, 41)),
             (rng.integers(0, 256, (37, 23, 3)).astype(np.uint8), rng.integers(0, 5, 37)),
             (rng.integers(0, 256, (19, 31, 4)).astype(np.uint8), np.full(19, 4)),                 # all Paeth: one band does it all
             ((np.arange(64 * 48).reshape(48, 64) * 7 % 65536).astype(np.uint16), np.array([4, 3, 2] + [1, 0, 2, 3, 4] * 9))]
    for img, filters in cases:
        H, W = img.shape[:2]
        for level in (0, 1, 9):
            b = _png_bytes(img, filters, level)
            ref = cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_UNCHANGED)
            want = img if img.ndim == 2 else img[..., :3]
            assert ref is not None and (ref if img.ndim == 2 else ref[..., :3][..., ::-1]).shape == want.shape
            assert ((ref if img.ndim == 2 else ref[..., :3][..., ::-1]) == want).all()                 # libpng accepts the file
            for band, team in ((1, 1), (3, 2), (4, 4), (1000, 3)):
                st, o = dec_png(model, b, H, W, band, team=team)
                assert st == 0 and (o == want).all(), (img.shape, level, band, team)
    # a filter byte > 4 is an error (libpng: "bad adaptive filter value")
    img, filters = cases[0]
    bad = filters.copy()
    bad[5] = 7
    assert dec_png(model, _png_bytes(img, bad), 41, 29)[0] == 2


def test_hd_frame_pair_equals_stock_decoders(model):
    """configs[3] geometry: one 1280x720 frame pair of the large room as cv::imwrite stores it."""
    from otslam_b200 import synth
    seq = synth.make_sequence("room", 40, intr=synth.HD_INTRINSICS, subsample=(0, 20))
    dep, rgb = seq.numpy()
    for k in range(len(dep)):
        ok, e = cv2.imencode(".jpg", rgb[k][..., ::-1])
        st, o = dec_jpg(model, e.tobytes(), 720, 1280)
        assert st == 0 and (o == cv2.imdecode(e, cv2.IMREAD_UNCHANGED)[..., ::-1]).all()
        ok, e = cv2.imencode(".png", dep[k])
        st, o = dec_png(model, e.tobytes(), 720, 1280, 4, team=4)
        assert st == 0 and (o == dep[k]).all()
