"""GPU decoding of capture files: the readers in front of the frame loop.

The reference reads every frame with o3d.io.read_image (/root/reference/3d_model/reconstruct_rgbd.py:88-89) -- libjpeg /
libpng on one core.  FrameDecoder hands a chunk of files (paths, or bytes already in memory) to
`otslam_decoder_decode_files` / `otslam_decoder_decode`: host threads only read and frame the files, the compressed bytes
cross PCIe, and inflate / PNG filters / Huffman / IDCT / upsampling / colour conversion run on the GPU into frame slots
that `integrate` feeds to the volume without a host round trip.  Pixels equal the stock decoders' bit for bit
(tests/test_imgcodec_model.py on the CPU, tests/test_gpu_zz_decode.py on the GPU)."""
import ctypes as C
import os

import numpy as np

from . import _lib

OK, UNSUPPORTED, CORRUPT = 0, 1, 2      # per-file status


class FrameDecoder:
    def __init__(self, height, width, max_frames, device=0):
        self.height, self.width, self.max_frames, self.device = int(height), int(width), int(max_frames), int(device)
        self._h = C.c_void_p()
        _lib.check(_lib.lib.otslam_decoder_create(self.device, self.height, self.width, self.max_frames, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib.otslam_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:       # noqa: BLE001 -- interpreter shutdown
            pass

    @staticmethod
    def _paths(paths, n):
        if paths is None:
            return None
        assert len(paths) == n
        return (C.c_char_p * max(n, 1))(*[os.fsencode(p) for p in paths])

    def decode_files(self, color_paths, depth_paths):
        """Decode files into slots 0..n-1; returns (colour status, depth status) int32 arrays (None for a kind not given)."""
        n = len(depth_paths if depth_paths is not None else color_paths)
        cs = np.empty(n, np.int32) if color_paths is not None else None
        ds = np.empty(n, np.int32) if depth_paths is not None else None
        cp, dp = self._paths(color_paths, n), self._paths(depth_paths, n)
        _lib.check(_lib.lib.otslam_decoder_decode_files(self._h, n, cp, dp, _lib.ptr(cs), _lib.ptr(ds)))
        return cs, ds

    @staticmethod
    def _blob(files):
        if files is None:
            return None, None
        off = np.zeros(len(files) + 1, np.int64)
        np.cumsum([len(f) if f is not None else 0 for f in files], out=off[1:])
        blob = np.frombuffer(b"".join(f for f in files if f is not None) or b"\0", np.uint8)
        return blob, off

    def decode_bytes(self, color_files, depth_files):
        """The same for files already in memory: lists of bytes objects (None / b'' = missing)."""
        n = len(depth_files if depth_files is not None else color_files)
        cb, co = self._blob(color_files)
        db, do = self._blob(depth_files)
        cs = np.empty(n, np.int32) if color_files is not None else None
        ds = np.empty(n, np.int32) if depth_files is not None else None
        _lib.check(_lib.lib.otslam_decoder_decode(self._h, n, _lib.ptr(cb), _lib.ptr(co), _lib.ptr(db), _lib.ptr(do),
                                                  _lib.ptr(cs), _lib.ptr(ds)))
        return cs, ds

    def put(self, slot, depth=None, rgb=None):
        d = None if depth is None else np.ascontiguousarray(depth, np.uint16)
        c = None if rgb is None else np.ascontiguousarray(rgb, np.uint8)
        assert d is None or d.shape == (self.height, self.width)
        assert c is None or c.shape == (self.height, self.width, 3)
        _lib.check(_lib.lib.otslam_decoder_put(self._h, int(slot), _lib.ptr(d), _lib.ptr(c)))

    def fetch(self, first, count, depth=True, rgb=True):
        d = np.empty((count, self.height, self.width), np.uint16) if depth else None
        c = np.empty((count, self.height, self.width, 3), np.uint8) if rgb else None
        _lib.check(_lib.lib.otslam_decoder_fetch(self._h, int(first), int(count), _lib.ptr(d), _lib.ptr(c)))
        return d, c

    def integrate(self, volume, slots, intr, extrinsics, depth_scale=1000.0, depth_trunc=3.0, object_ids=None):
        """volume.integrate for the decoded slots listed (ascending), in that order; volume = volume.TSDFVolume."""
        s = np.ascontiguousarray(slots, np.int32)
        k = np.ascontiguousarray(intr, np.float64).reshape(4)
        e = np.ascontiguousarray(extrinsics, np.float64).reshape(len(s), 16)
        ids = None if object_ids is None else np.ascontiguousarray(object_ids, np.int32).reshape(len(s))
        _lib.check(_lib.lib.otslam_decoder_integrate(self._h, volume._h, len(s), _lib.ptr(s), _lib.ptr(k), _lib.ptr(e),
                                                     float(depth_scale), float(depth_trunc), _lib.ptr(ids)))

    def profile(self):
        """Device times of the last decode (CUDA events, ms) and the compressed bytes it uploaded."""
        out = np.zeros(6, np.float64)
        _lib.check(_lib.lib.otslam_decoder_profile(self._h, _lib.ptr(out)))
        return {"inflate_ms": out[0], "png_filter_emit_ms": out[1], "jpeg_huffman_ms": out[2], "jpeg_idct_ms": out[3],
                "jpeg_color_ms": out[4], "compressed_bytes": int(out[5])}
