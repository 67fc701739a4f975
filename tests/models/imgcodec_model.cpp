// CPU model of csrc/imgcodec.cu: the same __host__ __device__ functions (csrc/imgcodec_core.h) composed the way the
// kernels compose them, compiled with g++ by tests/test_imgcodec_model.py and compared with OpenCV's decoders.
// Test infrastructure only -- the product library never runs these functions on the host.
#include <pthread.h>
#include <stdlib.h>

#include <thread>
#include <vector>

// the decoder's team barrier (a warp's __syncwarp() on the GPU) as a thread barrier, so that teams of several host threads
// exercise the same synchronisation points
static thread_local pthread_barrier_t* g_team_barrier = nullptr;
#define IC_TEAM_SYNC()                                              \
    do {                                                            \
        if (g_team_barrier) pthread_barrier_wait(g_team_barrier);   \
    } while (0)

#include "../../object-triggered-3d-slam_b200/csrc/imgcodec_core.h"

using namespace imgcodec;

// inflate with a team of `team` threads in lock step (team == 1: plain call)
static int team_inflate(const uint8_t* z, int64_t z_len, uint8_t* raw, int64_t cap, int team, int64_t* got) {
    InflateTables* T = new InflateTables;
    int st = IC_OK;
    if (team <= 1) {
        st = inflate_zlib(z, z_len, raw, cap, *T, got);
    } else {
        pthread_barrier_t bar;
        pthread_barrier_init(&bar, nullptr, (unsigned)team);
        std::vector<int> sts((size_t)team, IC_OK);
        std::vector<int64_t> gots((size_t)team, 0);
        std::vector<std::thread> th;
        for (int l = 0; l < team; ++l)
            th.emplace_back([&, l] {
                g_team_barrier = &bar;
                sts[l] = inflate_zlib(z, z_len, raw, cap, *T, &gots[l], l, team);
                g_team_barrier = nullptr;
            });
        for (auto& t : th) t.join();
        pthread_barrier_destroy(&bar);
        st = sts[0];
        *got = gots[0];
        for (int l = 1; l < team; ++l)
            if (sts[l] != st || gots[l] != *got) st = 77;          // the lanes must agree
    }
    delete T;
    return st;
}

extern "C" int model_decode_png(const uint8_t* file, int64_t size, int H, int W, int band_rows, int check_crc, uint8_t* out,
                                int* channels, int team) {
    static const Crc32 crc;
    std::vector<uint8_t> z((size_t)size);
    PngFrame f;
    if (png_parse(file, size, H, W, z.data(), size, check_crc ? &crc : nullptr, f) != IC_OK) return f.status;
    *channels = f.channels;
    const int64_t cap = (int64_t)(f.rowbytes + 1) * f.height;
    std::vector<uint8_t> raw((size_t)cap);
    int64_t got = 0;
    int st = team_inflate(z.data(), f.z_len, raw.data(), cap, team, &got);
    if (st != IC_OK) return st;
    if (got != cap) return IC_CORRUPT;                           // "Not enough image data"
    if (band_rows < 1) band_rows = 1;
    const int n_bands = (f.height + band_rows - 1) / band_rows;
    // bands in reverse order on purpose: they must be independent of each other
    std::vector<int> first(n_bands + 1);
    for (int b = 0; b <= n_bands; ++b) first[b] = b == n_bands ? f.height : png_band_first(raw.data(), f.height, f.rowbytes + 1, b * band_rows);
    for (int b = n_bands - 1; b >= 0; --b)
        if (png_unfilter_band(raw.data(), f.height, f.rowbytes, f.bpp, first[b], first[b + 1]) != IC_OK) st = IC_CORRUPT;
    if (st != IC_OK) return st;
    for (int y = 0; y < f.height; ++y)
        for (int x = 0; x < f.width; ++x) png_emit_pixel(raw.data(), f, x, y, out);
    return IC_OK;
}

extern "C" int model_decode_jpeg(const uint8_t* file, int64_t size, int H, int W, uint8_t* rgb) {
    JpegFrame f;
    JpegHeader hd;
    if (jpeg_parse(file, size, H, W, f, hd) != IC_OK) return f.status;
    JpegTables* T = new JpegTables();
    bool ok = true;
    for (int s = 0; s < 4; ++s)
        if (hd.h_set[s]) ok = ok && jpeg_build_huff(hd.hbits[s], hd.hvals[s], hd.hn[s], s < 2, T->h[s]);
    for (int t = 0; t < 4; ++t)
        for (int k = 0; k < 64; ++k) T->q[t][k] = hd.q[t][k];
    if (!ok) { delete T; return IC_CORRUPT; }
    for (int a = 0; a < 2; ++a)
        if (hd.h_set[2 + a]) jpeg_build_acfast(T->h[2 + a], T->acfast[a]);
    std::vector<int16_t> coef((size_t)f.n_blocks * 64, 0);
    std::vector<uint8_t> samples((size_t)f.n_blocks * 64);
    f.blk_off = 0;
    int st = jpeg_decode_scan(f, *T, file + f.scan_off, coef.data());
    if (st != IC_OK) { delete T; return st; }
    for (int c = 0; c < 3; ++c) {
        uint8_t* plane = samples.data() + (size_t)f.blk_base[c] * 64;
        const int stride = f.wblk[c] * 8;
        for (int by = 0; by < f.hblk[c]; ++by)
            for (int bx = 0; bx < f.wblk[c]; ++bx) {
                const int16_t* blk = coef.data() + ((size_t)f.blk_base[c] + (size_t)by * f.wblk[c] + bx) * 64;
                int32_t ws[8][8];                                 // [column][row]
                for (int col = 0; col < 8; ++col) idct_column(blk, T->q[f.tq[c]], col, ws[col]);
                for (int r = 0; r < 8; ++r) {
                    int32_t row[8];
                    for (int col = 0; col < 8; ++col) row[col] = ws[col][r];
                    uint8_t o[8];
                    idct_row(row, o);
                    for (int col = 0; col < 8; ++col) plane[(size_t)(by * 8 + r) * stride + bx * 8 + col] = o[col];
                }
            }
    }
    JpegPlanes P;
    P.y = samples.data() + (size_t)f.blk_base[0] * 64;
    P.cb = samples.data() + (size_t)f.blk_base[1] * 64;
    P.cr = samples.data() + (size_t)f.blk_base[2] * 64;
    P.ys = f.wblk[0] * 8;
    P.cs = f.wblk[1] * 8;
    P.hmax = f.hmax;
    P.vmax = f.vmax;
    P.cw = (f.width + f.hmax - 1) / f.hmax;
    P.ch = (f.height + f.vmax - 1) / f.vmax;
    for (int y = 0; y < f.height; ++y)
        for (int x = 0; x < f.width; ++x) jpeg_pixel_rgb(P, x, y, rgb + ((size_t)y * f.width + x) * 3);
    delete T;
    return IC_OK;
}
