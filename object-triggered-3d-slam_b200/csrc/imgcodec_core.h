// imgcodec_core.h -- decoders for the two file formats of a capture tree, written once as
// __host__ __device__ code: depth = 16-bit grey PNG, colour = baseline JPEG (or 8-bit RGB / RGBA PNG for the gt_ / plain
// capture tools), the bytes ScannerNode::save_files produces with cv::imwrite
// (/root/reference/ros2_ws/src/system_manager/src/scanner_node.cpp:268-283) and the reference reads back with
// o3d.io.read_image (/root/reference/3d_model/reconstruct_rgbd.py:88-89).
//
// The kernels in imgcodec.cu are thin wrappers around these functions; tests/imgcodec_model.cpp compiles the same
// functions with g++ so that the CPU suite checks them, byte for byte, against the stock decoders (OpenCV's libpng /
// libjpeg-turbo) without a GPU.  Nothing here is a product CPU path: the library only ever calls it from kernels
// (the container parsing -- chunk / marker walking, table set-up -- is host code, as it is I/O, not arithmetic).
//
// What "bit-exact" means here:
//   * PNG is lossless: inflate (RFC 1951) + the five scan-line filters (PNG spec 9.2) have one right answer.
//   * JPEG: the decoder restates the arithmetic libjpeg / libjpeg-turbo use by default -- Huffman decoding (ITU T.81 F.2),
//     the "islow" 13-bit fixed-point inverse DCT, the triangle-filter ("fancy") chroma upsampling for 4:2:0 / 4:2:2 and
//     the 16-bit fixed-point YCbCr -> RGB conversion -- so the RGB bytes equal cv2.imread's on files an encoder wrote.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define IC_FN __host__ __device__
#define IC_DEV_TABLE(type, name, n, ...) static __device__ const type name##_dev[n] = __VA_ARGS__;
#else
#define IC_FN
#define IC_DEV_TABLE(type, name, n, ...)
#endif
#if defined(__CUDA_ARCH__)
#define IC_TABLE(name) name##_dev
#else
#define IC_TABLE(name) name##_host
#endif
#define IC_DEFINE_TABLE(type, name, n, ...)             \
    static const type name##_host[n] = __VA_ARGS__;     \
    IC_DEV_TABLE(type, name, n, __VA_ARGS__)

namespace imgcodec {

enum { IC_OK = 0, IC_UNSUPPORTED = 1, IC_CORRUPT = 2 };

// =====================================================================================================================
// inflate (RFC 1950 / 1951)
// =====================================================================================================================
constexpr int kLitFast = 10;      // literal/length codes up to 10 bits resolve with one table lookup
constexpr int kDistFast = 8;

struct InflateTables {            // 3.6 KB: one per decoder (shared memory on the GPU)
    uint16_t lit_fast[1 << kLitFast];      // (symbol << 4) | code length, 0 = longer code
    uint16_t dist_fast[1 << kDistFast];
    uint16_t lit_count[16], lit_symbol[288];
    uint16_t dist_count[16], dist_symbol[32];
    uint16_t len_count[16], len_symbol[19];
    uint16_t lengths[320];
    int32_t build_status;             // lane 0's verdict on a table set, read by the whole team after the barrier
};

struct BitsLSB {                  // deflate packs bits starting at the least significant bit of each byte
    const uint8_t* p;
    const uint8_t* end;
    uint64_t buf;
    int cnt;                      // valid bits in buf; negative = the stream ended early
    IC_FN void refill() {         // afterwards cnt > 32 unless the input is exhausted
        if (cnt > 32) return;
        if (end - p >= 4) {
            const uint32_t w = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
            buf |= (uint64_t)w << cnt;
            cnt += 32;
            p += 4;
        } else {
            while (cnt <= 56 && p < end) {
                buf |= (uint64_t)(*p++) << cnt;
                cnt += 8;
            }
        }
    }
    IC_FN void drop(int n) {
        buf >>= n;
        cnt -= n;
    }
    IC_FN uint32_t bits(int n) {  // n <= 16
        const uint32_t v = (uint32_t)buf & ((1u << n) - 1u);
        drop(n);
        return v;
    }
};

IC_FN inline uint32_t bit_reverse(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) {
        r = (r << 1) | (code & 1u);
        code >>= 1;
    }
    return r;
}

// Canonical Huffman code from code lengths.  Returns 0 for a complete code, > 0 for an incomplete one (that many
// codes of the longest length unused), < 0 for an over-subscribed one.
IC_FN inline int huff_construct(uint16_t* count, uint16_t* symbol, const uint16_t* length, int n, uint16_t* fast, int fast_bits) {
    for (int l = 0; l < 16; ++l) count[l] = 0;
    for (int s = 0; s < n; ++s) count[length[s]]++;
    if (fast)
        for (int i = 0; i < (1 << fast_bits); ++i) fast[i] = 0;
    if (count[0] == n) return 0;
    int left = 1;
    for (int l = 1; l < 16; ++l) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return left;
    }
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    for (int s = 0; s < n; ++s)
        if (length[s]) symbol[offs[length[s]]++] = (uint16_t)s;
    if (fast) {
        uint32_t code = 0;
        int idx = 0;
        for (int l = 1; l <= fast_bits; ++l) {
            for (int k = 0; k < count[l]; ++k, ++idx, ++code) {
                const uint16_t e = (uint16_t)((symbol[idx] << 4) | l);
                for (uint32_t j = bit_reverse(code, l); j < (1u << fast_bits); j += (1u << l)) fast[j] = e;
            }
            code <<= 1;
        }
    }
    return left;
}

// One symbol; -1 = no such code.  The caller guarantees >= 15 bits in the buffer (or an exhausted input).
IC_FN inline int huff_decode(BitsLSB& b, const uint16_t* fast, int fast_bits, const uint16_t* count, const uint16_t* symbol) {
    if (fast) {
        const uint32_t e = fast[(uint32_t)b.buf & ((1u << fast_bits) - 1u)];
        if (e) {
            b.drop((int)(e & 15u));
            return (int)(e >> 4);
        }
    }
    int code = 0, first = 0, index = 0;
    uint32_t bits = (uint32_t)b.buf;
    for (int l = 1; l <= 15; ++l) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int c = count[l];
        if (code - c < first) {
            b.drop(l);
            return symbol[index + (code - first)];
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// ---- a TEAM of `nlanes` threads runs the decoder in lock step (on the GPU: the 32 lanes of a warp; on the host: 1).  Every
// lane decodes the same symbols from the same bytes (uniform control flow, table reads are broadcasts), so no lane ever
// waits for another to learn what comes next; the work that is parallel -- copying a match of up to 258 bytes, a stored
// block -- is split across the lanes (lane l moves bytes l, l + nlanes, ...).  Tables are built by lane 0 between two team
// barriers.  IC_TEAM_SYNC() = the barrier (orders the team's memory accesses): __syncwarp() in kernels, nothing for a
// team of one; the CPU model defines it as a thread barrier to run teams of several threads.
#ifndef IC_TEAM_SYNC
#if defined(__CUDA_ARCH__)
#define IC_TEAM_SYNC() __syncwarp()
#else
#define IC_TEAM_SYNC() ((void)0)
#endif
#endif

IC_FN inline int inflate_codes(BitsLSB& b, uint8_t* out, int64_t& pos64, int64_t cap64, const InflateTables& T, bool& full, int lane,
                               int nlanes) {
    // The loop is tuned for literals (98 % of the symbols of sensor-like depth, ncu profiles/decode_r02a.md): one table
    // load and one compare recognise "a literal whose code fits the fast table", and one refill (> 32 bits) feeds up to
    // three of them (<= 10 bits each) without further refill or bounds tests.  The input-underrun test is left to the
    // paths that end a run of literals (a run is bounded by `cap` in any case).  32-bit positions (an image's scan lines
    // are < 2^31 bytes: the decoder object caps height x width).
    const uint16_t* const lit_fast = T.lit_fast;
    const uint32_t fmask = (1u << kLitFast) - 1u;
    int pos = (int)pos64;
    const int cap = (int)cap64;
    int status = IC_OK;
    for (;;) {
        b.refill();
        uint32_t e = lit_fast[(uint32_t)b.buf & fmask];
        if (e - 1u < (256u << 4) - 1u) {                     // 1 <= e < 4096: fast entry, symbol < 256
            if (pos + 3 > cap) {                             // the last bytes of the image: one at a time
                if (pos >= cap) { full = true; break; }
                b.drop((int)(e & 15u));
                if (lane == 0) out[pos] = (uint8_t)(e >> 4);
                ++pos;
                continue;
            }
            b.drop((int)(e & 15u));
            if (lane == 0) out[pos] = (uint8_t)(e >> 4);
            ++pos;
            e = lit_fast[(uint32_t)b.buf & fmask];
            if (e - 1u >= (256u << 4) - 1u) continue;        // not a short literal: start over (refill, general path)
            b.drop((int)(e & 15u));
            if (lane == 0) out[pos] = (uint8_t)(e >> 4);
            ++pos;
            e = lit_fast[(uint32_t)b.buf & fmask];
            if (e - 1u >= (256u << 4) - 1u) continue;
            b.drop((int)(e & 15u));
            if (lane == 0) out[pos] = (uint8_t)(e >> 4);
            ++pos;
            continue;
        }
        int sym;
        if (e) {
            b.drop((int)(e & 15u));
            sym = (int)(e >> 4);
        } else {
            sym = huff_decode(b, nullptr, 0, T.lit_count, T.lit_symbol);
        }
        if (sym < 0 || b.cnt < 0) { status = IC_CORRUPT; break; }
        if (sym < 256) {                                     // a literal with a long code
            if (pos >= cap) { full = true; break; }
            if (lane == 0) out[pos] = (uint8_t)sym;
            ++pos;
            continue;
        }
        if (sym == 256) break;
        sym -= 257;
        if (sym >= 29) { status = IC_CORRUPT; break; }
        int len;
        if (sym < 8) len = 3 + sym;
        else if (sym == 28) len = 258;
        else {
            const int eb = (sym >> 2) - 1;
            len = 3 + ((4 + (sym & 3)) << eb) + (int)b.bits(eb);
        }
        b.refill();
        const int ds = huff_decode(b, T.dist_fast, kDistFast, T.dist_count, T.dist_symbol);
        if (ds < 0 || ds >= 30) { status = IC_CORRUPT; break; }
        int dist;
        if (ds < 4) dist = 1 + ds;
        else {
            const int eb = (ds >> 1) - 1;
            dist = 1 + ((2 + (ds & 1)) << eb) + (int)b.bits(eb);
        }
        if (b.cnt < 0 || dist > pos) { status = IC_CORRUPT; break; }
        if (len > cap - pos) { len = cap - pos; full = true; }
        uint8_t* dst = out + pos;
        const uint8_t* src = dst - dist;
        IC_TEAM_SYNC();                                      // everything before `pos` is written, by whichever lane
        if (dist >= len) {
            for (int i = lane; i < len; i += nlanes) dst[i] = src[i];
        } else {                                             // the match overlaps itself: the last `dist` bytes repeat
            for (int i = lane; i < len; i += nlanes) dst[i] = src[i % dist];
        }
        pos += len;
        if (full) break;
    }
    if (b.cnt < 0) status = IC_CORRUPT;                      // the stream ended inside a run of literals
    pos64 = pos;
    return status;
}

// zlib stream -> out[0..cap).  Stops when `cap` bytes exist (a PNG decoder ignores what follows the last scan line).
// *out_len = bytes produced.  The Adler-32 trailer is not verified (documented in DESIGN.md).  Every lane of the team calls
// this with the same arguments and gets the same result; out[] is complete after a final IC_TEAM_SYNC() by the caller.
IC_FN inline int inflate_zlib(const uint8_t* in, int64_t in_len, uint8_t* out, int64_t cap, InflateTables& T, int64_t* out_len,
                              int lane = 0, int nlanes = 1) {
    *out_len = 0;
    if (in_len < 2) return IC_CORRUPT;
    const int cmf = in[0], flg = in[1];
    if ((cmf & 15) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20)) return IC_CORRUPT;
    BitsLSB b{in + 2, in + in_len, 0, 0};
    int64_t pos = 0;
    bool full = false;
    int last;
    do {
        b.refill();
        last = (int)b.bits(1);
        const int type = (int)b.bits(2);
        if (b.cnt < 0) return IC_CORRUPT;
        if (type == 0) {
            b.drop(b.cnt & 7);
            b.p -= b.cnt >> 3;                                // hand the whole bytes still buffered back
            b.buf = 0;
            b.cnt = 0;
            if (b.end - b.p < 4) return IC_CORRUPT;
            const uint32_t len = (uint32_t)b.p[0] | ((uint32_t)b.p[1] << 8), nlen = (uint32_t)b.p[2] | ((uint32_t)b.p[3] << 8);
            b.p += 4;
            if (len != (~nlen & 0xFFFFu) || (int64_t)len > b.end - b.p) return IC_CORRUPT;
            int64_t n = len;
            if (n > cap - pos) { n = cap - pos; full = true; }
            for (int64_t i = lane; i < n; i += nlanes) out[pos + i] = b.p[i];
            pos += n;
            b.p += len;
        } else if (type == 1) {
            IC_TEAM_SYNC();                                   // nobody still decodes with the previous block's tables
            if (lane == 0) {
                int s = 0;
                for (; s < 144; ++s) T.lengths[s] = 8;
                for (; s < 256; ++s) T.lengths[s] = 9;
                for (; s < 280; ++s) T.lengths[s] = 7;
                for (; s < 288; ++s) T.lengths[s] = 8;
                huff_construct(T.lit_count, T.lit_symbol, T.lengths, 288, T.lit_fast, kLitFast);
                for (s = 0; s < 30; ++s) T.lengths[s] = 5;
                huff_construct(T.dist_count, T.dist_symbol, T.lengths, 30, T.dist_fast, kDistFast);
            }
            IC_TEAM_SYNC();
            const int r = inflate_codes(b, out, pos, cap, T, full, lane, nlanes);
            if (r != IC_OK) return r;
        } else if (type == 2) {
            b.refill();
            const int nlen = (int)b.bits(5) + 257, ndist = (int)b.bits(5) + 1, ncode = (int)b.bits(4) + 4;
            if (b.cnt < 0 || nlen > 286 || ndist > 30) return IC_CORRUPT;
            // order of the code-length code lengths: 16 17 18 0 8 7 9 6 10 5 11 4 | 12 3 13 2 14 1 15 (5 bits each)
            const uint64_t lo = 16ull | (17ull << 5) | (18ull << 10) | (0ull << 15) | (8ull << 20) | (7ull << 25) | (9ull << 30) |
                                (6ull << 35) | (10ull << 40) | (5ull << 45) | (11ull << 50) | (4ull << 55);
            const uint64_t hi = 12ull | (3ull << 5) | (13ull << 10) | (2ull << 15) | (14ull << 20) | (1ull << 25) | (15ull << 30);
            // the 19 lengths are small: every lane keeps its own copy in registers / local memory until lane 0 publishes them
            uint16_t cl[19];
            for (int i = 0; i < 19; ++i) cl[i] = 0;
            for (int i = 0; i < ncode; ++i) {
                b.refill();
                const int ord = (int)((i < 12 ? lo >> (5 * i) : hi >> (5 * (i - 12))) & 31u);
                cl[ord] = (uint16_t)b.bits(3);
            }
            if (b.cnt < 0) return IC_CORRUPT;
            IC_TEAM_SYNC();                                   // the previous block's tables (and T.lengths) are no longer read
            if (lane == 0) {
                for (int i = 0; i < 19; ++i) T.lengths[i] = cl[i];
                T.build_status = huff_construct(T.len_count, T.len_symbol, T.lengths, 19, nullptr, 0);
            }
            IC_TEAM_SYNC();
            if (T.build_status != 0) return IC_CORRUPT;
            // the literal/length + distance code lengths: decoded by every lane (the decisions depend on them), stored by lane 0
            int idx = 0, prev = 0;
            while (idx < nlen + ndist) {
                b.refill();
                const int sym = huff_decode(b, nullptr, 0, T.len_count, T.len_symbol);
                if (sym < 0 || b.cnt < 0) return IC_CORRUPT;
                if (sym < 16) {
                    if (lane == 0) T.lengths[idx] = (uint16_t)sym;
                    ++idx;
                    prev = sym;
                } else {
                    int rep, val = 0;
                    if (sym == 16) {
                        if (idx == 0) return IC_CORRUPT;
                        val = prev;
                        rep = 3 + (int)b.bits(2);
                    } else if (sym == 17) rep = 3 + (int)b.bits(3);
                    else rep = 11 + (int)b.bits(7);
                    if (b.cnt < 0 || idx + rep > nlen + ndist) return IC_CORRUPT;
                    if (lane == 0)
                        for (int k = 0; k < rep; ++k) T.lengths[idx + k] = (uint16_t)val;
                    idx += rep;
                    prev = val;
                }
            }
            IC_TEAM_SYNC();                                   // (the code-length tables are done with; lane 0's stores are complete)
            if (lane == 0) {
                int st = IC_OK;
                if (T.lengths[256] == 0) st = IC_CORRUPT;     // no end-of-block code
                if (st == IC_OK) {
                    const int e = huff_construct(T.lit_count, T.lit_symbol, T.lengths, nlen, T.lit_fast, kLitFast);
                    if (e < 0 || (e > 0 && nlen != T.lit_count[0] + T.lit_count[1])) st = IC_CORRUPT;   // incomplete: only a lone 1-bit code
                }
                if (st == IC_OK) {
                    const int e = huff_construct(T.dist_count, T.dist_symbol, T.lengths + nlen, ndist, T.dist_fast, kDistFast);
                    if (e < 0 || (e > 0 && ndist != T.dist_count[0] + T.dist_count[1])) st = IC_CORRUPT;
                }
                T.build_status = st;
            }
            IC_TEAM_SYNC();
            if (T.build_status != IC_OK) return IC_CORRUPT;
            const int r = inflate_codes(b, out, pos, cap, T, full, lane, nlanes);
            if (r != IC_OK) return r;
        } else {
            return IC_CORRUPT;
        }
    } while (!last && !full);
    *out_len = pos;
    return IC_OK;
}

// =====================================================================================================================
// PNG scan-line filters (PNG spec 9.2); cur is reconstructed in place, prev = the reconstructed line above (or null)
// =====================================================================================================================
IC_FN inline int png_unfilter_row(int ft, uint8_t* cur, const uint8_t* prev, int rowbytes, int bpp) {
    switch (ft) {
        case 0: return IC_OK;
        case 1:
            for (int i = bpp; i < rowbytes; ++i) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]);
            return IC_OK;
        case 2:
            if (prev)
                for (int i = 0; i < rowbytes; ++i) cur[i] = (uint8_t)(cur[i] + prev[i]);
            return IC_OK;
        case 3:
            for (int i = 0; i < rowbytes; ++i) {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0;
                cur[i] = (uint8_t)(cur[i] + ((a + b) >> 1));
            }
            return IC_OK;
        case 4:
            for (int i = 0; i < rowbytes; ++i) {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= bpp) ? prev[i - bpp] : 0;
                const int p = a + b - c;
                const int pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
                const int pr = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                cur[i] = (uint8_t)(cur[i] + pr);
            }
            return IC_OK;
        default: return IC_CORRUPT;
    }
}

struct PngFrame {                 // filled by png_parse (host), read by the kernels
    int32_t status;
    int32_t width, height;
    int32_t channels;             // 1 (grey), 3 (RGB), 4 (RGBA)
    int32_t bit_depth;            // 8 or 16
    int32_t rowbytes, bpp;
    int32_t pad_;
    int64_t z_off, z_len;         // the concatenated IDAT payload inside the staging blob
};

// Bands of scan lines that can be reconstructed independently: a band may only START at a line whose filter does not
// look at the line above (None / Sub).  band_first(b) = the first such line at or after b * rows_per_band.
IC_FN inline int png_band_first(const uint8_t* raw, int height, int stride, int row0) {
    if (row0 <= 0) return 0;
    int r = row0;
    while (r < height && raw[(int64_t)r * stride] > 1) ++r;
    return r < height ? r : height;
}
IC_FN inline int png_unfilter_band(uint8_t* raw, int height, int rowbytes, int bpp, int row_begin, int row_end) {
    const int stride = rowbytes + 1;
    int status = IC_OK;
    for (int r = row_begin; r < row_end; ++r) {
        uint8_t* cur = raw + (int64_t)r * stride;
        if (png_unfilter_row(cur[0], cur + 1, r > 0 ? cur + 1 - stride : nullptr, rowbytes, bpp) != IC_OK) status = IC_CORRUPT;
    }
    return status;
}

// one pixel of a reconstructed image -> the array the frame loop integrates: depth u16 (PNG stores big-endian samples),
// colour RGB u8 (alpha dropped, as the loop's cvtColor does)
IC_FN inline void png_emit_pixel(const uint8_t* raw, const PngFrame& f, int x, int y, uint8_t* out) {
    const uint8_t* px = raw + (int64_t)y * (f.rowbytes + 1) + 1 + (int64_t)x * f.bpp;
    if (f.channels == 1) {
        uint8_t* o = out + ((int64_t)y * f.width + x) * 2;
        o[0] = px[1];
        o[1] = px[0];
    } else {
        uint8_t* o = out + ((int64_t)y * f.width + x) * 3;
        o[0] = px[0];
        o[1] = px[1];
        o[2] = px[2];
    }
}

// =====================================================================================================================
// baseline JPEG
// =====================================================================================================================
constexpr int kJLook = 10;

struct JpegHuffTable {
    uint16_t look[1 << kJLook];   // (code length << 8) | symbol for codes of <= kJLook bits, 0 otherwise
    int32_t maxcode[18];          // largest code of each length (-1 if none); [17] = sentinel
    int32_t valoff[17];           // index of the first symbol of a length minus its first code
    uint8_t vals[256];
};
struct JpegTables {               // one per distinct (DHT, DQT) content of a batch; nearly always one per batch
    JpegHuffTable h[4];           // dc0, dc1, ac0, ac1
    uint16_t q[4][64];            // quantisation tables, zig-zag order as stored in the file
    // AC symbols whose code AND value bits fit the lookahead window, resolved by one load: bits 0-4 = bits consumed
    // (code + value), bits 5-10 = positions to skip first, bits 16-31 = the coefficient (int16); 0 = take the two-step
    // path.  End-of-block is "skip 63" (leaves the block), ZRL is "skip 15, store 0" (the block is pre-zeroed).
    uint32_t acfast[2][1 << kJLook];
};
struct JpegFrame {
    int32_t status;
    int32_t width, height;
    int32_t hs[3], vs[3], tq[3], td[3], ta[3];
    int32_t hmax, vmax;
    int32_t mcus_x, mcus_y;
    int32_t restart;              // MCUs per restart interval, 0 = none
    int32_t tables;               // index of the JpegTables set
    int32_t wblk[3], hblk[3];     // component planes in 8x8 blocks (padded to whole MCUs)
    int32_t blk_base[3];          // first block of each component inside the frame's block range
    int32_t n_blocks;             // blocks of the frame
    int64_t blk_off;              // first block of the frame inside the batch's coefficient / sample buffers
    int64_t scan_off, scan_len;   // entropy-coded segment inside the staging blob (ends before the closing marker)
};

IC_DEFINE_TABLE(uint8_t, kZigzag, 64,
                {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63})
// natural (row-major) position -> zig-zag index
IC_DEFINE_TABLE(uint8_t, kUnzigzag, 64,
                {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30,
                 41, 43, 9,  11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38,
                 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63})

struct BitsMSB {                  // JPEG packs bits from the most significant bit; 0xFF data bytes are followed by 0x00
    const uint8_t* p;
    const uint8_t* end;
    uint64_t buf;                 // next bit = bit 63
    int cnt;
    bool starved;                 // ran into a marker or the end: zero bits are fed from here on (as libjpeg does)
    IC_FN void refill() {         // afterwards cnt > 32
        if (cnt > 32) return;
        if (!starved && end - p >= 4) {
            const uint32_t b0 = p[0], b1 = p[1], b2 = p[2], b3 = p[3];
            if (b0 != 0xFF && b1 != 0xFF && b2 != 0xFF && b3 != 0xFF) {
                const uint32_t w = (b0 << 24) | (b1 << 16) | (b2 << 8) | b3;
                buf |= (uint64_t)w << (32 - cnt);
                cnt += 32;
                p += 4;
                return;
            }
        }
        while (cnt <= 56) {
            if (starved || p >= end) { starved = true; cnt = 64; return; }
            const uint32_t c = *p;
            if (c == 0xFF) {
                if (end - p < 2 || p[1] != 0) { starved = true; cnt = 64; return; }   // a marker: not consumed
                p += 2;
            } else {
                p += 1;
            }
            buf |= (uint64_t)c << (56 - cnt);
            cnt += 8;
        }
    }
    IC_FN void drop(int n) {
        buf <<= n;
        cnt -= n;
    }
    IC_FN int get(int n) {        // 1 <= n <= 16
        const int v = (int)(buf >> (64 - n));
        drop(n);
        return v;
    }
};

IC_FN inline int jpeg_huff_decode(BitsMSB& b, const JpegHuffTable& t) {
    const uint32_t e = t.look[(uint32_t)(b.buf >> (64 - kJLook))];
    if (e) {
        b.drop((int)(e >> 8));
        return (int)(e & 255u);
    }
    for (int l = kJLook + 1; l <= 16; ++l) {
        const int code = (int)(b.buf >> (64 - l));
        if (code <= t.maxcode[l]) {
            b.drop(l);
            return t.vals[(code + t.valoff[l]) & 255];
        }
    }
    b.drop(16);
    return -1;                    // libjpeg: "corrupt JPEG data: bad Huffman code", decodes as zero
}

IC_FN inline int jpeg_extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

// Entropy-decode the single interleaved scan of a baseline frame into coef[n_blocks][64] (zig-zag order, int16,
// pre-zeroed by the caller).  Returns IC_OK also for damaged scans (missing bits decode as zeros, like libjpeg's
// warning path); IC_CORRUPT is kept for structure errors the stock decoder treats as fatal.
IC_FN inline int jpeg_decode_scan(const JpegFrame& f, const JpegTables& T, const uint8_t* scan, int16_t* coef) {
    BitsMSB b{scan, scan + f.scan_len, 0, 0, false};
    int pred[3] = {0, 0, 0};
    int to_restart = f.restart;
    int next_rst = 0;
    for (int my = 0; my < f.mcus_y; ++my) {
        for (int mx = 0; mx < f.mcus_x; ++mx) {
            if (f.restart) {
                if (to_restart == 0) {
                    // byte-align, step over RSTn, reset the predictors (T.81 F.2.1.3.1)
                    b.buf = 0;
                    b.cnt = 0;
                    if (b.end - b.p >= 2 && b.p[0] == 0xFF && b.p[1] == (uint8_t)(0xD0 + next_rst)) {
                        b.p += 2;
                        b.starved = false;
                    } else {
                        b.starved = true;                       // resynchronisation is not attempted: zeros from here on
                    }
                    next_rst = (next_rst + 1) & 7;
                    pred[0] = pred[1] = pred[2] = 0;
                    to_restart = f.restart;
                }
                --to_restart;
            }
            for (int c = 0; c < 3; ++c) {
                const JpegHuffTable& dc = T.h[f.td[c]];
                const JpegHuffTable& ac = T.h[2 + f.ta[c]];
                for (int v = 0; v < f.vs[c]; ++v) {
                    for (int h = 0; h < f.hs[c]; ++h) {
                        int16_t* blk = coef + ((int64_t)f.blk_base[c] + (int64_t)(my * f.vs[c] + v) * f.wblk[c] + (mx * f.hs[c] + h)) * 64;
                        b.refill();
                        int s = jpeg_huff_decode(b, dc);
                        if (s < 0) s = 0;
                        if (s) {
                            const int r = b.get(s);             // s <= 15 (table check), <= 16 bits left after the code
                            pred[c] += jpeg_extend(r, s);
                        }
                        blk[0] = (int16_t)pred[c];
                        const uint32_t* acfast = T.acfast[f.ta[c]];
                        for (int k = 1; k < 64;) {
                            b.refill();
                            const uint32_t e = acfast[(uint32_t)(b.buf >> (64 - kJLook))];
                            if (e) {                                // run, size and value from one lookup
                                k += (int)((e >> 5) & 63u);
                                if (k < 64) blk[k] = (int16_t)(e >> 16);
                                ++k;
                                b.drop((int)(e & 31u));
                                continue;
                            }
                            int rs = jpeg_huff_decode(b, ac);
                            if (rs < 0) rs = 0;
                            const int r = rs >> 4;
                            s = rs & 15;
                            if (s) {
                                k += r;
                                const int v2 = jpeg_extend(b.get(s), s);
                                if (k < 64) blk[k] = (int16_t)v2;
                                ++k;
                            } else {
                                if (r != 15) break;
                                k += 16;
                            }
                        }
                    }
                }
            }
        }
    }
    return IC_OK;
}

// ---- inverse DCT: libjpeg's jidctint.c "islow" (Loeffler-Ligtenberg-Moschytz, 13-bit constants, 2 extra bits between passes)
constexpr int kConstBits = 13, kPass1Bits = 2;
IC_FN inline void idct_1d(const int32_t in[8], int32_t out[8], int shift) {
    // 32-bit two's-complement arithmetic throughout (what the SIMD builds of libjpeg-turbo do); unsigned so that the wrap of
    // absurd coefficients in a damaged file is defined behaviour
    typedef uint32_t U;
    // even part
    U z2 = (U)in[2], z3 = (U)in[6];
    U z1 = (z2 + z3) * (U)4433;                          // FIX(0.541196100)
    U tmp2 = z1 + z3 * (U)(-15137);                      // FIX(1.847759065)
    U tmp3 = z1 + z2 * (U)6270;                          // FIX(0.765366865)
    z2 = (U)in[0];
    z3 = (U)in[4];
    U tmp0 = (z2 + z3) << kConstBits;
    U tmp1 = (z2 - z3) << kConstBits;
    const U tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    // odd part
    tmp0 = (U)in[7];
    tmp1 = (U)in[5];
    tmp2 = (U)in[3];
    tmp3 = (U)in[1];
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    U z4 = tmp1 + tmp3;
    const U z5 = (z3 + z4) * (U)9633;                    // FIX(1.175875602)
    tmp0 *= (U)2446;                                     // FIX(0.298631336)
    tmp1 *= (U)16819;                                    // FIX(2.053119869)
    tmp2 *= (U)25172;                                    // FIX(3.072711026)
    tmp3 *= (U)12299;                                    // FIX(1.501321110)
    z1 *= (U)(-7373);                                    // FIX(0.899976223)
    z2 *= (U)(-20995);                                   // FIX(2.562915447)
    z3 *= (U)(-16069);                                   // FIX(1.961570560)
    z4 *= (U)(-3196);                                    // FIX(0.390180644)
    z3 += z5;
    z4 += z5;
    tmp0 += z1 + z3;
    tmp1 += z2 + z4;
    tmp2 += z2 + z3;
    tmp3 += z1 + z4;
    const U rnd = (U)1 << (shift - 1);
    out[0] = (int32_t)(tmp10 + tmp3 + rnd) >> shift;
    out[7] = (int32_t)(tmp10 - tmp3 + rnd) >> shift;
    out[1] = (int32_t)(tmp11 + tmp2 + rnd) >> shift;
    out[6] = (int32_t)(tmp11 - tmp2 + rnd) >> shift;
    out[2] = (int32_t)(tmp12 + tmp1 + rnd) >> shift;
    out[5] = (int32_t)(tmp12 - tmp1 + rnd) >> shift;
    out[3] = (int32_t)(tmp13 + tmp0 + rnd) >> shift;
    out[4] = (int32_t)(tmp13 - tmp0 + rnd) >> shift;
}
// column c of a block: coefficients (zig-zag order) x quantisation steps -> 8 workspace values
IC_FN inline void idct_column(const int16_t* blk_zz, const uint16_t* q_zz, int c, int32_t ws[8]) {
    int32_t in[8];
    for (int r = 0; r < 8; ++r) {
        const int zz = IC_TABLE(kUnzigzag)[r * 8 + c];
        in[r] = (int32_t)blk_zz[zz] * (int32_t)q_zz[zz];
    }
    idct_1d(in, ws, kConstBits - kPass1Bits);
}
// libjpeg's range-limit table, indexed with (x & 1023) after the +128 level shift
IC_FN inline uint8_t idct_range_limit(int32_t x) {
    const int v = x & 1023;
    if (v < 512) return (uint8_t)(v + 128 > 255 ? 255 : v + 128);
    const int w = v - 1024 + 128;
    return (uint8_t)(w < 0 ? 0 : w);
}
IC_FN inline void idct_row(const int32_t ws[8], uint8_t out[8]) {
    int32_t o[8];
    idct_1d(ws, o, kConstBits + kPass1Bits + 3);
    for (int i = 0; i < 8; ++i) out[i] = idct_range_limit(o[i]);
}

// ---- chroma upsampling (jdsample.c, do_fancy_upsampling) + YCbCr -> RGB (jdcolor.c) for one output pixel
struct JpegPlanes {
    const uint8_t* y;
    const uint8_t* cb;
    const uint8_t* cr;
    int ys, cs;                   // row strides of the luma / chroma planes (bytes)
    int cw, ch;                   // chroma size in samples that exist in the image (downsampled_width / _height)
    int hmax, vmax;               // 1 or 2
};
IC_FN inline int jpeg_chroma_at(const uint8_t* pl, int stride, int cw, int ch, int hmax, int vmax, int x, int y) {
    if (hmax == 1) return pl[(int64_t)y * stride + x];                   // vmax == 1 as well (4:4:4)
    const int c = x >> 1;
    if (vmax == 1) {                                                      // h2v1_fancy_upsample
        const uint8_t* row = pl + (int64_t)y * stride;
        const int v = row[c];
        if (x & 1) return c == cw - 1 ? v : (v * 3 + row[c + 1] + 2) >> 2;
        return c == 0 ? v : (v * 3 + row[c - 1] + 1) >> 2;
    }
    // h2v2_fancy_upsample: 3/4 nearer + 1/4 further line, then the same across; lines beyond the edge repeat the edge
    const int i = y >> 1;
    int j = (y & 1) ? i + 1 : i - 1;
    j = j < 0 ? 0 : (j > ch - 1 ? ch - 1 : j);
    const uint8_t* r0 = pl + (int64_t)i * stride;
    const uint8_t* r1 = pl + (int64_t)j * stride;
    const int cur = r0[c] * 3 + r1[c];
    if (x & 1) {
        if (c == cw - 1) return (cur * 4 + 7) >> 4;
        return (cur * 3 + (r0[c + 1] * 3 + r1[c + 1]) + 7) >> 4;
    }
    if (c == 0) return (cur * 4 + 8) >> 4;
    return (cur * 3 + (r0[c - 1] * 3 + r1[c - 1]) + 8) >> 4;
}
IC_FN inline uint8_t clamp_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
IC_FN inline void jpeg_pixel_rgb(const JpegPlanes& P, int x, int y, uint8_t rgb[3]) {
    const int yy = P.y[(int64_t)y * P.ys + x];
    const int cb = jpeg_chroma_at(P.cb, P.cs, P.cw, P.ch, P.hmax, P.vmax, x, y) - 128;
    const int cr = jpeg_chroma_at(P.cr, P.cs, P.cw, P.ch, P.hmax, P.vmax, x, y) - 128;
    rgb[0] = clamp_u8(yy + ((91881 * cr + 32768) >> 16));                               // FIX(1.40200)
    rgb[1] = clamp_u8(yy + ((-22554 * cb + 32768 - 46802 * cr) >> 16));                 // FIX(0.34414), FIX(0.71414)
    rgb[2] = clamp_u8(yy + ((116130 * cb + 32768) >> 16));                              // FIX(1.77200)
}

// =====================================================================================================================
// container parsing (host only)
// =====================================================================================================================
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline uint32_t be16(const uint8_t* p) { return ((uint32_t)p[0] << 8) | p[1]; }

struct Crc32 {                    // slice-by-8 (PNG spec annex D polynomial)
    uint32_t t[8][256];
    Crc32() {
        for (uint32_t n = 0; n < 256; ++n) {
            uint32_t c = n;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            t[0][n] = c;
        }
        for (uint32_t n = 0; n < 256; ++n)
            for (int s = 1; s < 8; ++s) t[s][n] = (t[s - 1][n] >> 8) ^ t[0][t[s - 1][n] & 255];
    }
    uint32_t run(const uint8_t* p, size_t n, uint32_t crc = 0) const {
        crc = ~crc;
        while (n >= 8) {
            const uint32_t a = ((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24)) ^ crc;
            crc = t[7][a & 255] ^ t[6][(a >> 8) & 255] ^ t[5][(a >> 16) & 255] ^ t[4][a >> 24] ^ t[3][p[4]] ^ t[2][p[5]] ^ t[1][p[6]] ^
                  t[0][p[7]];
            p += 8;
            n -= 8;
        }
        while (n--) crc = t[0][(crc ^ *p++) & 255] ^ (crc >> 8);
        return ~crc;
    }
};

// Walk the chunks of one PNG file, check what libpng checks before it hands out pixels (signature, IHDR, chunk CRCs
// of the critical chunks when `crc` is given), and append the IDAT payloads to dst (capacity dst_cap).  Fills f
// (z_off is left to the caller); returns the status also stored in f.status.
inline int png_parse(const uint8_t* file, int64_t size, int want_h, int want_w, uint8_t* dst, int64_t dst_cap, const Crc32* crc,
                     PngFrame& f) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    f = PngFrame{};
    f.status = IC_CORRUPT;
    if (size < 8 + 25) return f.status;
    for (int i = 0; i < 8; ++i)
        if (file[i] != sig[i]) return f.status;
    int64_t p = 8, zl = 0;
    bool have_hdr = false, seen_idat = false;
    while (p + 12 <= size) {
        const uint32_t len = be32(file + p);
        const uint8_t* type = file + p + 4;
        if ((int64_t)len > size - p - 12) break;
        const bool critical = !(type[0] & 0x20);
        if (crc && critical && crc->run(type, (size_t)len + 4) != be32(file + p + 8 + len)) return f.status;
        const uint8_t* d = file + p + 8;
        if (!have_hdr) {
            if (!(type[0] == 'I' && type[1] == 'H' && type[2] == 'D' && type[3] == 'R') || len != 13) return f.status;
            const uint32_t w = be32(d), h = be32(d + 4);
            const int depth = d[8], ctype = d[9];
            if (d[10] != 0 || d[11] != 0 || d[12] > 1 || w == 0 || h == 0 || w > 0x7FFFFFFFu || h > 0x7FFFFFFFu) return f.status;
            have_hdr = true;
            f.width = (int32_t)w;
            f.height = (int32_t)h;
            f.bit_depth = depth;
            f.channels = ctype == 0 ? 1 : (ctype == 2 ? 3 : (ctype == 6 ? 4 : 0));
            const bool ok = d[12] == 0 && ((ctype == 0 && depth == 16) || ((ctype == 2 || ctype == 6) && depth == 8));
            if (!ok || (int)h != want_h || (int)w != want_w) {           // a valid PNG of another kind: the stock decoder's business
                f.status = IC_UNSUPPORTED;
                return f.status;
            }
            f.bpp = f.channels * depth / 8;
            f.rowbytes = f.bpp * f.width;
        } else if (type[0] == 'I' && type[1] == 'D' && type[2] == 'A' && type[3] == 'T') {
            if (zl + (int64_t)len > dst_cap) { f.status = IC_UNSUPPORTED; return f.status; }   // larger than the staging slot
            memcpy(dst + zl, d, len);
            zl += len;
            seen_idat = true;
        } else if (type[0] == 'I' && type[1] == 'E' && type[2] == 'N' && type[3] == 'D') {
            break;
        } else if (type[0] == 'P' && type[1] == 'L' && type[2] == 'T' && type[3] == 'E') {
            // legal in RGB files (a suggested palette); ignored
        } else if (type[0] == 't' && type[1] == 'R' && type[2] == 'N' && type[3] == 'S') {
            f.status = IC_UNSUPPORTED;                                    // transparency key: OpenCV would synthesise alpha
            return f.status;
        } else if (critical) {
            return f.status;                                              // unknown critical chunk
        }
        p += 12 + (int64_t)len;
    }
    // (a missing IEND is not an error: libpng hands out the image once the IDATs are complete)
    if (!have_hdr || !seen_idat) return f.status;
    f.z_len = zl;
    f.status = IC_OK;
    return f.status;
}

// Build one Huffman decoding table from a DHT segment's counts / symbols (T.81 annex C + F.2.2.3 lookahead).
inline bool jpeg_build_huff(const uint8_t bits[16], const uint8_t* vals, int nvals, bool is_dc, JpegHuffTable& t) {
    t = JpegHuffTable{};
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        const int n = bits[l - 1];
        t.valoff[l] = k - code;
        if (n) {
            for (int i = 0; i < n; ++i, ++k, ++code) {
                if (k >= nvals) return false;
                if (is_dc && vals[k] > 15) return false;
                t.vals[k] = vals[k];
                if (l <= kJLook) {
                    const int first = code << (kJLook - l);
                    for (int j = 0; j < (1 << (kJLook - l)); ++j) t.look[first + j] = (uint16_t)((l << 8) | vals[k]);
                }
            }
            if (code > (1 << l)) return false;                            // more codes than the length can hold
            t.maxcode[l] = code - 1;
        } else {
            t.maxcode[l] = -1;
        }
        code <<= 1;
    }
    t.maxcode[17] = 0xFFFFF;
    return k == nvals;
}

// the one-lookup table of an AC Huffman table (JpegTables::acfast)
inline void jpeg_build_acfast(const JpegHuffTable& ac, uint32_t* out) {
    for (uint32_t peek = 0; peek < (1u << kJLook); ++peek) {
        out[peek] = 0;
        const uint32_t e = ac.look[peek];
        if (!e) continue;
        const int len = (int)(e >> 8), r = (int)((e & 255u) >> 4), sz = (int)(e & 15u);
        if (sz == 0) {
            out[peek] = (uint32_t)len | ((r == 15 ? 15u : 63u) << 5);
            continue;
        }
        if (len + sz > kJLook) continue;
        const int extra = (int)((peek >> (kJLook - len - sz)) & ((1u << sz) - 1u));
        const int val = jpeg_extend(extra, sz);
        out[peek] = (uint32_t)(len + sz) | ((uint32_t)r << 5) | ((uint32_t)(uint16_t)(int16_t)val << 16);
    }
}

struct JpegHeader {               // raw table bytes of one file: the batch decoder de-duplicates them into JpegTables sets
    uint8_t q[4][64];
    bool q_set[4];
    uint8_t hbits[4][16];
    uint8_t hvals[4][256];
    int hn[4];
    bool h_set[4];
};

// Walk the markers of one JPEG file up to the start of its scan.  Supported (status IC_OK): baseline sequential (SOF0),
// 8-bit, three components Y / Cb / Cr with luma sampling 1x1, 2x1 or 2x2 and chroma 1x1, one interleaved scan, 8-bit
// quantisation tables, JFIF colour convention (no Adobe marker).  Anything else that is still a JPEG is IC_UNSUPPORTED
// (handed to the stock decoder); broken structure is IC_CORRUPT.
inline int jpeg_parse(const uint8_t* file, int64_t size, int want_h, int want_w, JpegFrame& f, JpegHeader& hd) {
    f = JpegFrame{};
    hd = JpegHeader{};
    f.status = IC_CORRUPT;
    if (size < 4 || file[0] != 0xFF || file[1] != 0xD8) return f.status;
    int64_t p = 2;
    bool have_sof = false, saw_jfif = false;
    int comp_id[3] = {0, 0, 0};
    for (;;) {
        if (p + 4 > size) return f.status;
        if (file[p] != 0xFF) return f.status;
        while (p < size && file[p] == 0xFF) ++p;                          // fill bytes
        if (p >= size) return f.status;
        const int m = file[p++];
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) return f.status;                                   // EOI before any scan
        if (p + 2 > size) return f.status;
        const int64_t len = be16(file + p);
        if (len < 2 || p + len > size) return f.status;
        const uint8_t* d = file + p + 2;
        const int64_t n = len - 2;
        if (m == 0xDB) {                                                  // DQT
            int64_t i = 0;
            while (i < n) {
                const int pq = d[i] >> 4, tq = d[i] & 15;
                if (tq > 3) return f.status;
                if (pq != 0) { f.status = IC_UNSUPPORTED; return f.status; }
                if (i + 65 > n) return f.status;
                for (int k = 0; k < 64; ++k) hd.q[tq][k] = d[i + 1 + k];
                hd.q_set[tq] = true;
                i += 65;
            }
        } else if (m == 0xC4) {                                           // DHT
            int64_t i = 0;
            while (i < n) {
                if (i + 17 > n) return f.status;
                const int tc = d[i] >> 4, th = d[i] & 15;
                if (tc > 1 || th > 3) return f.status;
                if (th > 1) { f.status = IC_UNSUPPORTED; return f.status; }   // baseline allows two tables per class
                int cnt = 0;
                for (int k = 0; k < 16; ++k) cnt += d[i + 1 + k];
                if (cnt > 256 || i + 17 + cnt > n) return f.status;
                const int slot = tc * 2 + th;
                for (int k = 0; k < 16; ++k) hd.hbits[slot][k] = d[i + 1 + k];
                for (int k = 0; k < cnt; ++k) hd.hvals[slot][k] = d[i + 17 + k];
                hd.hn[slot] = cnt;
                hd.h_set[slot] = true;
                i += 17 + cnt;
            }
        } else if (m == 0xC0) {                                           // SOF0
            if (have_sof || n < 6) return f.status;
            const int prec = d[0], h = (int)be16(d + 1), w = (int)be16(d + 3), nc = d[5];
            if (prec != 8 || nc != 3 || h == 0 || w == 0) { f.status = (h == 0 || w == 0) ? IC_CORRUPT : IC_UNSUPPORTED; return f.status; }
            if (n < 6 + 3 * nc) return f.status;
            for (int c = 0; c < 3; ++c) {
                comp_id[c] = d[6 + 3 * c];
                f.hs[c] = d[7 + 3 * c] >> 4;
                f.vs[c] = d[7 + 3 * c] & 15;
                f.tq[c] = d[8 + 3 * c];
                if (f.tq[c] > 3) return f.status;
            }
            const bool samp_ok = f.hs[1] == 1 && f.vs[1] == 1 && f.hs[2] == 1 && f.vs[2] == 1 &&
                                 ((f.hs[0] == 1 && f.vs[0] == 1) || (f.hs[0] == 2 && f.vs[0] == 1) || (f.hs[0] == 2 && f.vs[0] == 2));
            if (!samp_ok || h != want_h || w != want_w) { f.status = IC_UNSUPPORTED; return f.status; }
            f.width = w;
            f.height = h;
            f.hmax = f.hs[0];
            f.vmax = f.vs[0];
            f.mcus_x = (w + 8 * f.hmax - 1) / (8 * f.hmax);
            f.mcus_y = (h + 8 * f.vmax - 1) / (8 * f.vmax);
            if (f.hmax == 2 && (w + 1) / 2 <= 2) { f.status = IC_UNSUPPORTED; return f.status; }   // libjpeg switches filters there
            int base = 0;
            for (int c = 0; c < 3; ++c) {
                f.wblk[c] = f.mcus_x * f.hs[c];
                f.hblk[c] = f.mcus_y * f.vs[c];
                f.blk_base[c] = base;
                base += f.wblk[c] * f.hblk[c];
            }
            f.n_blocks = base;
            have_sof = true;
        } else if (m == 0xC1 || m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
            f.status = IC_UNSUPPORTED;                                    // extended / progressive / lossless / arithmetic
            return f.status;
        } else if (m == 0xDD) {                                           // DRI
            if (n != 2) return f.status;
            f.restart = (int)be16(d);
        } else if (m == 0xE0) {                                           // APP0
            if (n >= 5 && d[0] == 'J' && d[1] == 'F' && d[2] == 'I' && d[3] == 'F' && d[4] == 0) saw_jfif = true;
        } else if (m == 0xEE) {                                           // Adobe: colour transform flag the JFIF rule does not cover
            f.status = IC_UNSUPPORTED;
            return f.status;
        } else if (m == 0xDA) {                                           // SOS
            if (!have_sof) return f.status;
            if (!saw_jfif && comp_id[0] == 'R' && comp_id[1] == 'G' && comp_id[2] == 'B') {   // libjpeg then takes the data for RGB
                f.status = IC_UNSUPPORTED;
                return f.status;
            }
            if (n < 1 || d[0] != 3 || n != 1 + 2 * 3 + 3) { f.status = (n >= 1 && d[0] >= 1 && d[0] <= 4) ? IC_UNSUPPORTED : IC_CORRUPT; return f.status; }
            for (int c = 0; c < 3; ++c) {
                if (d[1 + 2 * c] != comp_id[c]) { f.status = IC_UNSUPPORTED; return f.status; }
                f.td[c] = d[2 + 2 * c] >> 4;
                f.ta[c] = d[2 + 2 * c] & 15;
                if (f.td[c] > 1 || f.ta[c] > 1) { f.status = IC_UNSUPPORTED; return f.status; }
                if (!hd.h_set[f.td[c]] || !hd.h_set[2 + f.ta[c]] || !hd.q_set[f.tq[c]]) return f.status;
            }
            if (d[7] != 0 || d[8] != 63 || d[9] != 0) { f.status = IC_UNSUPPORTED; return f.status; }
            const int64_t s0 = p + len;
            int64_t e = s0;                                               // the scan ends at the first marker that is not RSTn
            while (e < size) {
                if (file[e] != 0xFF) { ++e; continue; }
                if (e + 1 >= size) { e = size; break; }
                const int nx = file[e + 1];
                if (nx == 0x00 || (nx >= 0xD0 && nx <= 0xD7)) { e += 2; continue; }
                if (nx == 0xFF) { ++e; continue; }
                break;
            }
            if (e > size) e = size;
            f.scan_off = s0;
            f.scan_len = e - s0;
            f.status = IC_OK;
            return f.status;
        }
        // APPn, COM and everything else: skipped
        p += len;
    }
}
}  // namespace imgcodec
