// ingest.cu — raw-scan ingest: parallel host-side decode of the Oxford radar scan container
// (8-bit grayscale, non-interlaced PNG, 400 x 3779) straight into a caller buffer — typically the pinned staging
// buffer rf_batch_upload_async reads.  Host code only (no kernels); zlib does the inflate.
//
// Replaces (reference file:line):
//   parseData.py:160-179   getDataFromImgPathsByIndex: cv2.imread(path, cv2.IMREAD_GRAYSCALE), one file at a time
//   utils.py:22-26         radarImgPathToTimestamp (the file stem is the timestamp)
// At the front end's frame rates the per-file cv2.imread (~20 ms) is the wall; files are independent, so they are
// decoded by a pool of threads, each inflating directly into its slice of the output.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

struct PngInfo { int width, height, bit_depth, color_type, interlace; };

// reads the whole file; returns "" on success
std::string read_file(const char* path, std::vector<uint8_t>& buf) {
    FILE* f = fopen(path, "rb");
    if (!f) return std::string("cannot open ") + path;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 0) { fclose(f); return std::string("cannot stat ") + path; }
    buf.resize((size_t)sz);
    size_t got = sz ? fread(buf.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    if (got != (size_t)sz) return std::string("short read on ") + path;
    return "";
}

std::string parse_header(const std::vector<uint8_t>& b, const char* path, PngInfo* info) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (b.size() < 33 || memcmp(b.data(), sig, 8) != 0) return std::string(path) + ": not a PNG file";
    if (be32(&b[8]) != 13 || memcmp(&b[12], "IHDR", 4) != 0) return std::string(path) + ": missing IHDR";
    info->width = (int)be32(&b[16]); info->height = (int)be32(&b[20]);
    info->bit_depth = b[24]; info->color_type = b[25]; info->interlace = b[28];
    return "";
}

inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// one file -> out[rows][cols]
std::string decode_one(const char* path, int rows, int cols, uint8_t* out) {
    std::vector<uint8_t> b;
    std::string e = read_file(path, b);
    if (!e.empty()) return e;
    PngInfo pi;
    e = parse_header(b, path, &pi);
    if (!e.empty()) return e;
    if (pi.bit_depth != 8 || pi.color_type != 0 || pi.interlace != 0)
        return std::string(path) + ": only 8-bit grayscale non-interlaced PNG (the radar scan container) is supported";
    if (pi.width != cols || pi.height != rows) {
        char m[256];
        snprintf(m, sizeof(m), "%s: image is %d x %d, expected %d x %d", path, pi.height, pi.width, rows, cols);
        return m;
    }
    // inflate the concatenated IDAT payloads into [rows][1 + cols] filtered scanlines
    std::vector<uint8_t> raw((size_t)rows * (cols + 1));
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit(&zs) != Z_OK) return "inflateInit failed";
    zs.next_out = raw.data();
    zs.avail_out = (uInt)raw.size();
    size_t pos = 8;
    bool done = false, seen_idat = false;
    int zrc = Z_OK;
    while (pos + 12 <= b.size()) {
        const uint32_t len = be32(&b[pos]);
        const uint8_t* type = &b[pos + 4];
        if (pos + 12 + (size_t)len > b.size()) { inflateEnd(&zs); return std::string(path) + ": truncated chunk"; }
        const uint32_t crc = be32(&b[pos + 8 + len]);
        if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), type, len + 4) != crc) { inflateEnd(&zs); return std::string(path) + ": chunk CRC mismatch"; }
        if (memcmp(type, "IDAT", 4) == 0 && !done) {   // (IDAT chunks after the end of the deflate stream carry nothing)
            seen_idat = true;
            zs.next_in = const_cast<uint8_t*>(&b[pos + 8]);
            zs.avail_in = len;
            zrc = inflate(&zs, Z_NO_FLUSH);
            if (zrc == Z_STREAM_END) done = true;
            else if (zrc != Z_OK && zrc != Z_BUF_ERROR) { inflateEnd(&zs); return std::string(path) + ": corrupt deflate stream"; }
        } else if (memcmp(type, "IEND", 4) == 0) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    const size_t produced = raw.size() - zs.avail_out;
    inflateEnd(&zs);
    if (!seen_idat || produced != raw.size() || (!done && zs.avail_out != 0)) return std::string(path) + ": image data incomplete";
    // undo the per-scanline filters (bytes per pixel = 1)
    const uint8_t* prev = nullptr;
    for (int r = 0; r < rows; ++r) {
        const uint8_t* src = &raw[(size_t)r * (cols + 1)];
        uint8_t* dst = out + (size_t)r * cols;
        const int ft = src[0];
        ++src;
        switch (ft) {
            case 0: memcpy(dst, src, (size_t)cols); break;
            case 1: { int a = 0; for (int x = 0; x < cols; ++x) { a = (src[x] + a) & 255; dst[x] = (uint8_t)a; } break; }
            case 2: for (int x = 0; x < cols; ++x) dst[x] = (uint8_t)(src[x] + (prev ? prev[x] : 0)); break;
            case 3: { int a = 0; for (int x = 0; x < cols; ++x) { a = (src[x] + ((a + (prev ? prev[x] : 0)) >> 1)) & 255; dst[x] = (uint8_t)a; } break; }
            case 4: {
                int a = 0, c = 0;
                for (int x = 0; x < cols; ++x) {
                    const int bb = prev ? prev[x] : 0;
                    a = (src[x] + paeth(a, bb, c)) & 255;
                    dst[x] = (uint8_t)a;
                    c = bb;
                }
                break;
            }
            default: return std::string(path) + ": unknown scanline filter";
        }
        prev = dst;
    }
    return "";
}

}  // namespace

extern "C" {

// size of one scan file (rows = azimuths, cols = 11 metadata bytes + range bins)
int rf_png_info(const char* path, int* rows, int* cols) {
    if (!path || !rows || !cols) return rf_fail(nullptr, RF_E_BADARG, "rf_png_info: null argument");
    std::vector<uint8_t> b;
    std::string e = read_file(path, b);
    PngInfo pi;
    if (e.empty()) e = parse_header(b, path, &pi);
    if (!e.empty()) return rf_fail(nullptr, RF_E_BADARG, "rf_png_info: %s", e.c_str());
    *rows = pi.height; *cols = pi.width;
    return RF_OK;
}

// n files -> out [n][rows][cols] u8, decoded by `threads` host threads (<= 0: hardware concurrency)
int rf_ingest_png(const char* const* paths, int n, int rows, int cols, int threads, uint8_t* out) {
    if (!paths || !out || n < 0 || rows < 1 || cols < 1) return rf_fail(nullptr, RF_E_BADARG, "rf_ingest_png: bad argument");
    if (n == 0) return RF_OK;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = nt < 1 ? 1 : (nt > n ? n : nt);
    std::atomic<int> next(0), failed(-1);
    std::vector<std::string> errs((size_t)n);
    auto work = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n) return;
            errs[i] = paths[i] ? decode_one(paths[i], rows, cols, out + (size_t)i * rows * cols) : std::string("null path");
            if (!errs[i].empty()) { int exp = -1; failed.compare_exchange_strong(exp, i); }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    for (int i = 0; i < n; ++i)
        if (!errs[i].empty()) return rf_fail(nullptr, RF_E_BADARG, "rf_ingest_png: %s", errs[i].c_str());
    return RF_OK;
}

}  // extern "C"
