// image_io.hpp — film writers, host-side mirror of `SaveImage` (reference include/nanogi/basic.hpp:506-672).
// The film is float RGB, row-major, ROW 0 = BOTTOM scanline (rt.hpp:135-140). Formats by extension:
//   .hdr  Radiance RGBE; FreeImage scanline y = bottom-up, so the file shows row H-1 first (basic.hpp:531-565)
//   .exr  3 x float32 channels "B","G","R", ZIP (16-line blocks), increasing-Y, y flipped (basic.hpp:566-621)
//   .png  8-bit RGB, gamma 1/2.2, clamp(int(pow(v,1/2.2)*255),0,255) (basic.hpp:622-659)
// Only zlib is used (deflate + crc32).
#pragma once
#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <sys/stat.h>

#include "logger.hpp"

namespace ngi {

namespace detail {
inline bool make_parent_dirs(const std::string& path) {  // basic.hpp:510-521
    const size_t s = path.find_last_of('/');
    if (s == std::string::npos) return true;
    const std::string parent = path.substr(0, s);
    if (parent.empty()) return true;
    std::string cur;
    for (size_t i = 0; i <= parent.size(); i++) {
        if (i == parent.size() || parent[i] == '/') {
            if (!cur.empty()) {
                struct stat st;
                if (stat(cur.c_str(), &st) != 0) {
                    NGI_LOG_INFO("Creating directory : " + cur);
                    if (mkdir(cur.c_str(), 0777) != 0) return false;
                }
            }
        }
        if (i < parent.size()) cur += parent[i];
    }
    return true;
}
inline std::string extension(const std::string& path) {
    const size_t d = path.find_last_of('.');
    const size_t s = path.find_last_of('/');
    if (d == std::string::npos || (s != std::string::npos && d < s)) return "";
    return path.substr(d);
}
inline void put32(std::vector<uint8_t>& b, uint32_t v) { for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i))); }
inline void put64(std::vector<uint8_t>& b, uint64_t v) { for (int i = 0; i < 8; i++) b.push_back((uint8_t)(v >> (8 * i))); }
inline void putstr(std::vector<uint8_t>& b, const char* s) { while (*s) b.push_back((uint8_t)*s++); b.push_back(0); }
inline void putf(std::vector<uint8_t>& b, float f) { uint32_t u; std::memcpy(&u, &f, 4); put32(b, u); }
inline void put32be(std::vector<uint8_t>& b, uint32_t v) { for (int i = 3; i >= 0; i--) b.push_back((uint8_t)(v >> (8 * i))); }
}  // namespace detail

// ---- Radiance .hdr ---------------------------------------------------------------------------
inline bool SaveHDR(const std::string& path, const float* film, int width, int height) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", height, width);
    std::vector<uint8_t> row((size_t)width * 4);
    for (int y = height - 1; y >= 0; y--) {  // file is top-down; film row 0 is the bottom
        for (int x = 0; x < width; x++) {
            const float* p = &film[((size_t)y * width + x) * 3];
            float r = p[0] > 0 ? p[0] : 0, g = p[1] > 0 ? p[1] : 0, b = p[2] > 0 ? p[2] : 0;
            float v = r > g ? r : g; if (b > v) v = b;
            uint8_t* o = &row[(size_t)x * 4];
            if (!(v > 1e-32f) || !std::isfinite(v)) { o[0] = o[1] = o[2] = o[3] = 0; }
            else {
                int e; const float m = std::frexp(v, &e) * 256.0f / v;
                o[0] = (uint8_t)(r * m); o[1] = (uint8_t)(g * m); o[2] = (uint8_t)(b * m); o[3] = (uint8_t)(e + 128);
            }
        }
        std::fwrite(row.data(), 1, row.size(), f);
    }
    std::fclose(f);
    return true;
}

// ---- OpenEXR scanline, ZIP, FLOAT B/G/R ------------------------------------------------------
inline bool SaveEXR(const std::string& path, const float* film, int width, int height) {
    using namespace detail;
    std::vector<uint8_t> out;
    put32(out, 20000630u);  // magic 0x76 0x2f 0x31 0x01
    put32(out, 2u);         // version 2, single-part scanline
    // channels (alphabetical): B, G, R — FLOAT (=2), pLinear 0, sampling 1,1
    putstr(out, "channels"); putstr(out, "chlist");
    put32(out, 3 * (2 + 4 + 4 + 4 + 4) + 1);
    for (const char* c : {"B", "G", "R"}) { putstr(out, c); put32(out, 2); out.push_back(0); out.push_back(0); out.push_back(0); out.push_back(0); put32(out, 1); put32(out, 1); }
    out.push_back(0);
    putstr(out, "compression"); putstr(out, "compression"); put32(out, 1); out.push_back(3);  // ZIP_COMPRESSION
    putstr(out, "dataWindow"); putstr(out, "box2i"); put32(out, 16); put32(out, 0); put32(out, 0); put32(out, (uint32_t)(width - 1)); put32(out, (uint32_t)(height - 1));
    putstr(out, "displayWindow"); putstr(out, "box2i"); put32(out, 16); put32(out, 0); put32(out, 0); put32(out, (uint32_t)(width - 1)); put32(out, (uint32_t)(height - 1));
    putstr(out, "lineOrder"); putstr(out, "lineOrder"); put32(out, 1); out.push_back(0);       // INCREASING_Y
    putstr(out, "pixelAspectRatio"); putstr(out, "float"); put32(out, 4); putf(out, 1.0f);
    putstr(out, "screenWindowCenter"); putstr(out, "v2f"); put32(out, 8); putf(out, 0.0f); putf(out, 0.0f);
    putstr(out, "screenWindowWidth"); putstr(out, "float"); put32(out, 4); putf(out, 1.0f);
    out.push_back(0);  // end of header

    const int linesPerBlock = 16;
    const int nBlocks = (height + linesPerBlock - 1) / linesPerBlock;
    const size_t tableAt = out.size();
    out.resize(out.size() + (size_t)nBlocks * 8, 0);
    std::vector<uint8_t> raw, tmp, comp;
    for (int blk = 0; blk < nBlocks; blk++) {
        const int y0 = blk * linesPerBlock, y1 = std::min(height, y0 + linesPerBlock);
        raw.clear();
        for (int y = y0; y < y1; y++) {
            const int fy = height - 1 - y;  // y flip, basic.hpp:583-589
            for (int c = 0; c < 3; c++) {   // B, G, R
                const int src = 2 - c;
                for (int x = 0; x < width; x++) {
                    const float v = film[((size_t)fy * width + x) * 3 + src];
                    uint32_t u; std::memcpy(&u, &v, 4);
                    for (int i = 0; i < 4; i++) raw.push_back((uint8_t)(u >> (8 * i)));
                }
            }
        }
        // ZIP preprocessing: de-interleave even/odd bytes, then delta predictor
        const size_t n = raw.size();
        tmp.resize(n);
        {
            size_t t1 = 0, t2 = (n + 1) / 2;
            for (size_t i = 0; i < n; i++) { if ((i & 1) == 0) tmp[t1++] = raw[i]; else tmp[t2++] = raw[i]; }
            int p = tmp[0];
            for (size_t i = 1; i < n; i++) { const int d = (int)tmp[i] - p + (128 + 256); p = tmp[i]; tmp[i] = (uint8_t)d; }
        }
        uLongf clen = compressBound((uLong)n);
        comp.resize(clen);
        if (compress2(comp.data(), &clen, tmp.data(), (uLong)n, Z_DEFAULT_COMPRESSION) != Z_OK) return false;
        const uint64_t off = out.size();
        for (int i = 0; i < 8; i++) out[tableAt + (size_t)blk * 8 + i] = (uint8_t)(off >> (8 * i));
        put32(out, (uint32_t)y0);
        if (clen < n) { put32(out, (uint32_t)clen); out.insert(out.end(), comp.begin(), comp.begin() + clen); }
        else { put32(out, (uint32_t)n); out.insert(out.end(), raw.begin(), raw.end()); }
    }
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok;
}

// ---- PNG 8-bit RGB ---------------------------------------------------------------------------
inline bool SavePNG(const std::string& path, const float* film, int width, int height) {
    using namespace detail;
    const double Exp = 1.0 / 2.2;  // basic.hpp:633
    std::vector<uint8_t> raw; raw.reserve(((size_t)width * 3 + 1) * height);
    for (int y = height - 1; y >= 0; y--) {  // PNG is top-down
        raw.push_back(0);  // filter: none
        for (int x = 0; x < width; x++)
            for (int c = 0; c < 3; c++) {
                const double v = std::pow((double)film[((size_t)y * width + x) * 3 + c], Exp) * 255.0;
                int q = std::isnan(v) ? 0 : (v > 255.0 ? 255 : (int)v);
                raw.push_back((uint8_t)(q < 0 ? 0 : q > 255 ? 255 : q));
            }
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), Z_DEFAULT_COMPRESSION) != Z_OK) return false;
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    auto chunk = [&](const char* type, const uint8_t* data, size_t len) {
        put32be(out, (uint32_t)len);
        const size_t at = out.size();
        for (int i = 0; i < 4; i++) out.push_back((uint8_t)type[i]);
        out.insert(out.end(), data, data + len);
        put32be(out, (uint32_t)crc32(0L, out.data() + at, (uInt)(len + 4)));
    };
    std::vector<uint8_t> ihdr;
    put32be(ihdr, (uint32_t)width); put32be(ihdr, (uint32_t)height);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk("IHDR", ihdr.data(), ihdr.size());
    chunk("IDAT", comp.data(), clen);
    chunk("IEND", nullptr, 0);
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok;
}

// SaveImage, basic.hpp:506-672
inline bool SaveImage(const std::string& path, const float* film, int width, int height) {
    if (!detail::make_parent_dirs(path)) { NGI_LOG_WARN("Failed to create output directory : " + path); return false; }
    const std::string ext = detail::extension(path);
    bool ok;
    if (ext == ".hdr") ok = SaveHDR(path, film, width, height);
    else if (ext == ".exr") ok = SaveEXR(path, film, width, height);
    else if (ext == ".png") ok = SavePNG(path, film, width, height);
    else { NGI_LOG_ERROR("Invalid extension: " + ext); return false; }
    if (!ok) { NGI_LOG_ERROR("Failed to save image : " + path); return false; }
    NGI_LOG_INFO("Successfully saved to " + path);
    return true;
}

}  // namespace ngi
