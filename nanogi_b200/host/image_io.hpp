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
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
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
// ---- image readers for TexR textures ---------------------------------------------------------------------------
// Host-side mirror of Texture::Load (reference include/nanogi/rt.hpp:168-258), which goes through FreeImage and keeps
// float RGB with ROW 0 = TOP scanline (FreeImage's bottom-up bitmap flipped, :214-217); 8-bit channels become v / 255
// with no gamma (:245-250). FreeImage is not available here, so the formats are read directly: PNG (8-bit RGB / RGBA /
// grey, non-interlaced; zlib inflate + the five scanline filters), Radiance .hdr (RGBE, flat or new-style RLE), binary
// PPM (P6) and PFM (PF / Pf). Anything else is an error, like an unsupported FreeImage type (:205-211).
namespace detail {
inline bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    out.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
    return true;
}
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline bool decode_png(const std::vector<uint8_t>& f, int& w, int& h, std::vector<float>& rgb, std::string& err) {
    size_t pos = 8;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat;
    w = h = 0;
    while (pos + 12 <= f.size()) {
        const uint32_t len = be32(&f[pos]);
        const std::string type((const char*)&f[pos + 4], 4);
        if (pos + 12 + len > f.size()) { err = "truncated PNG"; return false; }
        const uint8_t* d = &f[pos + 8];
        if (type == "IHDR") {
            if (len < 13) { err = "malformed PNG (IHDR shorter than 13 bytes)"; return false; }
            w = (int)be32(d); h = (int)be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12];
        }
        else if (type == "IDAT") idat.insert(idat.end(), d, d + len);
        else if (type == "IEND") break;
        pos += 12 + len;
    }
    if (w <= 0 || h <= 0) { err = "PNG without IHDR"; return false; }
    if (depth != 8 || interlace != 0 || !(ctype == 0 || ctype == 2 || ctype == 4 || ctype == 6)) { err = "unsupported PNG (need 8-bit, non-interlaced, grey/RGB/RGBA)"; return false; }
    const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 4 ? 2 : 4;
    const size_t stride = (size_t)w * ch;
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    uLongf rawlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) { err = "PNG inflate failed"; return false; }
    std::vector<uint8_t> img(stride * (size_t)h);
    for (int y = 0; y < h; y++) {
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t* in = &raw[(stride + 1) * y + 1];
        uint8_t* out = &img[stride * y];
        const uint8_t* up = y ? &img[stride * (y - 1)] : nullptr;
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= (size_t)ch ? out[i - ch] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)ch) ? up[i - ch] : 0;
            int v = in[i];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) / 2; break;
                case 4: { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
                default: err = "bad PNG filter"; return false;
            }
            out[i] = (uint8_t)v;
        }
    }
    rgb.resize((size_t)w * h * 3);
    for (size_t i = 0; i < (size_t)w * h; i++)
        for (int k = 0; k < 3; k++) rgb[3 * i + k] = (float)img[i * ch + (ch >= 3 ? k : 0)] / 255.0f;   // rt.hpp:245-250
    return true;
}

inline bool decode_hdr(const std::vector<uint8_t>& f, int& w, int& h, std::vector<float>& rgb, std::string& err) {
    size_t pos = 0;
    auto line = [&]() { std::string s; while (pos < f.size() && f[pos] != '\n') s.push_back((char)f[pos++]); pos++; return s; };
    std::string l = line();
    while (pos < f.size() && !l.empty()) l = line();          // header ends with an empty line
    l = line();
    if (std::sscanf(l.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) { err = "unsupported .hdr orientation (need -Y h +X w)"; return false; }
    std::vector<uint8_t> px((size_t)w * h * 4);
    for (int y = 0; y < h; y++) {
        uint8_t* row = &px[(size_t)y * w * 4];
        if (pos + 4 <= f.size() && w >= 8 && w < 32768 && f[pos] == 2 && f[pos + 1] == 2 && ((f[pos + 2] << 8) | f[pos + 3]) == w) {
            pos += 4;                                        // new-style RLE: the four channels separately
            for (int c = 0; c < 4; c++) {
                int x = 0;
                while (x < w) {
                    if (pos >= f.size()) { err = "truncated .hdr"; return false; }
                    int n = f[pos++];
                    if (n > 128) { n -= 128; if (pos >= f.size() || x + n > w) { err = "bad .hdr run"; return false; } const uint8_t v = f[pos++]; while (n--) row[4 * x++ + c] = v; }
                    else { if (pos + n > f.size() || x + n > w) { err = "bad .hdr run"; return false; } while (n--) row[4 * x++ + c] = f[pos++]; }
                }
            }
        } else {
            if (pos + (size_t)w * 4 > f.size()) { err = "truncated .hdr"; return false; }
            std::memcpy(row, &f[pos], (size_t)w * 4); pos += (size_t)w * 4;
        }
    }
    rgb.resize((size_t)w * h * 3);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        const int e = px[4 * i + 3];
        const float s = e ? std::ldexp(1.0f, e - 136) : 0.0f;
        for (int k = 0; k < 3; k++) rgb[3 * i + k] = e ? ((float)px[4 * i + k] + 0.5f) * s : 0.0f;
    }
    return true;
}

inline bool decode_pnm(const std::vector<uint8_t>& f, int& w, int& h, std::vector<float>& rgb, std::string& err) {
    size_t pos = 0;
    auto token = [&]() {
        std::string s;
        while (pos < f.size()) { if (f[pos] == '#') { while (pos < f.size() && f[pos] != '\n') pos++; } else if (std::isspace(f[pos])) pos++; else break; }
        while (pos < f.size() && !std::isspace(f[pos])) s.push_back((char)f[pos++]);
        return s;
    };
    const std::string magic = token();
    w = std::atoi(token().c_str()); h = std::atoi(token().c_str());
    const double third = std::atof(token().c_str());
    pos++;                                                   // single whitespace before the raster
    if (w <= 0 || h <= 0) { err = "bad PNM header"; return false; }
    rgb.resize((size_t)w * h * 3);
    if (magic == "P6") {
        if (third != 255 || pos + (size_t)w * h * 3 > f.size()) { err = "unsupported PPM (need maxval 255)"; return false; }
        for (size_t i = 0; i < (size_t)w * h * 3; i++) rgb[i] = (float)f[pos + i] / 255.0f;
        return true;
    }
    if (magic == "PF" || magic == "Pf") {                    // PFM: little endian when the scale is negative, rows bottom-up
        const int ch = magic == "PF" ? 3 : 1;
        if (third >= 0 || pos + (size_t)w * h * ch * 4 > f.size()) { err = "unsupported PFM (need little endian)"; return false; }
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                for (int k = 0; k < 3; k++) {
                    float v; std::memcpy(&v, &f[pos + (((size_t)(h - 1 - y) * w + x) * ch + (ch == 3 ? k : 0)) * 4], 4);
                    rgb[((size_t)y * w + x) * 3 + k] = v;
                }
        return true;
    }
    err = "unsupported PNM type " + magic;
    return false;
}
}  // namespace detail

// rgb: width * height * 3 floats, row 0 = top scanline
inline bool LoadImageRGB(const std::string& path, int& width, int& height, std::vector<float>& rgb, std::string& err) {
    std::vector<uint8_t> f;
    if (!detail::read_file(path, f) || f.size() < 8) { err = "Failed to load an image " + path; return false; }
    static const uint8_t png_sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    bool ok;
    if (std::memcmp(f.data(), png_sig, 8) == 0) ok = detail::decode_png(f, width, height, rgb, err);
    else if (f[0] == '#' && f[1] == '?') ok = detail::decode_hdr(f, width, height, rgb, err);
    else if (f[0] == 'P' && (f[1] == '6' || f[1] == 'F' || f[1] == 'f')) ok = detail::decode_pnm(f, width, height, rgb, err);
    else { err = "Unknown image format"; ok = false; }       // rt.hpp:176-183
    if (!ok) err = path + ": " + err;
    return ok;
}

// [b200, additive] lossless float film: PFM "PF", little endian, rows bottom-up = the film's own row order. Used by
// --resume-from (a resumable film needs the exact float values; .hdr is RGBE and .png 8-bit).
inline bool SavePFM(const std::string& path, const float* film, int width, int height) {
    FILE* fp = std::fopen(path.c_str(), "wb");
    if (!fp) return false;
    std::fprintf(fp, "PF\n%d %d\n-1.0\n", width, height);
    const size_t n = (size_t)width * height * 3;
    const bool ok = std::fwrite(film, sizeof(float), n, fp) == n;
    std::fclose(fp);
    return ok;
}

inline bool SaveImage(const std::string& path, const float* film, int width, int height) {
    if (!detail::make_parent_dirs(path)) { NGI_LOG_WARN("Failed to create output directory : " + path); return false; }
    const std::string ext = detail::extension(path);
    bool ok;
    if (ext == ".hdr") ok = SaveHDR(path, film, width, height);
    else if (ext == ".pfm") ok = SavePFM(path, film, width, height);
    else if (ext == ".exr") ok = SaveEXR(path, film, width, height);
    else if (ext == ".png") ok = SavePNG(path, film, width, height);
    else { NGI_LOG_ERROR("Invalid extension: " + ext); return false; }
    if (!ok) { NGI_LOG_ERROR("Failed to save image : " + path); return false; }
    NGI_LOG_INFO("Successfully saved to " + path);
    return true;
}

}  // namespace ngi
