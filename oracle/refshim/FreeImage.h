// Stand-in for the FreeImage subset the reference uses (Texture::Load rt.hpp:168-258, SaveImage basic.hpp:506-672).
// Written for this repository; see README.md. Bitmaps live in memory with FreeImage's bottom-up scanlines; files go through
// this repository's own readers / writers (nanogi_b200/host/image_io.hpp): .hdr (Radiance RGBE) and 8-bit .png on save;
// PNG / HDR / PPM / PFM on load, always presented as FIT_RGBF (an 8-bit channel v arrives as v / 255.0f, the same value
// the reference computes from FIT_BITMAP data at rt.hpp:245-250).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "../../nanogi_b200/host/logger.hpp"
#include "../../nanogi_b200/host/image_io.hpp"

typedef uint8_t BYTE;
typedef int32_t BOOL;
enum FREE_IMAGE_FORMAT { FIF_UNKNOWN = -1, FIF_BMP = 0, FIF_PNG = 13, FIF_PPM = 14, FIF_HDR = 26, FIF_PFM = 32 };
enum FREE_IMAGE_TYPE { FIT_UNKNOWN = 0, FIT_BITMAP = 1, FIT_RGBF = 11, FIT_RGBAF = 12 };
struct FIRGBF { float red, green, blue; };
struct FIRGBAF { float red, green, blue, alpha; };
#define FI_RGBA_RED 2
#define FI_RGBA_GREEN 1
#define FI_RGBA_BLUE 0
#define FI_RGBA_RED_MASK 0x00FF0000
#define FI_RGBA_GREEN_MASK 0x0000FF00
#define FI_RGBA_BLUE_MASK 0x000000FF
#define HDR_DEFAULT 0
#define PNG_DEFAULT 0

struct FIBITMAP { FREE_IMAGE_TYPE type; int width, height, bpp; std::vector<BYTE> data; size_t pitch; };

inline FIBITMAP* FreeImage_AllocateT(FREE_IMAGE_TYPE type, int width, int height, int bpp = 8, unsigned = 0, unsigned = 0, unsigned = 0) {
    FIBITMAP* b = new FIBITMAP;
    b->type = type; b->width = width; b->height = height;
    b->bpp = type == FIT_RGBF ? 96 : type == FIT_RGBAF ? 128 : bpp;
    b->pitch = ((size_t)width * b->bpp / 8 + 3) & ~(size_t)3;
    b->data.assign(b->pitch * height, 0);
    return b;
}
inline FIBITMAP* FreeImage_Allocate(int width, int height, int bpp, unsigned r = 0, unsigned g = 0, unsigned b = 0) { return FreeImage_AllocateT(FIT_BITMAP, width, height, bpp, r, g, b); }
inline void FreeImage_Unload(FIBITMAP* b) { delete b; }
inline BYTE* FreeImage_GetScanLine(FIBITMAP* b, int y) { return b->data.data() + b->pitch * (size_t)y; }   // scanline 0 = bottom row
inline unsigned FreeImage_GetWidth(FIBITMAP* b) { return (unsigned)b->width; }
inline unsigned FreeImage_GetHeight(FIBITMAP* b) { return (unsigned)b->height; }
inline unsigned FreeImage_GetBPP(FIBITMAP* b) { return (unsigned)b->bpp; }
inline FREE_IMAGE_TYPE FreeImage_GetImageType(FIBITMAP* b) { return b->type; }
inline BOOL FreeImage_FlipVertical(FIBITMAP* b) {
    std::vector<BYTE> tmp(b->pitch);
    for (int y = 0; y < b->height / 2; y++) {
        BYTE* a = FreeImage_GetScanLine(b, y); BYTE* c = FreeImage_GetScanLine(b, b->height - 1 - y);
        std::memcpy(tmp.data(), a, b->pitch); std::memcpy(a, c, b->pitch); std::memcpy(c, tmp.data(), b->pitch);
    }
    return 1;
}
inline FREE_IMAGE_FORMAT FreeImage_GetFIFFromFilename(const char* path) {
    const std::string p(path); const size_t d = p.find_last_of('.');
    const std::string e = d == std::string::npos ? "" : p.substr(d);
    return e == ".png" ? FIF_PNG : e == ".hdr" ? FIF_HDR : e == ".ppm" ? FIF_PPM : e == ".pfm" ? FIF_PFM : FIF_UNKNOWN;
}
inline FREE_IMAGE_FORMAT FreeImage_GetFileType(const char* path, int = 0) { return FreeImage_GetFIFFromFilename(path); }
inline BOOL FreeImage_FIFSupportsReading(FREE_IMAGE_FORMAT f) { return f != FIF_UNKNOWN; }
inline FIBITMAP* FreeImage_Load(FREE_IMAGE_FORMAT, const char* path, int = 0) {
    int w = 0, h = 0; std::vector<float> rgb; std::string err;
    if (!ngi::LoadImageRGB(path, w, h, rgb, err)) return nullptr;      // rgb: row 0 = top
    FIBITMAP* b = FreeImage_AllocateT(FIT_RGBF, w, h);
    for (int y = 0; y < h; y++) std::memcpy(FreeImage_GetScanLine(b, y), &rgb[(size_t)(h - 1 - y) * w * 3], (size_t)w * 12);
    return b;
}
inline BOOL FreeImage_Save(FREE_IMAGE_FORMAT fif, FIBITMAP* b, const char* path, int = 0) {
    // both writers of image_io.hpp take a film with row 0 = bottom, which is exactly FreeImage's scanline order
    std::vector<float> film((size_t)b->width * b->height * 3);
    if (fif == FIF_HDR && b->type == FIT_RGBF) {
        for (int y = 0; y < b->height; y++) std::memcpy(&film[(size_t)y * b->width * 3], FreeImage_GetScanLine(b, y), (size_t)b->width * 12);
        return ngi::SaveHDR(path, film.data(), b->width, b->height) ? 1 : 0;
    }
    if (fif == FIF_PNG && b->type == FIT_BITMAP && b->bpp == 24) {
        // the bitmap already holds the reference's tone-mapped bytes; ngi::SavePNG applies the same pow(x, 1/2.2) * 255 to a float
        // film, so hand it the inverse: ((v + 0.5) / 255)^2.2 lands back on byte v
        for (int y = 0; y < b->height; y++) {
            const BYTE* s = FreeImage_GetScanLine(b, y);
            for (int x = 0; x < b->width; x++) {
                const BYTE c[3] = {s[3 * x + FI_RGBA_RED], s[3 * x + FI_RGBA_GREEN], s[3 * x + FI_RGBA_BLUE]};
                for (int k = 0; k < 3; k++) film[((size_t)y * b->width + x) * 3 + k] = c[k] == 0 ? 0.0f : (float)std::pow((c[k] + 0.5) / 255.0, 2.2);
            }
        }
        return ngi::SavePNG(path, film.data(), b->width, b->height) ? 1 : 0;
    }
    return 0;
}
