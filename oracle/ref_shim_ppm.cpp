// ref_shim_ppm.cpp — C-ABI doorway into the reference's UNMODIFIED ppm_p6_lib
// (HW1/ppm_p6_lib/src/ppm_p6.cpp compiled in place next to this file; output
// oracle/_ref/libref_ppm.so, git-ignored).  Used by the host driver and the tests so the
// PPM P6 bytes come from the reference's own writer / quantiser (ppm_p6.cpp:137-155,257-301).
#include "ppm_p6.hpp"

#include <cstdint>
#include <cstring>
#include <string>

extern "C" {

// rgb: float[3*w*h] row-major.  Returns 0 on success, -1 on failure (message in err, if given).
int ref_ppm_write_rgbf(const char* path, int w, int h, const float* rgb, int maxval, int clamp,
                       int gamma2, int flip_y, char* err, int err_cap) {
    ppm_p6::Image img(w, h);
    auto& px = img.pixels();
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        px[i].r = rgb[3 * i]; px[i].g = rgb[3 * i + 1]; px[i].b = rgb[3 * i + 2];
    }
    ppm_p6::WriteOptions opt;
    opt.maxval = maxval; opt.clamp = clamp != 0; opt.gamma2 = gamma2 != 0; opt.flip_y = flip_y != 0;
    std::string e;
    bool ok = ppm_p6::write_p6(path, img, opt, &e);
    if (!ok && err && err_cap > 0) { std::strncpy(err, e.c_str(), (size_t)err_cap - 1); err[err_cap - 1] = 0; }
    return ok ? 0 : -1;
}

// Reads a P6 file back as doubles -> float rgb (caller passes capacity in floats).
int ref_ppm_read_rgbf(const char* path, int* w, int* h, float* rgb, uint64_t cap_floats) {
    ppm_p6::Image img;
    std::string e;
    if (!ppm_p6::read_p6(path, img, &e)) return -1;
    *w = img.width(); *h = img.height();
    size_t n = (size_t)img.width() * img.height();
    if (rgb) {
        if (cap_floats < 3 * n) return -2;
        const auto& px = img.pixels();
        for (size_t i = 0; i < n; ++i) { rgb[3 * i] = (float)px[i].r; rgb[3 * i + 1] = (float)px[i].g; rgb[3 * i + 2] = (float)px[i].b; }
    }
    return 0;
}

} // extern "C"
