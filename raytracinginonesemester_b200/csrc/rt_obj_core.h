// rt_obj_core.h — the per-line parsing functions of the device OBJ parser (rt_obj_device.cu), host/device compilable so that
// tests/emul can run them on the CPU against the host loader and glibc's strtof (tests/test_obj_core_on_host.py).
// They restate the cursor movements of the reference loader's line parser (GPUandCPU/include/MeshOBJ.h:260-427, as restated in
// host/mesh_ingest.cpp) and strtof's decimal syntax.
#pragma once

#include <stdint.h>
#include <string.h>

#include "rt_math.h"

struct Corner { int v, t, n; };

// Cursor over one line: [i, end) of the text; reads past the end give '\0' (the loader's fgets buffer is NUL-terminated).
struct DCur {
    const char* t; uint32_t i, end;
    RT_HD char at(uint32_t k) const { return k < end ? t[k] : '\0'; }
    RT_HD char c() const { return at(i); }
    RT_HD void ws() { while (c() == ' ' || c() == '\t') ++i; }
    RT_HD bool eol() const { const char x = c(); return x == '\0' || x == '\n' || x == '\r'; }
    RT_HD bool integer(int& v) {
        ws();
        bool neg = false;
        if (c() == '-') { neg = true; ++i; }
        if (c() < '0' || c() > '9') return false;
        unsigned a = 0;
        while (c() >= '0' && c() <= '9') { a = a * 10u + (unsigned)(c() - '0'); ++i; }
        v = neg ? -(int)a : (int)a;
        return true;
    }
    RT_HD void skip_token() { while (c() != '\0' && c() != '\n' && c() != ' ' && c() != '\t') ++i; }
};

RT_HD int resolve(int idx, uint32_t count) { return idx < 0 ? (int)count + idx : idx - 1; }

// "v", "v/t", "v//n", "v/t/n" — the same cursor movements as the loader's corner parser
RT_HD bool parse_corner(DCur& c, Corner& k, uint32_t nv, uint32_t nt, uint32_t nn) {
    int a = 0;
    if (!c.integer(a)) return false;
    k.v = resolve(a, nv); k.t = -1; k.n = -1;
    if (c.c() != '/') return true;
    ++c.i;
    if (c.c() == '/') {
        ++c.i;
        int n = 0;
        if (!c.integer(n)) return false;
        k.n = resolve(n, nn);
        return true;
    }
    int t = 0;
    if (c.integer(t)) k.t = resolve(t, nt);
    if (c.c() != '/') return true;
    ++c.i;
    int n = 0;
    if (c.integer(n)) k.n = resolve(n, nn);
    return true;
}

RT_HD int parse_face(DCur c, Corner k[4], uint32_t nv, uint32_t nt, uint32_t nn) {
    int n = 0;
    while (n < 4) {
        c.ws();
        if (c.c() == '\0' || c.c() == '\n') break;
        Corner q;
        if (!parse_corner(c, q, nv, nt, nn)) break;
        k[n++] = q;
        c.skip_token();
    }
    return n;
}

// 10^k, exact in fp64 for k <= 22
RT_HD double rt_obj_pow10(int k) {
    const double t[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    return t[k];
}

RT_HD long long rt_obj_double_bits(double d) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(d);
#else
    long long b; memcpy(&b, &d, sizeof b); return b;
#endif
}
RT_HD bool is_space(char x) { return x == ' ' || (x >= '\t' && x <= '\r'); }
RT_HD char lower(char x) { return (x >= 'A' && x <= 'Z') ? (char)(x + 32) : x; }

// strtof on the cursor.  0 = no conversion (cursor unchanged), 1 = value in `out`, 2 = syntax accepted but the value must be
// converted on the host (token [tok0, c.i)), 3 = a form this parser refuses (inf, nan, hexadecimal).
RT_HD int parse_real(DCur& c, float& out, uint32_t& tok0) {
    c.ws();
    uint32_t i = c.i;
    while (is_space(c.at(i))) ++i;
    tok0 = i;
    bool neg = false;
    if (c.at(i) == '+' || c.at(i) == '-') { neg = c.at(i) == '-'; ++i; }
    {
        const char a = lower(c.at(i)), b = lower(c.at(i + 1)), d = lower(c.at(i + 2));
        if ((a == 'i' && b == 'n' && d == 'f') || (a == 'n' && b == 'a' && d == 'n')) return 3;
        if (a == '0' && b == 'x') {
            const char h = lower(c.at(i + 2)), h2 = lower(c.at(i + 3));
            const bool hex1 = (h >= '0' && h <= '9') || (h >= 'a' && h <= 'f');
            const bool hex2 = (h2 >= '0' && h2 <= '9') || (h2 >= 'a' && h2 <= 'f');
            if (hex1 || (h == '.' && hex2)) return 3;
        }
    }
    unsigned long long m = 0;
    int nd = 0, e10 = 0;
    bool any = false, dot = false, dropped = false;
    for (;; ++i) {
        const char x = c.at(i);
        if (x >= '0' && x <= '9') {
            any = true;
            const unsigned d = (unsigned)(x - '0');
            if (m == 0 && d == 0) { if (dot) --e10; }
            else if (nd < 19) { m = m * 10ull + d; ++nd; if (dot) --e10; }
            else { if (d) dropped = true; if (!dot) ++e10; }
        } else if (x == '.' && !dot) dot = true;
        else break;
    }
    if (!any) return 0;
    if (lower(c.at(i)) == 'e') {
        uint32_t j = i + 1;
        bool eneg = false;
        if (c.at(j) == '+' || c.at(j) == '-') { eneg = c.at(j) == '-'; ++j; }
        if (c.at(j) >= '0' && c.at(j) <= '9') {
            int ex = 0;
            while (c.at(j) >= '0' && c.at(j) <= '9') { if (ex < 100000) ex = ex * 10 + (c.at(j) - '0'); ++j; }
            e10 += eneg ? -ex : ex;
            i = j;
        }
    }
    c.i = i;
    if (m == 0) { out = neg ? -0.0f : 0.0f; return 1; }
    if (e10 < -44 || e10 > 44) return 2;
    // m < 2^53, no dropped digits, |e| <= 22: both operands exact, ONE correctly rounded fp64 operation.  Otherwise — 17..19 digits
    // (the integer rounds on its way to fp64), more digits with the tail dropped (relative error < 1e-18), 22 < |e| <= 44 (two
    // exact powers of ten, two operations) — d is within 2.5 ulp of the value.
    const bool exact = !dropped && m < (1ull << 53) && e10 >= -22 && e10 <= 22;
    double d = (double)m;
    {
        const int a = e10 < 0 ? -e10 : e10, a1 = a > 22 ? 22 : a, a2 = a - a1;
        d = e10 >= 0 ? d * rt_obj_pow10(a1) : d / rt_obj_pow10(a1);
        if (a2) d = e10 >= 0 ? d * rt_obj_pow10(a2) : d / rt_obj_pow10(a2);
    }
    if (d < 1.1754943508222875e-38 || d > 3.4028234663852886e38) return 2;                 // sub-normal / overflowing fp32 results
    // Narrowing is right unless an fp32 rounding boundary (the midpoint of two neighbouring floats: low 29 fraction bits
    // 1000...0) lies within the error of d: exactly on it in the exact case, within 8 ulp otherwise.
    const long long low = rt_obj_double_bits(d) & 0x1fffffffll, off = low - 0x10000000ll;
    if (exact ? off == 0 : (off >= -8 && off <= 8)) return 2;
    const float f = (float)d;
    out = neg ? -f : f;
    return 1;
}

