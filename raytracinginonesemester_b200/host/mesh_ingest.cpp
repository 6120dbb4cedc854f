#include "mesh_ingest.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <tuple>

namespace rtb200 {
namespace {

// Cursor over one text line.
struct Cur {
    const char* p;
    void ws() { while (*p == ' ' || *p == '\t') ++p; }
    bool eol() const { return *p == '\0' || *p == '\n' || *p == '\r'; }
    bool integer(int& v) {
        ws();
        bool neg = false;
        if (*p == '-') { neg = true; ++p; }
        if (*p < '0' || *p > '9') return false;
        int a = 0;
        while (*p >= '0' && *p <= '9') { a = a * 10 + (*p - '0'); ++p; }
        v = neg ? -a : a;
        return true;
    }
    bool real(float& v) {
        ws();
        char* e = nullptr;
        v = std::strtof(p, &e);
        if (e == p) return false;
        p = e;
        return true;
    }
    void skip_token() { while (*p != '\0' && *p != '\n' && *p != ' ' && *p != '\t') ++p; }
};

struct Corner { int v = -1, t = -1, n = -1; };

int resolve(int idx, size_t count) { return idx < 0 ? (int)count + idx : idx - 1; }

// "v", "v/t", "v//n", "v/t/n"
bool parse_corner(Cur& c, Corner& k, size_t nv, size_t nt, size_t nn) {
    int a = 0;
    if (!c.integer(a)) return false;
    k = Corner{};
    k.v = resolve(a, nv);
    if (*c.p != '/') return true;
    ++c.p;
    if (*c.p == '/') {
        ++c.p;
        int n = 0;
        if (!c.integer(n)) return false;
        k.n = resolve(n, nn);
        return true;
    }
    int t = 0;
    if (c.integer(t)) k.t = resolve(t, nt);
    if (*c.p != '/') return true;
    ++c.p;
    int n = 0;
    if (c.integer(n)) k.n = resolve(n, nn);
    return true;
}

} // namespace

bool load_obj(const std::string& path, HostMesh& out, int& next_object_id, std::string* err) {
    out = HostMesh{};
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { if (err) *err = "cannot open " + path; return false; }
    std::vector<float> rp, rt, rn;                       // raw v / vt / vn
    bool has_uv = false, has_nrm = false, tag_seen = false;
    int obj_id = next_object_id;
    std::map<std::tuple<int, int, int>, uint32_t> seen;  // (v,t,n) -> unified vertex
    auto fail = [&](const char* why) { std::fclose(f); if (err) *err = path + ": " + why; return false; };
    auto vertex = [&](const Corner& k) -> uint32_t {
        auto key = std::make_tuple(k.v, k.t, k.n);
        auto it = seen.find(key);
        if (it != seen.end()) return it->second;
        uint32_t idx = (uint32_t)out.num_vertices();
        seen.emplace(key, idx);
        if (k.v < 0 || (size_t)k.v >= rp.size() / 3) throw 1;
        out.positions.insert(out.positions.end(), rp.begin() + 3 * k.v, rp.begin() + 3 * k.v + 3);
        if (has_uv) {
            float uv[2] = {0.f, 0.f};
            if (k.t >= 0 && (size_t)k.t < rt.size() / 2) { uv[0] = rt[2 * k.t]; uv[1] = rt[2 * k.t + 1]; }
            out.uvs.insert(out.uvs.end(), uv, uv + 2);
        }
        if (has_nrm) {
            float n[3] = {0.f, 0.f, 0.f};
            if (k.n >= 0 && (size_t)k.n < rn.size() / 3) { n[0] = rn[3 * k.n]; n[1] = rn[3 * k.n + 1]; n[2] = rn[3 * k.n + 2]; }
            out.normals.insert(out.normals.end(), n, n + 3);
        }
        return idx;
    };
    char line[1024];
    try {
        while (std::fgets(line, sizeof line, f)) {
            Cur c{line};
            c.ws();
            if (c.eol() || *c.p == '#') continue;
            if (*c.p == 'o' || *c.p == 'g') {
                // every tag after the first starts a new object; the first one names the current
                // object unless faces were already emitted under the implicit one
                if (tag_seen || !out.indices.empty()) obj_id = ++next_object_id;
                tag_seen = true;
                continue;
            }
            if (c.p[0] == 'v' && (c.p[1] == ' ' || c.p[1] == '\t')) {
                c.p += 1;
                float x, y, z;
                if (!c.real(x) || !c.real(y) || !c.real(z)) return fail("bad 'v' line");
                rp.insert(rp.end(), {x, y, z});
                continue;
            }
            if (c.p[0] == 'v' && c.p[1] == 't' && (c.p[2] == ' ' || c.p[2] == '\t')) {
                c.p += 2;
                float u, v;
                if (!c.real(u) || !c.real(v)) return fail("bad 'vt' line");
                rt.insert(rt.end(), {u, v});
                has_uv = true;
                continue;
            }
            if (c.p[0] == 'v' && c.p[1] == 'n' && (c.p[2] == ' ' || c.p[2] == '\t')) {
                c.p += 2;
                float x, y, z;
                if (!c.real(x) || !c.real(y) || !c.real(z)) return fail("bad 'vn' line");
                rn.insert(rn.end(), {x, y, z});
                has_nrm = true;
                continue;
            }
            if (c.p[0] == 'f' && (c.p[1] == ' ' || c.p[1] == '\t')) {
                c.p += 1;
                Corner k[4];
                int n = 0;
                while (n < 4) {
                    c.ws();
                    if (*c.p == '\0' || *c.p == '\n') break;
                    Corner q;
                    if (!parse_corner(c, q, rp.size() / 3, rt.size() / 2, rn.size() / 3)) break;
                    if (q.t >= 0) has_uv = true;
                    if (q.n >= 0) has_nrm = true;
                    k[n++] = q;
                    c.skip_token();
                }
                if (n < 3) return fail("face with fewer than 3 vertices");
                uint32_t a = vertex(k[0]), b = vertex(k[1]), d = vertex(k[2]);
                out.indices.insert(out.indices.end(), {a, b, d});
                out.tri_obj_ids.push_back(obj_id);
                if (n == 4) {
                    uint32_t e = vertex(k[3]);
                    out.indices.insert(out.indices.end(), {a, d, e});
                    out.tri_obj_ids.push_back(obj_id);
                }
                continue;
            }
        }
    } catch (int) {
        return fail("face references a missing vertex");
    }
    std::fclose(f);
    if (out.positions.empty() || out.indices.empty()) { if (err) *err = path + ": no geometry"; return false; }
    ++next_object_id;
    if (has_uv && out.uvs.size() / 2 != out.num_vertices()) { if (err) *err = path + ": uv stream misaligned"; return false; }
    if (has_nrm && out.normals.size() / 3 != out.num_vertices()) { if (err) *err = path + ": normal stream misaligned"; return false; }
    return true;
}

namespace {
inline void rotate_xyz(float v[3], const float deg[3]) {
    const float k = 0.01745329251994329577f;
    const float rx = deg[0] * k, ry = deg[1] * k, rz = deg[2] * k;
    const float cx = std::cos(rx), sx = std::sin(rx), cy = std::cos(ry), sy = std::sin(ry), cz = std::cos(rz), sz = std::sin(rz);
    float x = v[0], y = v[1], z = v[2];
    float y1 = cx * y - sx * z, z1 = sx * y + cx * z;           // about X
    float x2 = cy * x + sy * z1, z2 = -sy * x + cy * z1;         // about Y
    float x3 = cz * x2 - sz * y1, y3 = sz * x2 + cz * y1;        // about Z
    v[0] = x3; v[1] = y3; v[2] = z2;
}
} // namespace

void transform_mesh(HostMesh& m, const float position[3], const float rotation_deg[3], const float scale[3]) {
    for (size_t i = 0; i < m.num_vertices(); ++i) {
        float v[3] = {m.positions[3 * i] * scale[0], m.positions[3 * i + 1] * scale[1], m.positions[3 * i + 2] * scale[2]};
        rotate_xyz(v, rotation_deg);
        for (int k = 0; k < 3; ++k) m.positions[3 * i + k] = v[k] + position[k];
    }
    for (size_t i = 0; i < m.normals.size() / 3; ++i) {
        float n[3] = {m.normals[3 * i], m.normals[3 * i + 1], m.normals[3 * i + 2]};
        for (int k = 0; k < 3; ++k) if (std::fabs(scale[k]) > 1e-8f) n[k] /= scale[k];
        rotate_xyz(n, rotation_deg);
        const float len2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
        if (len2 > 1e-12f) {
            const float inv = 1.0f / std::sqrt(len2);
            for (int k = 0; k < 3; ++k) m.normals[3 * i + k] = n[k] * inv;
        } else {
            m.normals[3 * i] = 0.f; m.normals[3 * i + 1] = 0.f; m.normals[3 * i + 2] = 1.f;
        }
    }
}

void append_mesh(HostMesh& dst, const HostMesh& src) {
    const uint32_t base = (uint32_t)dst.num_vertices();
    const bool dst_had_vertices = base != 0;
    dst.positions.insert(dst.positions.end(), src.positions.begin(), src.positions.end());
    if (!dst.normals.empty() || !src.normals.empty()) {
        if (dst.normals.empty() && dst_had_vertices) dst.normals.assign(3 * (size_t)base, 0.f);
        if (!src.normals.empty()) dst.normals.insert(dst.normals.end(), src.normals.begin(), src.normals.end());
        else dst.normals.resize(dst.normals.size() + src.positions.size(), 0.f);
    }
    if (!dst.uvs.empty() || !src.uvs.empty()) {
        if (dst.uvs.empty() && dst_had_vertices) dst.uvs.assign(2 * (size_t)base, 0.f);
        if (!src.uvs.empty()) dst.uvs.insert(dst.uvs.end(), src.uvs.begin(), src.uvs.end());
        else dst.uvs.resize(dst.uvs.size() + 2 * src.num_vertices(), 0.f);
    }
    for (uint32_t i : src.indices) dst.indices.push_back(i + base);
    dst.tri_obj_ids.insert(dst.tri_obj_ids.end(), src.tri_obj_ids.begin(), src.tri_obj_ids.end());
}

} // namespace rtb200
