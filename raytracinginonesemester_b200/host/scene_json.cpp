#include "scene_json.hpp"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>

namespace rtb200 {
namespace {

struct J {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<J> arr;
    std::vector<std::pair<std::string, J>> obj;   // first match wins, like the reference's json_get
    const J* get(const char* key) const {
        if (kind != Obj) return nullptr;
        for (auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

struct Reader {
    const std::string& s; size_t i = 0; std::string err;
    explicit Reader(const std::string& t) : s(t) {}
    void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
    bool fail(const std::string& m) { if (err.empty()) err = m + " at offset " + std::to_string(i); return false; }
    bool value(J& out) {
        ws();
        if (i >= s.size()) return fail("unexpected end");
        char c = s[i];
        if (c == '{') return object(out);
        if (c == '[') return array(out);
        if (c == '"') { out.kind = J::Str; return string(out.str); }
        if (!s.compare(i, 4, "true")) { out.kind = J::Bool; out.b = true; i += 4; return true; }
        if (!s.compare(i, 5, "false")) { out.kind = J::Bool; out.b = false; i += 5; return true; }
        if (!s.compare(i, 4, "null")) { out.kind = J::Null; i += 4; return true; }
        char* e = nullptr;
        double v = std::strtod(s.c_str() + i, &e);
        if (e == s.c_str() + i) return fail("bad value");
        out.kind = J::Num; out.num = v; i = (size_t)(e - s.c_str());
        return true;
    }
    bool string(std::string& out) {
        ++i;
        out.clear();
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\' && i + 1 < s.size()) {
                char n = s[i + 1];
                out += n == 'n' ? '\n' : n == 't' ? '\t' : n;
                i += 2;
            } else out += s[i++];
        }
        if (i >= s.size()) return fail("unterminated string");
        ++i;
        return true;
    }
    bool array(J& out) {
        out.kind = J::Arr; ++i; ws();
        if (i < s.size() && s[i] == ']') { ++i; return true; }
        while (true) {
            J v;
            if (!value(v)) return false;
            out.arr.push_back(std::move(v));
            ws();
            if (i < s.size() && s[i] == ',') { ++i; continue; }
            if (i < s.size() && s[i] == ']') { ++i; return true; }
            return fail("expected ',' or ']'");
        }
    }
    bool object(J& out) {
        out.kind = J::Obj; ++i; ws();
        if (i < s.size() && s[i] == '}') { ++i; return true; }
        while (true) {
            ws();
            if (i >= s.size() || s[i] != '"') return fail("expected key");
            std::string k;
            if (!string(k)) return false;
            ws();
            if (i >= s.size() || s[i] != ':') return fail("expected ':'");
            ++i;
            J v;
            if (!value(v)) return false;
            out.obj.emplace_back(std::move(k), std::move(v));
            ws();
            if (i < s.size() && s[i] == ',') { ++i; continue; }
            if (i < s.size() && s[i] == '}') { ++i; return true; }
            return fail("expected ',' or '}'");
        }
    }
};

bool vec3(const J* v, float out[3]) {
    if (!v || v->kind != J::Arr || v->arr.size() != 3) return false;
    for (auto& e : v->arr) if (e.kind != J::Num) return false;
    for (int k = 0; k < 3; ++k) out[k] = (float)v->arr[k].num;
    return true;
}
bool number(const J* v, double& out) { if (!v || v->kind != J::Num) return false; out = v->num; return true; }

void read_light(const J& o, rt_light& l) {
    vec3(o.get("position"), l.position);
    vec3(o.get("color"), l.color);
    double d;
    if (number(o.get("intensity"), d)) l.intensity = (int)d;     // Light::intensity is an int (scene.h:24, quirk Q4)
}

} // namespace

rt_material default_material() {
    rt_material m{};
    m.albedo[0] = m.albedo[1] = m.albedo[2] = 0.8f; m.kd = 1.0f;
    m.specular_color[0] = m.specular_color[1] = m.specular_color[2] = 0.04f;
    m.ks = 0.0f; m.shininess = 32.0f; m.kr = 0.0f;
    return m;
}

bool parse_scene_text(const std::string& text, SceneDesc& sc, std::string* err) {
    Reader rd(text);
    J root;
    if (!rd.value(root)) { if (err) *err = rd.err; return false; }
    if (root.kind != J::Obj) { if (err) *err = "Root is not an object"; return false; }
    double d;
    if (const J* st = root.get("settings")) {
        if (number(st->get("max_bounces"), d)) sc.max_depth = (int)d;
        if (number(st->get("spp"), d)) { sc.spp = (int)d; if (sc.spp < 1) sc.spp = 1; }
        const J* db = st->get("diffuse_bounce");
        if (db && db->kind == J::Bool) sc.diffuse_bounce = db->b;
    }
    vec3(root.get("miss_color"), sc.miss_color);
    if (const J* cam = root.get("camera")) {
        if (number(cam->get("focal_length_mm"), d)) sc.focal_length_mm = d;
        if (number(cam->get("sensor_height_mm"), d)) sc.sensor_height_mm = d;
        if (number(cam->get("pixel_width"), d)) sc.pixel_width = (int)d;
        if (number(cam->get("pixel_height"), d)) sc.pixel_height = (int)d;
        vec3(cam->get("position"), sc.cam_pos);
        vec3(cam->get("look_at"), sc.cam_look_at);
        vec3(cam->get("up"), sc.cam_up);
    }
    sc.lights.clear();
    auto fresh_light = [] { rt_light l{}; l.color[0] = l.color[1] = l.color[2] = 1.0f; l.intensity = 1; return l; };
    const J* lights = root.get("lights");
    if (lights && lights->kind == J::Arr)
        for (auto& item : lights->arr) {
            if (item.kind != J::Obj) continue;
            rt_light l = fresh_light();
            read_light(item, l);
            sc.lights.push_back(l);
        }
    if (sc.lights.empty()) {
        const J* light = root.get("light");
        if (light && light->kind == J::Obj) { rt_light l = fresh_light(); read_light(*light, l); sc.lights.push_back(l); }
    }
    const J* arr = root.get("scene");
    if (!arr || arr->kind != J::Arr) { if (err) *err = "Missing 'scene' array"; return false; }
    sc.objects.clear();
    for (auto& item : arr->arr) {
        if (item.kind != J::Obj) continue;
        SceneObjectDesc o;
        o.material = default_material();
        const J* v;
        if ((v = item.get("name")) && v->kind == J::Str) o.name = v->str;
        if ((v = item.get("type")) && v->kind == J::Str) o.type = v->str;
        if ((v = item.get("path")) && v->kind == J::Str) o.path = v->str;
        const J* tr = item.get("transform");
        if (tr && tr->kind == J::Obj) {
            vec3(tr->get("position"), o.position);
            vec3(tr->get("rotation"), o.rotation);
            vec3(tr->get("scale"), o.scale);
        }
        const J* mt = item.get("material");
        if (mt && mt->kind == J::Obj) {
            vec3(mt->get("albedo"), o.material.albedo);
            vec3(mt->get("specular_color"), o.material.specular_color);
            vec3(mt->get("emission"), o.material.emission);
            if (number(mt->get("kd"), d)) o.material.kd = (float)d;
            if (number(mt->get("ks"), d)) o.material.ks = (float)d;
            if (number(mt->get("shininess"), d)) o.material.shininess = (float)d;
            if (number(mt->get("kr"), d)) o.material.kr = (float)d;
        }
        if (!o.path.empty()) sc.objects.push_back(o);
    }
    if (sc.objects.empty()) { if (err) *err = "Scene contains no valid objects"; return false; }
    return true;
}

bool load_scene_file(const std::string& path, SceneDesc& out, std::string* err) {
    std::ifstream f(path);
    if (!f) { if (err) *err = "Failed to open scene file: " + path; return false; }
    std::stringstream buf;
    buf << f.rdbuf();
    return parse_scene_text(buf.str(), out, err);
}

} // namespace rtb200
