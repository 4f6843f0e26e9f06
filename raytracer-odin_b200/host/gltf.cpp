// gltf.cpp — read_gltf / populate_scene (input.odin:13-259) and finish_scene (raytracer.odin:62-91)
// for the C++ host.  Arithmetic is f32 with every operation individually rounded (the file is
// compiled with -ffp-contract=off), sums taken left to right like a generic matrix product.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <thread>

#include "json.hpp"
#include "scene.hpp"

namespace orh {
namespace {

struct Mat4 { // m[c][r]: column-major like Odin's matrix[4,4]f32 and cgltf's float[16]
    float m[4][4];
};

Mat4 identity() {
    Mat4 r{};
    for (int i = 0; i < 4; i++) r.m[i][i] = 1.0f;
    return r;
}
Mat4 mul(const Mat4& a, const Mat4& b) { // (a*b)[r,c] = sum_k a[r,k] * b[k,c]
    Mat4 o;
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++)
            o.m[c][r] = ((a.m[0][r] * b.m[c][0] + a.m[1][r] * b.m[c][1]) + a.m[2][r] * b.m[c][2]) + a.m[3][r] * b.m[c][3];
    return o;
}
void mul_vec(const Mat4& t, const float v[4], float out[3]) { // (t * v).xyz
    for (int r = 0; r < 3; r++) out[r] = ((t.m[0][r] * v[0] + t.m[1][r] * v[1]) + t.m[2][r] * v[2]) + t.m[3][r] * v[3];
}
void normalize3(float v[3]) { // linalg.normalize: v / sqrt(dot(v, v))
    const float len = std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    v[0] /= len; v[1] /= len; v[2] /= len;
}

// cgltf_node_transform_local (called at input.odin:100)
Mat4 node_transform_local(const Json& node) {
    Mat4 o{};
    float* lm = &o.m[0][0];
    if (node.has("matrix") && node.at("matrix").size() == 16) {
        for (int i = 0; i < 16; i++) lm[i] = (float)node.at("matrix").at((size_t)i).number(0.0);
        return o;
    }
    float t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
    if (node.has("translation")) for (int i = 0; i < 3; i++) t[i] = (float)node.at("translation").at((size_t)i).number(0.0);
    if (node.has("rotation")) for (int i = 0; i < 4; i++) q[i] = (float)node.at("rotation").at((size_t)i).number(i == 3 ? 1.0 : 0.0);
    if (node.has("scale")) for (int i = 0; i < 3; i++) s[i] = (float)node.at("scale").at((size_t)i).number(1.0);
    const float qx = q[0], qy = q[1], qz = q[2], qw = q[3];
    const float sx = s[0], sy = s[1], sz = s[2];
    lm[0] = (1 - 2 * qy * qy - 2 * qz * qz) * sx;
    lm[1] = (2 * qx * qy + 2 * qz * qw) * sx;
    lm[2] = (2 * qx * qz - 2 * qy * qw) * sx;
    lm[3] = 0.f;
    lm[4] = (2 * qx * qy - 2 * qz * qw) * sy;
    lm[5] = (1 - 2 * qx * qx - 2 * qz * qz) * sy;
    lm[6] = (2 * qy * qz + 2 * qx * qw) * sy;
    lm[7] = 0.f;
    lm[8] = (2 * qx * qz + 2 * qy * qw) * sz;
    lm[9] = (2 * qy * qz - 2 * qx * qw) * sz;
    lm[10] = (1 - 2 * qx * qx - 2 * qy * qy) * sz;
    lm[11] = 0.f;
    lm[12] = t[0]; lm[13] = t[1]; lm[14] = t[2]; lm[15] = 1.f;
    return o;
}

std::string percent_decode(const std::string& s) { // net.percent_decode (input.odin:55)
    std::string o;
    for (size_t i = 0; i < s.size(); i++) {
        if (s[i] == '%' && i + 2 < s.size() + 0 && std::isxdigit((unsigned char)s[i + 1]) && std::isxdigit((unsigned char)s[i + 2])) {
            o.push_back((char)std::strtol(s.substr(i + 1, 2).c_str(), nullptr, 16));
            i += 2;
        } else {
            o.push_back(s[i]);
        }
    }
    return o;
}

bool base64_decode(const std::string& in, size_t from, std::vector<uint8_t>* out) {
    uint32_t acc = 0;
    int bits = 0;
    for (size_t i = from; i < in.size(); i++) {
        const char c = in[i];
        int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A';
        else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
        else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+' || c == '-') v = 62;
        else if (c == '/' || c == '_') v = 63;
        else if (c == '=') break;
        else if (c == '\n' || c == '\r') continue;
        else return false;
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) { bits -= 8; out->push_back((uint8_t)((acc >> bits) & 0xff)); }
    }
    return true;
}

std::string dir_of(const std::string& path) {
    const size_t k = path.find_last_of('/');
    return k == std::string::npos ? std::string(".") : (k == 0 ? std::string("/") : path.substr(0, k));
}
std::string join(const std::string& root, const std::string& rel) {
    if (!rel.empty() && rel[0] == '/') return rel;
    return root + "/" + rel;
}

struct Accessor {
    const uint8_t* base = nullptr;
    size_t count = 0, stride = 0;
    int component = 0, ncomp = 0;
    bool normalized = false;
    size_t comp_size = 0;
};

struct Loader {
    Json j;
    std::string root;
    std::vector<std::vector<uint8_t>> buffers;
    HostScene* scene = nullptr;
    std::map<std::string, int> texture_cache;
    std::string err;

    bool fail(const std::string& m) { err = m; return false; }

    bool accessor(int64_t idx, Accessor* a) {
        const Json& acc = j.at("accessors").at((size_t)idx);
        if (idx < 0 || acc.is_null()) return fail("accessor index out of range");
        if (!acc.has("bufferView")) return fail("accessor without a bufferView (sparse accessors are not supported)");
        const Json& bv = j.at("bufferViews").at((size_t)acc.at("bufferView").integer(-1));
        if (bv.is_null()) return fail("bufferView index out of range");
        const int64_t bi = bv.at("buffer").integer(-1);
        if (bi < 0 || (size_t)bi >= buffers.size()) return fail("buffer index out of range");
        a->component = (int)acc.at("componentType").integer(0);
        switch (a->component) {
        case 5120: case 5121: a->comp_size = 1; break;
        case 5122: case 5123: a->comp_size = 2; break;
        case 5125: case 5126: a->comp_size = 4; break;
        default: return fail("unsupported accessor component type");
        }
        const std::string& ty = acc.at("type").str;
        a->ncomp = ty == "SCALAR" ? 1 : ty == "VEC2" ? 2 : ty == "VEC3" ? 3 : ty == "VEC4" ? 4 : ty == "MAT4" ? 16 : 0;
        if (!a->ncomp) return fail("unsupported accessor type");
        a->count = (size_t)acc.at("count").integer(0);
        a->normalized = acc.at("normalized").kind == Json::Bool && acc.at("normalized").b;
        const size_t elem = a->comp_size * (size_t)a->ncomp;
        const size_t bstride = (size_t)bv.at("byteStride").integer(0);
        a->stride = bstride ? bstride : elem;
        const size_t off = (size_t)bv.at("byteOffset").integer(0) + (size_t)acc.at("byteOffset").integer(0);
        const auto& buf = buffers[(size_t)bi];
        if (a->count && off + a->stride * (a->count - 1) + elem > buf.size()) return fail("accessor reads past the end of its buffer");
        a->base = buf.data() + off;
        return true;
    }
    // cgltf_accessor_read_float: float passthrough, normalised integers scaled, others cast
    static void read_float(const Accessor& a, size_t index, float* out, int n) {
        const uint8_t* p = a.base + a.stride * index;
        for (int c = 0; c < n; c++) {
            if (c >= a.ncomp) { out[c] = 0.0f; continue; }
            const uint8_t* e = p + a.comp_size * (size_t)c;
            float v;
            switch (a.component) {
            case 5126: std::memcpy(&v, e, 4); break;
            case 5120: { int8_t x; std::memcpy(&x, e, 1); v = a.normalized ? std::max((float)x / 127.0f, -1.0f) : (float)x; break; }
            case 5121: { uint8_t x = *e; v = a.normalized ? (float)x / 255.0f : (float)x; break; }
            case 5122: { int16_t x; std::memcpy(&x, e, 2); v = a.normalized ? std::max((float)x / 32767.0f, -1.0f) : (float)x; break; }
            case 5123: { uint16_t x; std::memcpy(&x, e, 2); v = a.normalized ? (float)x / 65535.0f : (float)x; break; }
            default: { uint32_t x; std::memcpy(&x, e, 4); v = (float)x; break; }
            }
            out[c] = v;
        }
    }
    static uint64_t read_index(const Accessor& a, size_t index) { // cgltf_accessor_read_index
        const uint8_t* e = a.base + a.stride * index;
        switch (a.component) {
        case 5121: return *e;
        case 5123: { uint16_t x; std::memcpy(&x, e, 2); return x; }
        case 5125: { uint32_t x; std::memcpy(&x, e, 4); return x; }
        case 5120: { int8_t x; std::memcpy(&x, e, 1); return (uint64_t)(int64_t)x; }
        case 5122: { int16_t x; std::memcpy(&x, e, 2); return (uint64_t)(int64_t)x; }
        default: { float x; std::memcpy(&x, e, 4); return (uint64_t)x; }
        }
    }

    // load_sampler / load_image (input.odin:50-90): -1 == nil sampler
    bool load_sampler(const Json& view, int32_t* out) {
        *out = -1;
        if (view.is_null()) return true;
        const Json& tex = j.at("textures").at((size_t)view.at("index").integer(-1));
        if (tex.is_null()) return fail("texture index out of range");
        const Json& img = j.at("images").at((size_t)tex.at("source").integer(-1));
        if (img.is_null() || img.at("uri").kind != Json::String) return fail("Failed to decode image path");
        const std::string path = join(root, percent_decode(img.at("uri").str));
        auto it = texture_cache.find(path);
        if (it != texture_cache.end()) { *out = it->second; return true; }
        Texture t;
        std::string e;
        if (!load_texture(path, &t, &e)) return fail(e);
        *out = (int32_t)scene->textures.size();
        texture_cache[path] = *out;
        scene->textures.push_back(std::move(t));
        return true;
    }

    // triangles the hierarchy under `node_idx` will emit (one pre-pass, so the triangle array is allocated once)
    size_t count_triangles(int64_t node_idx, int depth) const {
        if (depth > 512) return 0;
        const Json& node = j.at("nodes").at((size_t)node_idx);
        if (node_idx < 0 || node.is_null()) return 0;
        size_t n = 0;
        if (node.has("mesh")) {
            const Json& prims = j.at("meshes").at((size_t)node.at("mesh").integer(-1)).at("primitives");
            for (size_t pi = 0; pi < prims.size(); pi++) {
                const Json& prim = prims.at(pi);
                const Json& acc = j.at("accessors").at((size_t)(prim.has("indices") ? prim.at("indices").integer(-1)
                                                                                    : prim.at("attributes").at("POSITION").integer(-1)));
                n += (size_t)std::max<int64_t>(acc.at("count").integer(0), 0) / 3;
            }
        }
        const Json& kids = node.at("children");
        for (size_t k = 0; k < kids.size(); k++) n += count_triangles(kids.at(k).integer(-1), depth + 1);
        return n;
    }

    bool populate(int64_t node_idx, const Mat4& parent, int depth) { // populate_scene input.odin:92-233
        if (depth > 512) return fail("node hierarchy too deep (cycle?)");
        const Json& node = j.at("nodes").at((size_t)node_idx);
        if (node_idx < 0 || node.is_null()) return fail("node index out of range");
        const Mat4 transform = mul(parent, node_transform_local(node));
        if (node.has("camera")) { // :103-109
            const Json& cam = j.at("cameras").at((size_t)node.at("camera").integer(-1));
            for (int r = 0; r < 3; r++) {
                scene->cam.pos[r] = transform.m[3][r];
                scene->cam.basis[0 + r] = transform.m[0][r];
                scene->cam.basis[3 + r] = transform.m[1][r];
                scene->cam.basis[6 + r] = -transform.m[2][r];
            }
            scene->cam.fov_x = (float)cam.at("perspective").at("yfov").number(0.0);
        }
        if (node.has("mesh")) {
            const Json& mesh = j.at("meshes").at((size_t)node.at("mesh").integer(-1));
            if (mesh.is_null()) return fail("mesh index out of range");
            const Json& prims = mesh.at("primitives");
            for (size_t pi = 0; pi < prims.size(); pi++) {
                const Json& prim = prims.at(pi);
                const Json& attrs = prim.at("attributes");
                if (!attrs.has("POSITION")) return fail("No position accessor found in mesh primitive");
                if (!prim.has("material")) return fail("mesh primitive without a material (the reference dereferences nil, input.odin:138)");
                const Json& gm = j.at("materials").at((size_t)prim.at("material").integer(-1));
                if (gm.is_null()) return fail("material index out of range");
                const Json& pbr = gm.at("pbrMetallicRoughness");
                ort_material mat{};
                for (int c = 0; c < 3; c++) {
                    mat.color_factor[c] = pbr.has("baseColorFactor") ? (float)pbr.at("baseColorFactor").at((size_t)c).number(1.0) : 1.0f;
                    mat.emission_factor[c] = gm.has("emissiveFactor") ? (float)gm.at("emissiveFactor").at((size_t)c).number(0.0) : 0.0f;
                }
                if (!load_sampler(pbr.at("baseColorTexture"), &mat.color_texture)) return false;
                if (!load_sampler(gm.at("emissiveTexture"), &mat.emission_texture)) return false;
                mat.roughness_factor = (float)pbr.at("roughnessFactor").number(1.0);
                mat.metallic_factor = (float)pbr.at("metallicFactor").number(1.0);
                if (!load_sampler(pbr.at("metallicRoughnessTexture"), &mat.metallic_roughness_texture)) return false;
                if (!load_sampler(gm.at("normalTexture"), &mat.normal_texture)) return false;
                const Json& es = gm.at("extensions").at("KHR_materials_emissive_strength");
                if (!es.is_null()) { // :157-159
                    const float k = (float)es.at("emissiveStrength").number(1.0);
                    for (int c = 0; c < 3; c++) mat.emission_factor[c] *= k;
                }
                const int64_t material_index = (int64_t)scene->materials.size();
                scene->materials.push_back(mat);

                Accessor pos, nrm, uv, tan, idx;
                if (!accessor(attrs.at("POSITION").integer(-1), &pos)) return false;
                const bool has_n = attrs.has("NORMAL"), has_uv = attrs.has("TEXCOORD_0"), has_t = attrs.has("TANGENT");
                const bool has_i = prim.has("indices");
                if (has_n && !accessor(attrs.at("NORMAL").integer(-1), &nrm)) return false;
                if (has_uv && !accessor(attrs.at("TEXCOORD_0").integer(-1), &uv)) return false;
                if (has_t && !accessor(attrs.at("TANGENT").integer(-1), &tan)) return false;
                if (has_i && !accessor(prim.at("indices").integer(-1), &idx)) return false;
                // normal_transform = cofactor(mat3(transform)) (:203)
                float cof[3][3]; // cof[r][c]
                for (int r = 0; r < 3; r++)
                    for (int c = 0; c < 3; c++) {
                        const int r0 = r == 0 ? 1 : 0, r1 = r == 2 ? 1 : 2, c0 = c == 0 ? 1 : 0, c1 = c == 2 ? 1 : 2;
                        const float minor = transform.m[c0][r0] * transform.m[c1][r1] - transform.m[c1][r0] * transform.m[c0][r1];
                        cof[r][c] = ((r + c) % 2 == 0) ? minor : -minor;
                    }
                const size_t num_vertices = has_i ? idx.count : pos.count;
                const size_t n_new = num_vertices / 3, first_out = scene->triangles.size();
                if (scene->triangles.capacity() < first_out + n_new) // geometric growth: an exact reserve per primitive is quadratic
                    scene->triangles.reserve(std::max(first_out + n_new, 2 * scene->triangles.capacity()));
                scene->triangles.resize(first_out + n_new);
                ort_triangle* out_tris = scene->triangles.data() + first_out;
                std::atomic<int> bad{0}; // 1 position, 2 normal, 3 uv, 4 tangent index out of range
                auto work = [&](size_t t0, size_t t1) {
                    for (size_t i = t0; i < t1; i++) {
                        float P[3][3], N[3][3] = {}, UV[3][2] = {}, T[3][4] = {};
                        for (size_t k = 0; k < 3; k++) {
                            const size_t index = has_i ? (size_t)read_index(idx, i * 3 + k) : i * 3 + k;
                            if (index >= pos.count) { bad.store(1); return; }
                            float raw[4];
                            read_float(pos, index, raw, 3);
                            const float v4[4] = {raw[0], raw[1], raw[2], 1.0f};
                            mul_vec(transform, v4, P[k]);
                            if (has_n) {
                                if (index >= nrm.count) { bad.store(2); return; }
                                read_float(nrm, index, raw, 3);
                                for (int r = 0; r < 3; r++) N[k][r] = (cof[r][0] * raw[0] + cof[r][1] * raw[1]) + cof[r][2] * raw[2];
                                normalize3(N[k]);
                            }
                            if (has_uv) {
                                if (index >= uv.count) { bad.store(3); return; }
                                read_float(uv, index, UV[k], 2);
                            }
                            float t4[4] = {0, 0, 0, 0};
                            if (has_t) {
                                if (index >= tan.count) { bad.store(4); return; }
                                read_float(tan, index, t4, 4);
                            }
                            // tangents[i].xyz = normalize((transform * {t, 0}).xyz): a zero tangent becomes NaN (:193-195)
                            const float tv[4] = {t4[0], t4[1], t4[2], 0.0f};
                            mul_vec(transform, tv, T[k]);
                            normalize3(T[k]);
                            T[k][3] = t4[3];
                        }
                        ort_triangle tri{};
                        float e1[3], e2[3], ng[3];
                        for (int r = 0; r < 3; r++) { e1[r] = P[1][r] - P[0][r]; e2[r] = P[2][r] - P[0][r]; }
                        ng[0] = e1[1] * e2[2] - e1[2] * e2[1];
                        ng[1] = e1[2] * e2[0] - e1[0] * e2[2];
                        ng[2] = e1[0] * e2[1] - e1[1] * e2[0];
                        normalize3(ng);
                        for (int r = 0; r < 3; r++) {
                            tri.p[r] = P[0][r]; tri.u[r] = e1[r]; tri.v[r] = e2[r]; tri.ng[r] = ng[r];
                            tri.n1[r] = has_n ? N[0][r] : ng[r];
                            tri.n2[r] = has_n ? N[1][r] : ng[r];
                            tri.n3[r] = has_n ? N[2][r] : ng[r];
                        }
                        std::memcpy(tri.tex1, UV[0], 8); std::memcpy(tri.tex2, UV[1], 8); std::memcpy(tri.tex3, UV[2], 8);
                        std::memcpy(tri.tan1, T[0], 16); std::memcpy(tri.tan2, T[1], 16); std::memcpy(tri.tan3, T[2], 16);
                        tri.material_index = material_index;
                        out_tris[i] = tri;
                    }
                };
                // big primitives are flattened on several host threads (disjoint output ranges)
                const size_t nthr = n_new >= 65536 ? std::min<size_t>(16, std::max(1u, std::thread::hardware_concurrency())) : 1;
                if (nthr <= 1) work(0, n_new);
                else {
                    std::vector<std::thread> pool;
                    const size_t per = (n_new + nthr - 1) / nthr;
                    for (size_t t = 0; t < nthr; t++) {
                        const size_t a = t * per, b = std::min(n_new, a + per);
                        if (a < b) pool.emplace_back(work, a, b);
                    }
                    for (auto& th : pool) th.join();
                }
                switch (bad.load()) {
                case 1: return fail("Failed to read position data from accessor");
                case 2: return fail("Failed to read normal data from accessor");
                case 3: return fail("Failed to read UV data from accessor");
                case 4: return fail("Failed to read tangent data from accessor");
                default: break;
                }
            }
        }
        const Json& kids = node.at("children");
        for (size_t k = 0; k < kids.size(); k++)
            if (!populate(kids.at(k).integer(-1), transform, depth + 1)) return false;
        return true;
    }
};

} // namespace

bool read_gltf(const std::string& path, HostScene* out, std::string* err) {
    std::vector<uint8_t> bytes;
    if (!read_file(path, &bytes)) { *err = "Failed to open input file: " + path; return false; }
    Loader L;
    std::string perr;
    if (!JsonParser::parse(std::string(bytes.begin(), bytes.end()), &L.j, &perr)) {
        *err = "Failed to parse .gltf file: " + perr; return false;
    }
    L.root = dir_of(path);
    L.scene = out;
    *out = HostScene();
    const Json& bufs = L.j.at("buffers");
    for (size_t i = 0; i < bufs.size(); i++) { // cgltf.load_buffers (input.odin:35-41)
        const Json& uri = bufs.at(i).at("uri");
        std::vector<uint8_t> b;
        if (uri.kind != Json::String) { *err = "Failed to load buffers from .gltf file: buffer without uri"; return false; }
        if (uri.str.compare(0, 5, "data:") == 0) {
            const size_t comma = uri.str.find(',');
            if (comma == std::string::npos || !base64_decode(uri.str, comma + 1, &b)) {
                *err = "Failed to load buffers from .gltf file: bad data URI"; return false;
            }
        } else if (!read_file(join(L.root, percent_decode(uri.str)), &b)) {
            *err = "Failed to load buffers from .gltf file: " + uri.str; return false;
        }
        L.buffers.push_back(std::move(b));
    }
    ort_material dummy{}; // Material{} (input.odin:44): nil samplers
    dummy.color_texture = dummy.emission_texture = dummy.metallic_roughness_texture = dummy.normal_texture = -1;
    out->materials.push_back(dummy);
    for (int i = 0; i < 9; i++) out->cam.basis[i] = (i % 4 == 0) ? 1.0f : 0.0f;

    const Mat4 ident = identity();
    const Json* roots = nullptr;
    if (L.j.has("scene")) roots = &L.j.at("scenes").at((size_t)L.j.at("scene").integer(0)).at("nodes"); // input.odin:236-248
    else if (L.j.at("scenes").size() > 0) roots = &L.j.at("scenes").at((size_t)0).at("nodes");
    {
        size_t total = 0;
        if (roots) for (size_t i = 0; i < roots->size(); i++) total += L.count_triangles(roots->at(i).integer(-1), 0);
        else for (size_t i = 0; i < L.j.at("nodes").size(); i++) total += L.count_triangles((int64_t)i, 0);
        if (total < ((size_t)1 << 31)) out->triangles.reserve(total);
    }
    bool ok = true;
    if (roots) {
        for (size_t i = 0; ok && i < roots->size(); i++) ok = L.populate(roots->at(i).integer(-1), ident, 0);
    } else {
        for (size_t i = 0; ok && i < L.j.at("nodes").size(); i++) ok = L.populate((int64_t)i, ident, 0);
    }
    if (!ok) { *err = L.err; return false; }
    return true;
}

bool finish_scene(HostScene* s, int bvh_device, std::string* err) {
    // light_surfaces: norm_l1(emission_factor) > 1e-6, collected in glTF order BEFORE the scene
    // build reorders scene.trigs (raytracer.odin:63-66)
    s->light_triangles.clear();
    for (const ort_triangle& t : s->triangles) {
        if (t.material_index < 0 || (size_t)t.material_index >= s->materials.size()) { *err = "triangle material_index out of range"; return false; }
        const float* e = s->materials[(size_t)t.material_index].emission_factor;
        const float l1 = (std::fabs(e[0]) + std::fabs(e[1])) + std::fabs(e[2]);
        if (l1 > 1e-6f) s->light_triangles.push_back(t);
    }
    auto build = [&](std::vector<ort_triangle>& tris, std::vector<ort_bvh_node>* nodes) {
        const int64_t n = (int64_t)tris.size(), cap = std::max<int64_t>(2 * n, 1);
        nodes->assign((size_t)cap, ort_bvh_node{});
        const int64_t cnt = bvh_device >= 0 ? ort_bvh_build_device(bvh_device, n ? tris.data() : nullptr, n, nodes->data(), cap)
                                            : ort_bvh_build(n ? tris.data() : nullptr, n, nodes->data(), cap);
        if (cnt < 0) {
            *err = bvh_device >= 0 ? std::string("ort_bvh_build_device: ") + ort_bvh_build_device_error() : "ort_bvh_build failed";
            return false;
        }
        nodes->resize((size_t)cnt);
        return true;
    };
    if (!build(s->triangles, &s->bvh)) return false;          // raytracer.odin:72
    if (!build(s->light_triangles, &s->light_bvh)) return false; // raytracer.odin:75
    s->finished = true;
    return true;
}

void scene_view(const HostScene& s, ort_scene* out, std::vector<ort_texture>* tex_keep) {
    auto view = [](const Texture& t) {
        ort_texture o{};
        o.data = t.data();
        o.width = t.width; o.height = t.height; o.channels = t.channels;
        o.is_f32 = t.is_f32 ? 1 : 0;
        o.stride = (int64_t)t.width * t.channels; // textures.odin:65
        return o;
    };
    tex_keep->clear();
    tex_keep->reserve(s.textures.size() + 1);
    for (const Texture& t : s.textures) tex_keep->push_back(view(t));
    *out = ort_scene{};
    out->cam = s.cam;
    out->triangles = s.triangles.data(); out->n_triangles = (int64_t)s.triangles.size();
    out->bvh = s.bvh.data(); out->n_bvh_nodes = (int64_t)s.bvh.size();
    out->light_triangles = s.light_triangles.data(); out->n_light_triangles = (int64_t)s.light_triangles.size();
    out->light_bvh = s.light_bvh.data(); out->n_light_bvh_nodes = (int64_t)s.light_bvh.size();
    out->materials = s.materials.data(); out->n_materials = (int64_t)s.materials.size();
    out->textures = tex_keep->data(); out->n_textures = (int64_t)s.textures.size();
    if (s.has_env) {
        tex_keep->push_back(view(s.env_map));
        out->textures = tex_keep->data(); // no reallocation: capacity was reserved
        out->env_map = &tex_keep->back();
    }
}

// ------------------------------------------------------------------------------------------------
// output.odin
// ------------------------------------------------------------------------------------------------
void get_rgb_image(const ort_sample_stats* pixels, int w, int h, std::vector<uint8_t>* rgb) {
    rgb->assign((size_t)w * h * 3, 0);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        const ort_sample_stats& s = pixels[i];
        for (int c = 0; c < 3; c++) {
            float raw = s.total[c] / (float)s.count;         // mode Mean (output.odin:38)
            raw = raw > 0.0f ? raw : 0.0f;                    // linalg.max(raw, 0); NaN -> 0
            float tm = (raw * (2.51f * raw + 0.03f)) / (raw * (2.43f * raw + 0.59f) + 0.14f); // tone_mapping_aces :21-28
            tm = tm < 0.0f ? 0.0f : (tm > 1.0f ? 1.0f : tm);
            const float g = std::pow(tm, (float)(1.0 / 2.2));
            const float r = std::round(g * 255.0f);
            (*rgb)[i * 3 + (size_t)c] = (r == r) ? (uint8_t)r : 0;
        }
    }
}

bool save_result(const ort_sample_stats* pixels, int w, int h, const std::string& path, std::string* err) {
    std::vector<uint8_t> rgb;
    get_rgb_image(pixels, w, h, &rgb);
    auto ends_with = [&](const char* suf) {
        const size_t n = std::strlen(suf);
        return path.size() >= n && path.compare(path.size() - n, n, suf) == 0;
    };
    if (ends_with(".ppm")) return write_ppm(path, w, h, rgb.data(), err);
    if (ends_with(".png")) return write_png(path, w, h, rgb.data(), err);
    *err = "Unsupported file format: " + path; // output.odin:105
    return false;
}

} // namespace orh
