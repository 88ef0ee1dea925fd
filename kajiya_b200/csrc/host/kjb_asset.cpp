// glTF 2.0 scene -> TriangleMesh (include/kjb_asset.h).
//
// Mirrors kajiya-asset: import_gltf.rs:32-160 (buffer / image source resolution: data: URIs, file: URIs, relative paths, GLB BIN chunk,
// buffer views), mesh.rs:100-112 (node tree walk), :120-262 (load_gltf_material), :278-441 (LoadGltfScene::run).
// The reference reads the document through the `gltf` crate (git b9c04be) and does its vector maths with glam 0.22; neither is vendored
// in /root/reference, so their arithmetic is restated here from the published sources in the same operation order (f32, column-major,
// left-to-right sums, no fused multiply-add): node TRS -> matrix as gltf's Transform::matrix() (T * R * S), parent * child as glam's
// Mat4 * Mat4, point/direction transforms as glam's Mat4 * Vec4, normalize as v * (1 / sqrt(dot)).  "Parity unpinned" for the last ulp
// of those floats (no Rust toolchain here to run the reference); indices, material ids, maps and counts are exact by construction.
#include "../../../include/kjb_asset.h"
#include "kjb_asset_json.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace kjb_asset_detail {
static thread_local std::string g_error;
void set_error(const std::string& m) { g_error = m; }
bool decode_image(const uint8_t* bytes, size_t n, std::vector<uint8_t>& rgba, uint32_t& w, uint32_t& h);
bool decode_dds(const uint8_t* p, size_t n, std::vector<uint8_t>& texels, uint32_t& W, uint32_t& H, uint32_t& mips, uint32_t& srgb);
void build_mips(const uint8_t* rgba8, uint32_t width, uint32_t height, bool use_mips, const uint32_t* swizzle, std::vector<uint8_t>& out, uint32_t& ow, uint32_t& oh, uint32_t& levels);
}  // namespace kjb_asset_detail
using namespace kjb_asset_detail;

namespace {

// ------------------------------------------------------------------------------------------------ small f32 algebra (glam / gltf::math order)
struct M4 { float c[4][4]; };   // c[col][row]
M4 m4_identity() { M4 m; memset(&m, 0, sizeof m); m.c[0][0] = m.c[1][1] = m.c[2][2] = m.c[3][3] = 1.0f; return m; }
// glam Mat4::mul_mat4: every result column is self * rhs.col, with mul_vec4 = ((x_axis*v.x + y_axis*v.y) + z_axis*v.z) + w_axis*v.w
void m4_mul_vec4(const M4& m, const float v[4], float out[4]) {
    for (int r = 0; r < 4; ++r) { float s = m.c[0][r] * v[0]; s = s + m.c[1][r] * v[1]; s = s + m.c[2][r] * v[2]; s = s + m.c[3][r] * v[3]; out[r] = s; }
}
M4 m4_mul(const M4& a, const M4& b) { M4 o; for (int j = 0; j < 4; ++j) m4_mul_vec4(a, b.c[j], o.c[j]); return o; }
// glam Mat4::from_scale_rotation_translation / quat_to_axes
M4 m4_from_srt(const float s[3], const float q[4], const float t[3]) {
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float x2 = x + x, y2 = y + y, z2 = z + z;
    const float xx = x * x2, xy = x * y2, xz = x * z2, yy = y * y2, yz = y * z2, zz = z * z2, wx = w * x2, wy = w * y2, wz = w * z2;
    M4 m = m4_identity();
    const float ax[3][3] = {{1.0f - (yy + zz), xy + wz, xz - wy}, {xy - wz, 1.0f - (xx + zz), yz + wx}, {xz + wy, yz - wx, 1.0f - (xx + yy)}};
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) m.c[c][r] = ax[c][r] * s[c];
    m.c[3][0] = t[0]; m.c[3][1] = t[1]; m.c[3][2] = t[2];
    return m;
}
// gltf::scene::Transform::Decomposed -> matrix(): T * R * S with gltf::math's (cgmath-derived) Matrix4 product, each entry a left-to-right dot
M4 gltf_mul(const M4& a, const M4& b) {
    M4 o;
    for (int j = 0; j < 4; ++j) for (int r = 0; r < 4; ++r) { float s = a.c[0][r] * b.c[j][0]; s = s + a.c[1][r] * b.c[j][1]; s = s + a.c[2][r] * b.c[j][2]; s = s + a.c[3][r] * b.c[j][3]; o.c[j][r] = s; }
    return o;
}
M4 gltf_trs(const float t[3], const float q[4], const float s[3]) {
    M4 T = m4_identity(); T.c[3][0] = t[0]; T.c[3][1] = t[1]; T.c[3][2] = t[2];
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float x2 = x + x, y2 = y + y, z2 = z + z;
    const float xx2 = x2 * x, xy2 = x2 * y, xz2 = x2 * z, yy2 = y2 * y, yz2 = y2 * z, zz2 = z2 * z, sy2 = y2 * w, sz2 = z2 * w, sx2 = x2 * w;
    M4 R = m4_identity();
    R.c[0][0] = 1.0f - yy2 - zz2; R.c[0][1] = xy2 + sz2; R.c[0][2] = xz2 - sy2;
    R.c[1][0] = xy2 - sz2; R.c[1][1] = 1.0f - xx2 - zz2; R.c[1][2] = yz2 + sx2;
    R.c[2][0] = xz2 + sy2; R.c[2][1] = yz2 - sx2; R.c[2][2] = 1.0f - xx2 - yy2;
    M4 S = m4_identity(); S.c[0][0] = s[0]; S.c[1][1] = s[1]; S.c[2][2] = s[2];
    return gltf_mul(gltf_mul(T, R), S);
}
double m4_det(const M4& m) {   // only the sign is used (winding flip)
    double a[4][4]; for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) a[r][c] = m.c[c][r];
    double det = 1.0;
    for (int i = 0; i < 4; ++i) {
        int piv = i; for (int r = i + 1; r < 4; ++r) if (std::fabs(a[r][i]) > std::fabs(a[piv][i])) piv = r;
        if (a[piv][i] == 0.0) return 0.0;
        if (piv != i) { for (int c = 0; c < 4; ++c) std::swap(a[piv][c], a[i][c]); det = -det; }
        det *= a[i][i];
        for (int r = i + 1; r < 4; ++r) { const double f = a[r][i] / a[i][i]; for (int c = i; c < 4; ++c) a[r][c] -= f * a[i][c]; }
    }
    return det;
}
void normalize3(float v[3]) {   // glam Vec3::normalize: self * self.length_recip(), length_recip = 1 / sqrt(x*x + y*y + z*z)
    const float d = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const float inv = 1.0f / std::sqrt(d);
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
}

// ------------------------------------------------------------------------------------------------ URI / file helpers (import_gltf.rs:32-86)
bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { set_error("cannot open " + path); return false; }
    fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); set_error("cannot size " + path); return false; }
    out.resize(size_t(n));
    const size_t got = n ? fread(out.data(), 1, size_t(n), f) : 0;
    fclose(f);
    if (got != size_t(n)) { set_error("short read on " + path); return false; }
    return true;
}
bool base64_decode(const std::string& s, std::vector<uint8_t>& out) {
    uint32_t acc = 0; int bits = 0;
    for (char ch : s) {
        int v;
        if (ch >= 'A' && ch <= 'Z') v = ch - 'A'; else if (ch >= 'a' && ch <= 'z') v = ch - 'a' + 26; else if (ch >= '0' && ch <= '9') v = ch - '0' + 52;
        else if (ch == '+' || ch == '-') v = 62; else if (ch == '/' || ch == '_') v = 63; else if (ch == '=') break; else if (ch == '\n' || ch == '\r') continue; else return false;
        acc = (acc << 6) | uint32_t(v); bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back(uint8_t(acc >> bits)); acc &= (1u << bits) - 1; }
    }
    return true;
}
std::string percent_decode(const std::string& s) {   // urlencoding::decode (images only, import_gltf.rs:128)
    std::string o;
    auto hex = [](char c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1; };
    for (size_t i = 0; i < s.size(); ++i) {
        if (s[i] == '%' && i + 2 < s.size() && hex(s[i + 1]) >= 0 && hex(s[i + 2]) >= 0) { o += char(hex(s[i + 1]) * 16 + hex(s[i + 2])); i += 2; }
        else o += s[i];
    }
    return o;
}
std::string dirname_of(const std::string& path) { const size_t k = path.find_last_of("/\\"); return k == std::string::npos ? std::string("./") : path.substr(0, k + 1); }
// Scheme::read
bool read_uri(const std::string& base, const std::string& uri, std::vector<uint8_t>& out) {
    if (uri.find(':') != std::string::npos) {
        if (uri.compare(0, 5, "data:") == 0) {
            const size_t k = uri.find(";base64,");
            const std::string payload = k == std::string::npos ? uri.substr(5) : uri.substr(k + 8);
            if (!base64_decode(payload, out)) { set_error("bad base64 payload in data: URI"); return false; }
            return true;
        }
        if (uri.compare(0, 7, "file://") == 0) return read_file(uri.substr(7), out);
        if (uri.compare(0, 5, "file:") == 0) return read_file(uri.substr(5), out);
        set_error("unsupported URI scheme: " + uri.substr(0, uri.find(':') + 1));
        return false;
    }
    return read_file(base + uri, out);
}

// ------------------------------------------------------------------------------------------------ document + accessors
struct Doc {
    Json root; std::string base;
    std::vector<std::vector<uint8_t>> buffers;
    const Json& list(const char* k) const { static const Json empty; const Json* j = root.get(k); return j && j->is(Json::Array) ? *j : empty; }
};
int component_bytes(int64_t ct) { return ct == 5120 || ct == 5121 ? 1 : ct == 5122 || ct == 5123 ? 2 : ct == 5125 || ct == 5126 ? 4 : 0; }
int type_components(const std::string& t) { return t == "SCALAR" ? 1 : t == "VEC2" ? 2 : t == "VEC3" ? 3 : t == "VEC4" ? 4 : t == "MAT2" ? 4 : t == "MAT3" ? 9 : t == "MAT4" ? 16 : 0; }

struct View { const uint8_t* p = nullptr; size_t stride = 0, count = 0; int64_t ctype = 0; int comps = 0; bool normalized = false; };
bool view_of(const Doc& d, const Json& bv_list, int64_t bv_index, size_t byte_offset, size_t elem_bytes, size_t count, View& v, const char* what) {
    if (bv_index < 0) { set_error(std::string(what) + ": bad bufferView index"); return false; }
    const Json& bv = bv_list[size_t(bv_index)];
    if (!bv.is(Json::Object)) { set_error(std::string(what) + ": bad bufferView index"); return false; }
    const int64_t bi = bv.i("buffer", -1);
    if (bi < 0 || size_t(bi) >= d.buffers.size()) { set_error(std::string(what) + ": bad buffer index"); return false; }
    bool ok = true;
    const size_t bv_off = bv.size("byteOffset", 0, ok), len = bv.size("byteLength", 0, ok);
    const size_t stride = bv.has("byteStride") ? bv.size("byteStride", 0, ok) : elem_bytes;
    if (!ok) { set_error(std::string(what) + ": bufferView byteOffset / byteLength / byteStride is not a non-negative integer"); return false; }
    const size_t buf_size = d.buffers[size_t(bi)].size();
    // every comparison is written so that nothing can wrap: the view lies inside its buffer, the accessor inside its view
    if (bv_off > buf_size || len > buf_size - bv_off) { set_error(std::string(what) + ": buffer view reaches past its buffer"); return false; }
    if (count) {
        if (elem_bytes > len || byte_offset > len - elem_bytes) { set_error(std::string(what) + ": accessor reaches past its buffer view"); return false; }
        const size_t room = len - byte_offset - elem_bytes;   // bytes available for the (count - 1) strides
        if (count - 1 != 0 && (stride == 0 ? false : (count - 1) > room / stride)) { set_error(std::string(what) + ": accessor reaches past its buffer view"); return false; }
    }
    v.p = d.buffers[size_t(bi)].data() + bv_off + byte_offset; v.stride = stride;
    return true;
}
double load_component(const uint8_t* p, int64_t ct) {
    switch (ct) {
        case 5120: return double(*reinterpret_cast<const int8_t*>(p));
        case 5121: return double(*p);
        case 5122: { int16_t v; memcpy(&v, p, 2); return double(v); }
        case 5123: { uint16_t v; memcpy(&v, p, 2); return double(v); }
        case 5125: { uint32_t v; memcpy(&v, p, 4); return double(v); }
        default: { float v; memcpy(&v, p, 4); return double(v); }
    }
}
// Reads an accessor into rows of `comps` raw component values kept in their storage type's value (no normalisation), sparse substitution applied.
struct AccessorData { std::vector<float> f; std::vector<uint32_t> u; int64_t ctype = 0; int comps = 0; size_t count = 0; };
bool read_accessor(const Doc& d, int64_t index, bool as_uint, AccessorData& out, const char* what) {
    const Json& acc = d.list("accessors")[size_t(index)];
    if (index < 0 || !acc.is(Json::Object)) { set_error(std::string(what) + ": bad accessor index"); return false; }
    bool fields_ok = true;
    out.ctype = acc.i("componentType", 0); out.comps = type_components(acc.s("type")); out.count = acc.size("count", 0, fields_ok);
    const size_t acc_off = acc.size("byteOffset", 0, fields_ok);
    const int cb = component_bytes(out.ctype);
    if (!cb || !out.comps) { set_error(std::string(what) + ": bad accessor componentType/type"); return false; }
    if (!fields_ok || out.count > (size_t(1) << 31)) { set_error(std::string(what) + ": accessor count / byteOffset is not a plausible non-negative integer"); return false; }
    const size_t n = out.count * size_t(out.comps);   // <= 2^31 * 16: cannot wrap
    const Json& views = d.list("bufferViews");
    View v;   // validate the extent BEFORE sizing the output: a hostile count must not drive a multi-gigabyte allocation
    if (acc.has("bufferView")) { if (!view_of(d, views, acc.i("bufferView", -1), acc_off, size_t(cb) * out.comps, out.count, v, what)) return false; }
    else if (out.count > (size_t(1) << 28)) { set_error(std::string(what) + ": accessor without a buffer view is implausibly large"); return false; }
    if (as_uint) out.u.assign(n, 0); else out.f.assign(n, 0.0f);
    auto store = [&](size_t row, const uint8_t* p) {
        for (int c = 0; c < out.comps; ++c) {
            const uint8_t* q = p + size_t(c) * cb;
            if (as_uint) { uint32_t v = 0; if (cb == 1) v = *q; else if (cb == 2) { uint16_t t; memcpy(&t, q, 2); v = t; } else memcpy(&v, q, 4); out.u[row * out.comps + c] = v; }
            else out.f[row * out.comps + c] = float(load_component(q, out.ctype));
        }
    };
    if (acc.has("bufferView")) for (size_t i = 0; i < out.count; ++i) store(i, v.p + i * v.stride);
    if (const Json* sp = acc.get("sparse")) {
        bool sp_ok = true;
        const size_t sc = sp->size("count", 0, sp_ok);
        const Json* ji = sp->get("indices"); const Json* jv = sp->get("values");
        if (!ji || !jv || !sp_ok || sc > out.count) { set_error(std::string(what) + ": malformed sparse accessor"); return false; }
        const size_t ji_off = ji->size("byteOffset", 0, sp_ok), jv_off = jv->size("byteOffset", 0, sp_ok);
        if (!sp_ok) { set_error(std::string(what) + ": malformed sparse accessor"); return false; }
        const int64_t ict = ji->i("componentType", 0); const int icb = component_bytes(ict);
        if (!icb) { set_error(std::string(what) + ": bad sparse index type"); return false; }
        View vi, vv;
        if (!view_of(d, views, ji->i("bufferView", -1), ji_off, size_t(icb), sc, vi, what)) return false;
        if (!view_of(d, views, jv->i("bufferView", -1), jv_off, size_t(cb) * out.comps, sc, vv, what)) return false;
        for (size_t k = 0; k < sc; ++k) {
            const size_t row = size_t(load_component(vi.p + k * vi.stride, ict));
            if (row >= out.count) { set_error(std::string(what) + ": sparse index out of range"); return false; }
            store(row, vv.p + k * vv.stride);
        }
    }
    return true;
}
// gltf::mesh::util casting iterators: normalised integers -> f32
float norm_to_f32(float raw, int64_t ct) {
    switch (ct) {
        case 5121: return raw / 255.0f;
        case 5123: return raw / 65535.0f;
        case 5120: return std::max(raw / 127.0f, -1.0f);
        case 5122: return std::max(raw / 32767.0f, -1.0f);
        default: return raw;
    }
}

// ------------------------------------------------------------------------------------------------ the asset
struct Texture { std::vector<uint8_t> texels; uint32_t w = 1, h = 1, mips = 1, srgb = 0; };
}  // namespace

struct kjb_asset {
    std::vector<float> positions, normals, uvs, colors, tangents;
    std::vector<uint32_t> material_ids, indices;
    std::vector<kjb_mesh_material> materials;
    std::vector<Texture> maps;
    std::vector<kjb_texture_desc> map_descs;
    uint32_t stats[4] = {0, 0, 0, 0};
};

namespace {
struct ImageSource { bool loaded = false, ok = false, is_dds = false; std::vector<uint8_t> rgba; uint32_t w = 0, h = 0, dds_mips = 0, dds_srgb = 0; std::string where; };   // DDS: rgba holds the whole decoded chain

struct Importer {
    Doc d; kjb_asset* out = nullptr;
    std::vector<ImageSource> images;

    bool load_document(const std::string& path) {
        std::vector<uint8_t> file;
        if (!read_file(path, file)) return false;
        d.base = dirname_of(path);
        std::vector<uint8_t> blob; bool have_blob = false;
        std::string err;
        if (file.size() >= 12 && !memcmp(file.data(), "glTF", 4)) {   // GLB container: header, JSON chunk, optional BIN chunk
            uint32_t version, total; memcpy(&version, file.data() + 4, 4); memcpy(&total, file.data() + 8, 4);
            if (version != 2 || total > file.size()) { set_error("glb: unsupported version or truncated file"); return false; }
            size_t off = 12; bool have_json = false;
            while (off + 8 <= total) {
                uint32_t clen, ctype; memcpy(&clen, file.data() + off, 4); memcpy(&ctype, file.data() + off + 4, 4);
                if (off + 8 + size_t(clen) > total) { set_error("glb: chunk reaches past the end of the file"); return false; }
                const uint8_t* cp = file.data() + off + 8;
                if (ctype == 0x4E4F534Au && !have_json) { JsonParser jp(reinterpret_cast<const char*>(cp), clen); if (!jp.parse(d.root, err)) { set_error("glb: " + err); return false; } have_json = true; }
                else if (ctype == 0x004E4942u && !have_blob) { blob.assign(cp, cp + clen); have_blob = true; }
                off += 8 + size_t(clen);
            }
            if (!have_json) { set_error("glb: no JSON chunk"); return false; }
        } else {
            JsonParser jp(reinterpret_cast<const char*>(file.data()), file.size());
            if (!jp.parse(d.root, err)) { set_error("gltf: " + err); return false; }
        }
        if (!d.root.is(Json::Object)) { set_error("gltf: top level is not an object"); return false; }
        // import_buffer_data (import_gltf.rs:89-116)
        const Json& bufs = d.list("buffers");
        for (size_t i = 0; i < bufs.size(); ++i) {
            std::vector<uint8_t> data;
            if (bufs[i].has("uri")) { if (!read_uri(d.base, bufs[i].s("uri"), data)) return false; }
            else { if (!have_blob) { set_error("gltf: buffer without uri but no GLB BIN chunk"); return false; } data.swap(blob); have_blob = false; }
            bool bl_ok = true;
            const size_t want = bufs[i].size("byteLength", 0, bl_ok);
            if (!bl_ok) { set_error("gltf: buffer byteLength is not a non-negative integer"); return false; }
            if (data.size() < want) { char m[160]; snprintf(m, sizeof m, "gltf: buffer %zu is %zu bytes, document says %zu", i, data.size(), want); set_error(m); return false; }
            while (data.size() % 4) data.push_back(0);
            d.buffers.push_back(std::move(data));
        }
        images.resize(d.list("images").size());
        return true;
    }

    // import_image_data (import_gltf.rs:119-160) + LoadImage: decoded on first use, once per glTF image
    const ImageSource* image(size_t index) {
        if (index >= images.size()) { set_error("gltf: texture names an unknown image"); return nullptr; }
        ImageSource& im = images[index];
        if (im.loaded) { if (!im.ok) set_error(im.where); return im.ok ? &im : nullptr; }
        im.loaded = true;
        const Json& ji = d.list("images")[index];
        std::vector<uint8_t> bytes;
        if (ji.has("uri")) {
            const std::string uri = percent_decode(ji.s("uri"));
            if (!read_uri(d.base, uri, bytes)) { im.where = g_error; return nullptr; }
        } else if (ji.has("bufferView")) {
            const int64_t bvi = ji.i("bufferView", -1);
            static const Json no_view;
            const Json& bv = bvi >= 0 ? d.list("bufferViews")[size_t(bvi)] : no_view;
            const int64_t bi = bv.i("buffer", -1);
            bool bv_ok = true;
            const size_t off = bv.size("byteOffset", 0, bv_ok), len = bv.size("byteLength", 0, bv_ok);
            if (!bv_ok || bi < 0 || size_t(bi) >= d.buffers.size() || off > d.buffers[size_t(bi)].size() || len > d.buffers[size_t(bi)].size() - off) { im.where = "gltf: image buffer view out of range"; set_error(im.where); return nullptr; }
            bytes.assign(d.buffers[size_t(bi)].begin() + off, d.buffers[size_t(bi)].begin() + off + len);
        } else { im.where = "gltf: image has neither uri nor bufferView"; set_error(im.where); return nullptr; }
        if (bytes.size() >= 4 && !memcmp(bytes.data(), "DDS ", 4)) {   // RawImage::Dds (image.rs:70-84): the file's own format and mips win over TexParams
            im.is_dds = true;
            if (!decode_dds(bytes.data(), bytes.size(), im.rgba, im.w, im.h, im.dds_mips, im.dds_srgb)) { im.where = g_error; return nullptr; }
        } else if (!decode_image(bytes.data(), bytes.size(), im.rgba, im.w, im.h)) { im.where = g_error; return nullptr; }
        im.ok = true; out->stats[3]++;
        return &im;
    }

    // KHR_texture_transform -> [f32; 6] (mesh.rs:127-146)
    static void texture_transform(const Json* info, float m[6]) {
        static const float dflt[6] = {1, 0, 0, 1, 0, 0};
        memcpy(m, dflt, sizeof dflt);
        const Json* ext = info ? info->get("extensions") : nullptr;
        const Json* tt = ext ? ext->get("KHR_texture_transform") : nullptr;
        if (!tt) return;
        const float r = float(tt->f("rotation", 0.0));
        float s[2] = {1, 1}, o[2] = {0, 0};
        if (const Json* js = tt->get("scale")) { s[0] = float((*js)[0].number_or(1)); s[1] = float((*js)[1].number_or(1)); }
        if (const Json* jo = tt->get("offset")) { o[0] = float((*jo)[0].number_or(0)); o[1] = float((*jo)[1].number_or(0)); }
        const float cr = float(std::cos(double(r))), sr = float(std::sin(double(r)));   // f32::cos / f32::sin, rounded once
        m[0] = cr * s[0]; m[1] = sr * s[1]; m[2] = -sr * s[0]; m[3] = cr * s[1]; m[4] = o[0]; m[5] = o[1];
    }

    bool make_map(const Json* info, const uint8_t placeholder[4], bool srgb, const uint32_t* swizzle, Texture& t) {
        if (!info) { t.texels.assign(placeholder, placeholder + 4); t.w = t.h = t.mips = 1; t.srgb = 0; return true; }   // CreatePlaceholderImage: 1x1, UNORM
        const Json& tex = d.list("textures")[size_t(info->i("index", -1))];
        if (!tex.is(Json::Object) || !tex.has("source")) { set_error("gltf: texture without an image source"); return false; }
        const ImageSource* im = image(size_t(tex.i("source", -1)));
        if (!im) return false;
        if (im->is_dds) { t.texels = im->rgba; t.w = im->w; t.h = im->h; t.mips = im->dds_mips; t.srgb = im->dds_srgb; return true; }   // process_dds (image.rs:285-335)
        build_mips(im->rgba.data(), im->w, im->h, true, swizzle, t.texels, t.w, t.h, t.mips);
        t.srgb = srgb ? 1 : 0;
        return true;
    }

    // load_gltf_material (mesh.rs:120-262): maps in the order normal, spec, albedo, emissive
    bool material(const Json* mat) {
        static const Json empty_obj = [] { Json j; j.kind = Json::Object; return j; }();
        const Json& m = mat ? *mat : empty_obj;
        const Json* pbr = m.get("pbrMetallicRoughness");
        const Json& p = pbr ? *pbr : empty_obj;
        kjb_mesh_material mm; memset(&mm, 0, sizeof mm);
        static const float dflt[6] = {1, 0, 0, 1, 0, 0};
        for (int k = 0; k < 4; ++k) memcpy(mm.map_transforms + 6 * k, dflt, sizeof dflt);

        const Json* albedo = p.get("baseColorTexture");
        if (!albedo) { const Json* ext = m.get("extensions"); const Json* sg = ext ? ext->get("KHR_materials_pbrSpecularGlossiness") : nullptr; if (sg) albedo = sg->get("diffuseTexture"); }
        const Json* normal = m.get("normalTexture");
        const Json* spec = p.get("metallicRoughnessTexture");
        const Json* emissive = m.get("emissiveTexture");
        if (albedo) texture_transform(albedo, mm.map_transforms + 0);
        if (spec) texture_transform(spec, mm.map_transforms + 12);
        if (emissive) texture_transform(emissive, mm.map_transforms + 18);

        static const uint8_t PH_NORMAL[4] = {127, 127, 255, 255}, PH_SPEC[4] = {255, 255, 127, 255}, PH_WHITE[4] = {255, 255, 255, 255};
        static const uint32_t SPEC_SWIZZLE[4] = {1, 2, 0, 3};
        const uint32_t base = uint32_t(out->maps.size());
        Texture t[4];
        if (!make_map(normal, PH_NORMAL, false, nullptr, t[0])) return false;
        if (!make_map(spec, PH_SPEC, false, SPEC_SWIZZLE, t[1])) return false;
        if (!make_map(albedo, PH_WHITE, true, nullptr, t[2])) return false;
        if (!make_map(emissive, PH_WHITE, true, nullptr, t[3])) return false;
        for (int k = 0; k < 4; ++k) { out->maps.push_back(std::move(t[k])); mm.maps[k] = base + uint32_t(k); }

        const Json* bc = p.get("baseColorFactor");
        for (int k = 0; k < 4; ++k) mm.base_color_mult[k] = bc ? float((*bc)[size_t(k)].number_or(1.0)) : 1.0f;
        mm.roughness_mult = float(p.f("roughnessFactor", 1.0));
        mm.metalness_factor = float(p.f("metallicFactor", 1.0));
        const Json* em = m.get("emissiveFactor");
        for (int k = 0; k < 3; ++k) mm.emissive[k] = em ? float((*em)[size_t(k)].number_or(0.0)) : 0.0f;
        mm.flags = 0;
        out->materials.push_back(mm);
        return true;
    }

    // returns false on error; *leave_node is set where the reference's closure `return`s (mesh.rs:306-318,353-355): the node's remaining
    // primitives are not visited, the material pushed for this one stays
    bool primitive(const Json& prim, const M4& xform, bool flip, bool* leave_node) {
        const uint32_t material_index = uint32_t(out->materials.size());
        const Json& mats = d.list("materials");
        const Json* mat = prim.has("material") ? &mats[size_t(prim.i("material", -1))] : nullptr;
        if (mat && !mat->is(Json::Object)) { set_error("gltf: primitive names an unknown material"); return false; }
        if (!material(mat)) return false;

        const Json* attrs = prim.get("attributes");
        if (!attrs || !attrs->has("POSITION") || !attrs->has("NORMAL")) { out->stats[2]++; *leave_node = true; return true; }
        AccessorData pos, nrm;
        if (!read_accessor(d, attrs->i("POSITION", -1), false, pos, "POSITION") || !read_accessor(d, attrs->i("NORMAL", -1), false, nrm, "NORMAL")) return false;
        if (pos.comps != 3 || nrm.comps != 3 || nrm.count < pos.count) { set_error("gltf: POSITION/NORMAL must be VEC3 of equal count"); return false; }
        const size_t nv = pos.count;

        std::vector<float> tangents(nv * 4, 0.0f);
        for (size_t i = 0; i < nv; ++i) tangents[4 * i] = 1.0f;
        if (attrs->has("TANGENT")) {
            AccessorData t; if (!read_accessor(d, attrs->i("TANGENT", -1), false, t, "TANGENT")) return false;
            if (t.comps == 4) for (size_t i = 0; i < std::min(nv, t.count) * 4; ++i) tangents[i] = t.f[i];
        }
        std::vector<float> uvs(nv * 2, 0.0f);
        if (attrs->has("TEXCOORD_0")) {
            AccessorData t; if (!read_accessor(d, attrs->i("TEXCOORD_0", -1), false, t, "TEXCOORD_0")) return false;
            if (t.comps == 2) for (size_t i = 0; i < std::min(nv, t.count) * 2; ++i) uvs[i] = norm_to_f32(t.f[i], t.ctype);
        }
        std::vector<float> colors(nv * 4, 1.0f);
        if (attrs->has("COLOR_0")) {
            AccessorData t; if (!read_accessor(d, attrs->i("COLOR_0", -1), false, t, "COLOR_0")) return false;
            if (t.comps == 3 || t.comps == 4) for (size_t i = 0; i < std::min(nv, t.count); ++i) for (int c = 0; c < t.comps; ++c) colors[4 * i + c] = norm_to_f32(t.f[i * t.comps + c], t.ctype);
        }
        std::vector<uint32_t> indices;
        if (prim.has("indices")) {
            AccessorData t; if (!read_accessor(d, prim.i("indices", -1), true, t, "indices")) return false;
            indices.swap(t.u);
        } else {
            if (nv == 0) { out->stats[2]++; *leave_node = true; return true; }
            if (prim.i("mode", 4) != 4) { set_error("gltf: non-indexed primitives must be triangle lists"); return false; }
            indices.resize(nv); for (size_t i = 0; i < nv; ++i) indices[i] = uint32_t(i);
        }
        for (uint32_t i : indices) if (i >= nv) { set_error("gltf: index out of range"); return false; }
        if (flip) for (size_t i = 0; i + 2 < indices.size(); i += 3) std::swap(indices[i], indices[i + 2]);

        const uint32_t base_index = uint32_t(out->positions.size() / 3);
        for (uint32_t i : indices) out->indices.push_back(i + base_index);
        out->colors.insert(out->colors.end(), colors.begin(), colors.end());
        out->material_ids.insert(out->material_ids.end(), nv, material_index);
        for (size_t i = 0; i < nv; ++i) {
            const float v[4] = {pos.f[3 * i], pos.f[3 * i + 1], pos.f[3 * i + 2], 1.0f}; float o[4];
            m4_mul_vec4(xform, v, o);
            out->positions.insert(out->positions.end(), o, o + 3);
        }
        for (size_t i = 0; i < nv; ++i) {
            const float v[4] = {nrm.f[3 * i], nrm.f[3 * i + 1], nrm.f[3 * i + 2], 0.0f}; float o[4];
            m4_mul_vec4(xform, v, o); normalize3(o);
            out->normals.insert(out->normals.end(), o, o + 3);
        }
        for (size_t i = 0; i < nv; ++i) {
            const float v[4] = {tangents[4 * i], tangents[4 * i + 1], tangents[4 * i + 2], 0.0f}; float o[4];
            m4_mul_vec4(xform, v, o); normalize3(o);
            o[3] = tangents[4 * i + 3] * (flip ? -1.0f : 1.0f);
            out->tangents.insert(out->tangents.end(), o, o + 4);
        }
        out->uvs.insert(out->uvs.end(), uvs.begin(), uvs.end());
        out->stats[1]++;
        return true;
    }

    M4 node_matrix(const Json& node) const {
        if (const Json* jm = node.get("matrix")) { M4 m; for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) m.c[c][r] = float((*jm)[size_t(c * 4 + r)].number_or(c == r ? 1.0 : 0.0)); return m; }
        float t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
        if (const Json* j = node.get("translation")) for (int k = 0; k < 3; ++k) t[k] = float((*j)[size_t(k)].number_or(0));
        if (const Json* j = node.get("rotation")) for (int k = 0; k < 4; ++k) q[k] = float((*j)[size_t(k)].number_or(k == 3 ? 1 : 0));
        if (const Json* j = node.get("scale")) for (int k = 0; k < 3; ++k) s[k] = float((*j)[size_t(k)].number_or(1));
        return gltf_trs(t, q, s);
    }

    // iter_gltf_node_tree (mesh.rs:100-112)
    bool walk(size_t node_index, const M4& parent, int depth) {
        const Json& nodes = d.list("nodes");
        const Json& node = nodes[node_index];
        if (!node.is(Json::Object)) { set_error("gltf: scene names an unknown node"); return false; }
        if (depth > 4096) { set_error("gltf: node hierarchy too deep (cycle?)"); return false; }
        const M4 xform = m4_mul(parent, node_matrix(node));
        out->stats[0]++;
        if (node.has("mesh")) {
            const Json& mesh = d.list("meshes")[size_t(node.i("mesh", -1))];
            if (!mesh.is(Json::Object)) { set_error("gltf: node names an unknown mesh"); return false; }
            const bool flip = m4_det(xform) < 0.0;
            const Json* prims = mesh.get("primitives");
            bool leave = false;
            for (size_t k = 0; prims && k < prims->size() && !leave; ++k) if (!primitive((*prims)[k], xform, flip, &leave)) return false;
        }
        if (const Json* ch = node.get("children")) for (size_t k = 0; k < ch->size(); ++k) if (!walk(size_t((*ch)[k].number_or(-1)), xform, depth + 1)) return false;
        return true;
    }

    bool run(const std::string& path, float scale, const float* rot) {
        if (!load_document(path)) return false;
        const Json& scenes = d.list("scenes");
        const Json* scene = nullptr;
        if (d.root.has("scene")) scene = &scenes[size_t(d.root.i("scene", 0))];
        if (!scene || !scene->is(Json::Object)) scene = scenes.size() ? &scenes[0] : nullptr;
        if (!scene || !scene->is(Json::Object)) { set_error("No default scene found in gltf"); return false; }
        const float s[3] = {scale, scale, scale}, q[4] = {rot ? rot[0] : 0.0f, rot ? rot[1] : 0.0f, rot ? rot[2] : 0.0f, rot ? rot[3] : 1.0f}, t[3] = {0, 0, 0};
        const M4 root = m4_from_srt(s, q, t);
        if (const Json* ns = scene->get("nodes")) for (size_t k = 0; k < ns->size(); ++k) if (!walk(size_t((*ns)[k].number_or(-1)), root, 0)) return false;
        out->map_descs.resize(out->maps.size());
        for (size_t i = 0; i < out->maps.size(); ++i) { const Texture& t2 = out->maps[i]; out->map_descs[i].texels = t2.texels.data(); out->map_descs[i].width = t2.w; out->map_descs[i].height = t2.h; out->map_descs[i].mip_count = t2.mips; out->map_descs[i].srgb = t2.srgb; }
        return true;
    }
};
}  // namespace

extern "C" {

int kjb_asset_load_gltf(const char* path, float scale, const float rotation_xyzw[4], kjb_asset** out) {
    if (out) *out = nullptr;
    if (!path || !out) { set_error("kjb_asset_load_gltf: null argument"); return 1; }
    Importer imp; imp.out = new kjb_asset();
    bool ok = false;
    try { ok = imp.run(path, scale, rotation_xyzw); }   // errors are values at this boundary: nothing may unwind into the caller
    catch (const std::exception& e) { set_error(std::string("kjb_asset_load_gltf: ") + e.what()); }
    if (!ok) { delete imp.out; return 1; }
    *out = imp.out;
    return 0;
}
void kjb_asset_destroy(kjb_asset* a) { delete a; }
const char* kjb_asset_last_error(void) { return g_error.c_str(); }

int kjb_asset_get_mesh(const kjb_asset* a, kjb_mesh_desc* out) {
    if (!a || !out) { set_error("kjb_asset_get_mesh: null argument"); return 1; }
    memset(out, 0, sizeof *out);
    out->positions = a->positions.data(); out->normals = a->normals.data(); out->uvs = a->uvs.data(); out->colors = a->colors.data();
    out->material_ids = a->material_ids.data(); out->indices = a->indices.data();
    out->vertex_count = uint32_t(a->positions.size() / 3); out->index_count = uint32_t(a->indices.size());
    out->materials = a->materials.data(); out->material_count = uint32_t(a->materials.size());
    out->maps = a->map_descs.data(); out->map_count = uint32_t(a->map_descs.size());
    out->use_lights = 0;
    return 0;
}
const float* kjb_asset_tangents(const kjb_asset* a) { return a ? a->tangents.data() : nullptr; }
int kjb_asset_stats(const kjb_asset* a, uint32_t out[4]) { if (!a || !out) return 1; memcpy(out, a->stats, sizeof a->stats); return 0; }

int kjb_asset_decode_image(const uint8_t* bytes, uint64_t byte_count, uint8_t** out_rgba8, uint32_t* out_width, uint32_t* out_height) {
    if (!bytes || !out_rgba8 || !out_width || !out_height) { set_error("kjb_asset_decode_image: null argument"); return 1; }
    std::vector<uint8_t> rgba; uint32_t w = 0, h = 0;
    try { if (!decode_image(bytes, size_t(byte_count), rgba, w, h)) return 1; }
    catch (const std::exception& e) { set_error(std::string("kjb_asset_decode_image: ") + e.what()); return 1; }
    uint8_t* p = static_cast<uint8_t*>(malloc(rgba.size() ? rgba.size() : 1));
    if (!p) { set_error("out of memory"); return 1; }
    memcpy(p, rgba.data(), rgba.size());
    *out_rgba8 = p; *out_width = w; *out_height = h;
    return 0;
}
int kjb_asset_build_mips(const uint8_t* rgba8, uint32_t width, uint32_t height, uint32_t use_mips, const uint32_t channel_swizzle[4],
                         uint8_t** out_texels, uint64_t* out_bytes, uint32_t* out_width, uint32_t* out_height, uint32_t* out_mip_count) {
    if (!rgba8 || !width || !height || !out_texels || !out_bytes || !out_width || !out_height || !out_mip_count) { set_error("kjb_asset_build_mips: bad argument"); return 1; }
    std::vector<uint8_t> texels; uint32_t w, h, levels;
    try { build_mips(rgba8, width, height, use_mips != 0, channel_swizzle, texels, w, h, levels); }
    catch (const std::exception& e) { set_error(std::string("kjb_asset_build_mips: ") + e.what()); return 1; }
    uint8_t* p = static_cast<uint8_t*>(malloc(texels.size()));
    if (!p) { set_error("out of memory"); return 1; }
    memcpy(p, texels.data(), texels.size());
    *out_texels = p; *out_bytes = texels.size(); *out_width = w; *out_height = h; *out_mip_count = levels;
    return 0;
}
void kjb_asset_free_buffer(void* p) { free(p); }

}  // extern "C"
