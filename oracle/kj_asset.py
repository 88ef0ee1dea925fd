"""CPU restatement of the asset path — TEST INFRASTRUCTURE ONLY: nothing under kajiya_b200/ may import this.

numpy restatement of kajiya-asset's glTF import, written independently of kajiya_b200/csrc/host/kjb_asset*.cpp (python json / base64 /
struct instead of the C++ parser, PIL as the PNG/JPEG bit-stream decoder, vectorised numpy float32 for the arithmetic):
  * load_gltf_scene      <- LoadGltfScene::run            crates/lib/kajiya-asset/src/mesh.rs:278-441
  * _node_tree           <- iter_gltf_node_tree            mesh.rs:100-112
  * load_gltf_material   <- load_gltf_material             mesh.rs:120-262
  * _buffers / _image    <- import_buffer_data / import_image_data   import_gltf.rs:89-160
  * process_rgba8        <- CreateGpuImage::process_rgba8 (TexCompressionMode::None branch)   image.rs:130-283
  * resize_lanczos3      <- image::imageops::resize(FilterType::Lanczos3) of the `image` crate 0.23.14 (not vendored in the reference tree)

PARITY UNPINNED: the reference has no tests or golden vectors for its asset crate and no Rust toolchain exists here to run it, so this
oracle is pinned only against (a) the glTF 2.0 / PNG specifications through independent decoders (python json, PIL) on the reference's
own bundled assets and on committed fixtures and (b) hand-computed expectations in tests/test_asset.py.  The float arithmetic of the
third-party crates (gltf b9c04be Transform::matrix, glam 0.22 Mat4/Vec3) is restated from their published sources.
"""
import base64, json, os, struct, io
import numpy as np

F = np.float32


# ------------------------------------------------------------------ f32 algebra in glam / gltf::math operation order
def _mul_vec4(m, v):
    """glam Mat4 * Vec4 for rows of v: ((x_axis*v.x + y_axis*v.y) + z_axis*v.z) + w_axis*v.w ; m[col][row]"""
    v = np.asarray(v, F)
    s = m[0] * v[..., 0:1]
    s = s + m[1] * v[..., 1:2]
    s = s + m[2] * v[..., 2:3]
    s = s + m[3] * v[..., 3:4]
    return s.astype(F)


def _mat_mul(a, b):
    return np.stack([_mul_vec4(a, b[j]) for j in range(4)]).astype(F)


def _from_scale_rotation_translation(scale, q, t):
    x, y, z, w = [F(c) for c in q]
    x2, y2, z2 = x + x, y + y, z + z
    xx, xy, xz, yy, yz, zz, wx, wy, wz = x * x2, x * y2, x * z2, y * y2, y * z2, z * z2, w * x2, w * y2, w * z2
    m = np.zeros((4, 4), F); m[3, 3] = 1
    m[0, :3] = np.array([F(1) - (yy + zz), xy + wz, xz - wy], F) * F(scale[0])
    m[1, :3] = np.array([xy - wz, F(1) - (xx + zz), yz + wx], F) * F(scale[1])
    m[2, :3] = np.array([xz + wy, yz - wx, F(1) - (xx + yy)], F) * F(scale[2])
    m[3, :3] = np.asarray(t, F)
    return m


def _gltf_mul(a, b):
    """gltf::math Matrix4 product: every entry a left-to-right 4-term dot (m[col][row])"""
    o = np.zeros((4, 4), F)
    for j in range(4):
        for r in range(4):
            s = a[0, r] * b[j, 0]
            s = F(s + a[1, r] * b[j, 1]); s = F(s + a[2, r] * b[j, 2]); s = F(s + a[3, r] * b[j, 3])
            o[j, r] = s
    return o


def _node_matrix(node):
    if "matrix" in node:
        return np.array(node["matrix"], np.float64).astype(F).reshape(4, 4)   # column-major list -> m[col][row]
    t = np.array(node.get("translation", [0, 0, 0]), np.float64).astype(F)
    q = np.array(node.get("rotation", [0, 0, 0, 1]), np.float64).astype(F)
    s = np.array(node.get("scale", [1, 1, 1]), np.float64).astype(F)
    T = np.eye(4, dtype=F); T[3, :3] = t
    x, y, z, w = q
    x2, y2, z2 = x + x, y + y, z + z
    xx2, xy2, xz2, yy2, yz2, zz2, sy2, sz2, sx2 = x2 * x, x2 * y, x2 * z, y2 * y, y2 * z, z2 * z, y2 * w, z2 * w, x2 * w
    R = np.eye(4, dtype=F)
    R[0, :3] = [F(1) - yy2 - zz2, xy2 + sz2, xz2 - sy2]
    R[1, :3] = [xy2 - sz2, F(1) - xx2 - zz2, yz2 + sx2]
    R[2, :3] = [xz2 + sy2, yz2 - sx2, F(1) - xx2 - yy2]
    S = np.eye(4, dtype=F); S[0, 0], S[1, 1], S[2, 2] = s
    return _gltf_mul(_gltf_mul(T, R), S)


def _normalize(v):
    d = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]).astype(F)
    d = (d + v[:, 2] * v[:, 2]).astype(F)
    inv = (F(1) / np.sqrt(d)).astype(F)
    return (v * inv[:, None]).astype(F)


# ------------------------------------------------------------------ document
def _read_uri(base, uri):
    if ":" in uri:
        if uri.startswith("data:"):
            payload = uri.split(";base64,", 1)[1] if ";base64," in uri else uri[5:]
            return base64.b64decode(payload)
        if uri.startswith("file://"):
            return open(uri[7:], "rb").read()
        if uri.startswith("file:"):
            return open(uri[5:], "rb").read()
        raise ValueError("unsupported URI scheme")
    return open(os.path.join(base, uri), "rb").read()


def _document(path):
    raw = open(path, "rb").read()
    blob = None
    if raw[:4] == b"glTF":
        version, total = struct.unpack_from("<II", raw, 4)
        off, doc = 12, None
        while off + 8 <= total:
            clen, ctype = struct.unpack_from("<II", raw, off)
            chunk = raw[off + 8:off + 8 + clen]
            if ctype == 0x4E4F534A and doc is None:
                doc = json.loads(chunk.decode("utf-8"))
            elif ctype == 0x004E4942 and blob is None:
                blob = chunk
            off += 8 + clen
    else:
        doc = json.loads(raw.decode("utf-8-sig"))
    base = os.path.dirname(path) or "."
    buffers = []
    for b in doc.get("buffers", []):
        data = _read_uri(base, b["uri"]) if "uri" in b else blob
        assert data is not None and len(data) >= b["byteLength"]
        buffers.append(bytes(data) + b"\0" * (-len(data) % 4))
    return doc, buffers, base


_CT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NC = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT2": 4, "MAT3": 9, "MAT4": 16}


def _view(doc, buffers, bv_index, byte_offset, dtype, ncomp, count):
    bv = doc["bufferViews"][bv_index]
    buf = buffers[bv["buffer"]]
    off = bv.get("byteOffset", 0) + byte_offset
    esz = np.dtype(dtype).itemsize * ncomp
    stride = bv.get("byteStride", esz)
    out = np.zeros((count, ncomp), dtype)
    for i in range(count):
        out[i] = np.frombuffer(buf, dtype, ncomp, off + i * stride)
    return out


def _accessor(doc, buffers, index):
    a = doc["accessors"][index]
    dt, nc, n = _CT[a["componentType"]], _NC[a["type"]], a["count"]
    if "bufferView" in a:
        bv = doc["bufferViews"][a["bufferView"]]
        esz = np.dtype(dt).itemsize * nc
        if bv.get("byteStride", esz) == esz:   # fast path: tightly packed
            out = np.frombuffer(buffers[bv["buffer"]], dt, n * nc, bv.get("byteOffset", 0) + a.get("byteOffset", 0)).reshape(n, nc).copy()
        else:
            out = _view(doc, buffers, a["bufferView"], a.get("byteOffset", 0), dt, nc, n)
    else:
        out = np.zeros((n, nc), dt)
    if "sparse" in a:
        sp = a["sparse"]
        idx = _view(doc, buffers, sp["indices"]["bufferView"], sp["indices"].get("byteOffset", 0), _CT[sp["indices"]["componentType"]], 1, sp["count"])[:, 0]
        val = _view(doc, buffers, sp["values"]["bufferView"], sp["values"].get("byteOffset", 0), dt, nc, sp["count"])
        out[idx.astype(np.int64)] = val
    return out, a["componentType"]


def _to_f32(arr, ct):
    a = arr.astype(F)
    if ct == 5121: return (a / F(255)).astype(F)
    if ct == 5123: return (a / F(65535)).astype(F)
    if ct == 5120: return np.maximum(a / F(127), F(-1)).astype(F)
    if ct == 5122: return np.maximum(a / F(32767), F(-1)).astype(F)
    return a


# ------------------------------------------------------------------ images
def decode_image(data):
    """LoadImage::run -> uint8[h, w, 4] (8-bit sources; PIL does the entropy decoding)"""
    from PIL import Image
    im = Image.open(io.BytesIO(data)); im.load()
    if im.mode in ("I;16", "I;16B", "I"):
        raise NotImplementedError("16-bit sources are checked against analytic expectations instead")
    return np.asarray(im.convert("RGBA"), np.uint8).copy()


def _sinc(t):
    t = np.asarray(t, F)
    a = (t * F(np.pi)).astype(F)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = (np.sin(a.astype(np.float64)).astype(F) / a).astype(F)
    return np.where(t == 0, F(1), r).astype(F)


def _lanczos3(x):
    x = np.asarray(x, F)
    return np.where(np.abs(x) < F(3), _sinc(x) * _sinc((x / F(3)).astype(F)), F(0)).astype(F)


def _taps(out_i, in_n, out_n):
    ratio = F(in_n) / F(out_n)
    sratio = F(1) if ratio < 1 else ratio
    support = F(3) * sratio
    centre = F((F(out_i) + F(0.5)) * ratio)
    left = int(min(max(int(np.floor(F(centre - support))), 0), in_n - 1))
    right = int(min(max(int(np.ceil(F(centre + support))), left + 1), in_n))
    c = F(centre - F(0.5))
    w = _lanczos3(((np.arange(left, right).astype(F) - c) / sratio).astype(F))
    s = F(0)
    for x in w:
        s = F(s + x)
    return left, (w / s).astype(F)


def resize_lanczos3(src, dw, dh):
    """imageops::resize: vertical pass into f32 (unrounded), horizontal pass, clamp, round half away from zero"""
    sh, sw = src.shape[:2]
    srcf = src.astype(F)
    tmp = np.zeros((dh, sw, 4), F)
    for oy in range(dh):
        left, w = _taps(oy, sh, dh)
        acc = np.zeros((sw, 4), F)
        for k, wk in enumerate(w):
            acc = (acc + srcf[left + k] * wk).astype(F)
        tmp[oy] = acc
    out = np.zeros((dh, dw, 4), np.uint8)
    for ox in range(dw):
        left, w = _taps(ox, sw, dw)
        acc = np.zeros((dh, 4), F)
        for k, wk in enumerate(w):
            acc = (acc + tmp[:, left + k] * wk).astype(F)
        v = np.clip(acc, F(0), F(255))
        out[:, ox] = np.where(v - np.floor(v) >= F(0.5), np.floor(v) + 1, np.floor(v)).astype(np.uint8)
    return out


def process_rgba8(rgba, use_mips=True, swizzle=None):
    """-> list of mip levels (uint8[h, w, 4]), level 0 first"""
    img = np.ascontiguousarray(rgba, np.uint8)
    h, w = img.shape[:2]
    if w > 2048 or h > 2048:
        w2, h2 = min(w, 2048), min(h, 2048)
        img = resize_lanczos3(img, w2, h2); w, h = w2, h2
    levels = max(int(w).bit_length(), int(h).bit_length()) if use_mips else 1

    def finish(level):
        level = level.copy()
        if swizzle is not None:   # in-place, channel after channel (image.rs:214-223)
            for c in range(4):
                level[..., c] = level[..., swizzle[c]]
        return level

    out = []
    for l in range(levels):
        out.append(finish(img))
        if l + 1 < levels:
            w, h = max(1, w // 2), max(1, h // 2)
            img = resize_lanczos3(img, w, h)
    return out


# ------------------------------------------------------------------ materials + scene
_DEFAULT_XF = [1.0, 0.0, 0.0, 1.0, 0.0, 0.0]


def _texture_transform(info):
    tt = (info or {}).get("extensions", {}).get("KHR_texture_transform")
    if tt is None:
        return np.array(_DEFAULT_XF, F)
    r = F(tt.get("rotation", 0.0)); s = np.array(tt.get("scale", [1, 1]), np.float64).astype(F); o = np.array(tt.get("offset", [0, 0]), np.float64).astype(F)
    c, sn = F(np.cos(np.float64(r))), F(np.sin(np.float64(r)))
    return np.array([c * s[0], sn * s[1], -sn * s[0], c * s[1], o[0], o[1]], F)


class _Images:
    def __init__(self, doc, buffers, base):
        self.doc, self.buffers, self.base, self.cache = doc, buffers, base, {}

    def get(self, index):
        if index not in self.cache:
            ji = self.doc["images"][index]
            if "uri" in ji:
                from urllib.parse import unquote
                data = _read_uri(self.base, unquote(ji["uri"]))
            else:
                bv = self.doc["bufferViews"][ji["bufferView"]]
                data = self.buffers[bv["buffer"]][bv.get("byteOffset", 0):bv.get("byteOffset", 0) + bv["byteLength"]]
            self.cache[index] = ("dds",) + decode_dds(bytes(data)) if bytes(data[:4]) == b"DDS " else decode_image(data)
        return self.cache[index]


def load_gltf_material(doc, mat, images):
    mat = mat or {}
    pbr = mat.get("pbrMetallicRoughness", {})
    xf = np.tile(np.array(_DEFAULT_XF, F), (4, 1))
    albedo = pbr.get("baseColorTexture") or mat.get("extensions", {}).get("KHR_materials_pbrSpecularGlossiness", {}).get("diffuseTexture")
    normal, spec, emissive = mat.get("normalTexture"), pbr.get("metallicRoughnessTexture"), mat.get("emissiveTexture")
    if albedo: xf[0] = _texture_transform(albedo)
    if spec: xf[2] = _texture_transform(spec)
    if emissive: xf[3] = _texture_transform(emissive)

    def make(info, placeholder, srgb, swizzle):
        if not info:
            return dict(levels=[np.array(placeholder, np.uint8).reshape(1, 1, 4)], srgb=0)
        src = doc["textures"][info["index"]]["source"]
        img = images.get(src)
        if isinstance(img, tuple) and img[0] == "dds":   # RawImage::Dds: the file's mips and format, TexParams ignored (image.rs:285-335)
            return dict(levels=img[1], srgb=img[2])
        return dict(levels=process_rgba8(img, True, swizzle), srgb=int(srgb))

    maps = [make(normal, [127, 127, 255, 255], False, None), make(spec, [255, 255, 127, 255], False, [1, 2, 0, 3]),
            make(albedo, [255, 255, 255, 255], True, None), make(emissive, [255, 255, 255, 255], True, None)]
    material = dict(base_color=np.array(pbr.get("baseColorFactor", [1, 1, 1, 1]), np.float64).astype(F), roughness=F(pbr.get("roughnessFactor", 1.0)),
                    metallic=F(pbr.get("metallicFactor", 1.0)), emissive=np.array(mat.get("emissiveFactor", [0, 0, 0]), np.float64).astype(F), flags=0, map_transforms=xf)
    return maps, material


def load_gltf_scene(path, scale=1.0, rotation=(0.0, 0.0, 0.0, 1.0)):
    doc, buffers, base = _document(path)
    images = _Images(doc, buffers, base)
    scenes = doc.get("scenes", [])
    scene = scenes[doc["scene"]] if "scene" in doc and doc["scene"] < len(scenes) else (scenes[0] if scenes else None)
    if scene is None:
        raise ValueError("No default scene found in gltf")
    res = dict(positions=[], normals=[], colors=[], uvs=[], tangents=[], material_ids=[], indices=[], materials=[], maps=[])
    count = [0]

    def process(node, xform):
        if "mesh" not in node:
            return
        flip = np.linalg.det(xform.astype(np.float64)) < 0
        for prim in doc["meshes"][node["mesh"]]["primitives"]:
            mi = len(res["materials"])
            maps, material = load_gltf_material(doc, doc["materials"][prim["material"]] if "material" in prim else None, images)
            material["maps"] = [len(res["maps"]) + k for k in range(4)]
            res["materials"].append(material); res["maps"] += maps
            at = prim.get("attributes", {})
            if "POSITION" not in at or "NORMAL" not in at:
                return
            pos = _accessor(doc, buffers, at["POSITION"])[0].astype(F); nrm = _accessor(doc, buffers, at["NORMAL"])[0].astype(F)
            nv = len(pos)
            tan = np.tile(np.array([1, 0, 0, 0], F), (nv, 1))
            if "TANGENT" in at: tan = _accessor(doc, buffers, at["TANGENT"])[0].astype(F)
            uvs = np.zeros((nv, 2), F)
            if "TEXCOORD_0" in at:
                a, ct = _accessor(doc, buffers, at["TEXCOORD_0"]); uvs = _to_f32(a, ct)
            col = np.ones((nv, 4), F)
            if "COLOR_0" in at:
                a, ct = _accessor(doc, buffers, at["COLOR_0"]); col[:, :a.shape[1]] = _to_f32(a, ct)
            if "indices" in prim:
                idx = _accessor(doc, buffers, prim["indices"])[0][:, 0].astype(np.uint32)
            else:
                if nv == 0:
                    return
                assert prim.get("mode", 4) == 4
                idx = np.arange(nv, dtype=np.uint32)
            if flip:
                full = len(idx) // 3 * 3
                tri = idx[:full].reshape(-1, 3)[:, ::-1].reshape(-1)
                idx = np.concatenate([tri, idx[full:]])
            res["indices"].append(idx + np.uint32(count[0])); res["colors"].append(col); res["material_ids"].append(np.full(nv, mi, np.uint32))
            res["positions"].append(_mul_vec4(xform, np.concatenate([pos, np.ones((nv, 1), F)], 1))[:, :3])
            res["normals"].append(_normalize(_mul_vec4(xform, np.concatenate([nrm, np.zeros((nv, 1), F)], 1))[:, :3]))
            t3 = _normalize(_mul_vec4(xform, np.concatenate([tan[:, :3], np.zeros((nv, 1), F)], 1))[:, :3])
            res["tangents"].append(np.concatenate([t3, (tan[:, 3:4] * F(-1.0 if flip else 1.0)).astype(F)], 1))
            res["uvs"].append(uvs)
            count[0] += nv

    def walk(ni, xform):
        node = doc["nodes"][ni]
        xf = _mat_mul(xform, _node_matrix(node))
        process(node, xf)
        for c in node.get("children", []):
            walk(c, xf)

    root = _from_scale_rotation_translation([scale] * 3, rotation, [0, 0, 0])
    for ni in scene.get("nodes", []):
        walk(ni, root)

    def cat(k, shape, dt):
        return np.concatenate(res[k]).astype(dt) if res[k] else np.zeros(shape, dt)
    return dict(positions=cat("positions", (0, 3), F), normals=cat("normals", (0, 3), F), colors=cat("colors", (0, 4), F), uvs=cat("uvs", (0, 2), F),
                tangents=cat("tangents", (0, 4), F), material_ids=cat("material_ids", (0,), np.uint32), indices=cat("indices", (0,), np.uint32),
                materials=res["materials"], maps=res["maps"])


# ------------------------------------------------------------------ DDS (image.rs:70-84, 285-335): DX10 header, BC1_SRGB / BC3 / BC5 -> RGBA8 mip chain
def _bc1_palette(c0, c1, force_four):
    def expand(c):
        r, g, b = (c >> 11) & 31, (c >> 5) & 63, c & 31
        return np.stack([(r << 3) | (r >> 2), (g << 2) | (g >> 4), (b << 3) | (b >> 2)], -1).astype(np.int32)
    e0, e1 = expand(c0.astype(np.int32)), expand(c1.astype(np.int32))
    four = (c0 > c1) | force_four
    p2 = np.where(four[:, None], (2 * e0 + e1 + 1) // 3, (e0 + e1) // 2)
    p3 = np.where(four[:, None], (e0 + 2 * e1 + 1) // 3, 0)
    a3 = np.where(four, 255, 0)
    pal = np.zeros((len(c0), 4, 4), np.int32)
    pal[:, 0, :3], pal[:, 1, :3], pal[:, 2, :3], pal[:, 3, :3] = e0, e1, p2, p3
    pal[:, :3, 3] = 255; pal[:, 3, 3] = a3
    return pal


def _bc4_values(b0, b1, snorm):
    if snorm:
        s0, s1 = b0.astype(np.int8).astype(np.int32), b1.astype(np.int8).astype(np.int32)
        a0, a1 = np.maximum(s0.astype(F) / F(127), F(-1)), np.maximum(s1.astype(F) / F(127), F(-1)); six = s0 > s1; lo, hi = F(-1), F(1)
    else:
        a0, a1 = b0.astype(F) / F(255), b1.astype(F) / F(255); six = b0 > b1; lo, hi = F(0), F(1)
    vals = np.zeros((len(b0), 8), F); vals[:, 0], vals[:, 1] = a0, a1
    for i in range(1, 7):
        v6 = ((F(7 - i) * a0 + F(i) * a1) / F(7)).astype(F)
        v4 = ((F(5 - i) * a0 + F(i) * a1) / F(5)).astype(F) if i < 5 else (np.full(len(b0), lo if i == 5 else hi, F))
        vals[:, i + 1] = np.where(six, v6, v4)
    u = (vals * F(0.5) + F(0.5)).astype(F) if snorm else vals
    return np.floor(np.clip(u, F(0), F(1)) * F(255) + F(0.5)).astype(np.uint8)


def decode_dds(data):
    """-> (list of uint8[h, w, 4] levels, srgb)"""
    assert data[:4] == b"DDS " and struct.unpack_from("<I", data, 4)[0] == 124
    H, W = struct.unpack_from("<II", data, 12); mips = max(1, struct.unpack_from("<I", data, 28)[0])
    assert struct.unpack_from("<I", data, 84)[0] == 0x30315844, "only DX10-header files"
    dxgi = struct.unpack_from("<I", data, 128)[0]
    kind, srgb = {72: ("bc1", 1), 77: ("bc3", 0), 78: ("bc3", 1), 83: ("bc5u", 0), 84: ("bc5s", 0)}[dxgi]
    bb = 8 if kind == "bc1" else 16
    off, levels = 148, []
    for l in range(mips):
        w, h = max(1, W >> l), max(1, H >> l); bw, bh = (max(w, 4) + 3) // 4, (max(h, 4) + 3) // 4
        blocks = np.frombuffer(data, np.uint8, bw * bh * bb, off).reshape(bw * bh, bb); off += bw * bh * bb
        out = np.zeros((bw * bh, 16, 4), np.uint8)

        def colours(cb, force_four):
            c0 = cb[:, 0].astype(np.uint32) | (cb[:, 1].astype(np.uint32) << 8); c1 = cb[:, 2].astype(np.uint32) | (cb[:, 3].astype(np.uint32) << 8)
            pal = _bc1_palette(c0, c1, force_four)
            idx = cb[:, 4].astype(np.uint32) | (cb[:, 5].astype(np.uint32) << 8) | (cb[:, 6].astype(np.uint32) << 16) | (cb[:, 7].astype(np.uint32) << 24)
            sel = (idx[:, None] >> (2 * np.arange(16, dtype=np.uint32))[None, :]) & 3
            return np.take_along_axis(pal, sel[:, :, None].astype(np.int64).repeat(4, 2), 1).astype(np.uint8)

        def alphas(ab, snorm):
            vals = _bc4_values(ab[:, 0], ab[:, 1], snorm)
            bits = np.zeros(len(ab), np.uint64)
            for i in range(6): bits |= ab[:, 2 + i].astype(np.uint64) << np.uint64(8 * i)
            sel = (bits[:, None] >> (np.uint64(3) * np.arange(16, dtype=np.uint64))[None, :]) & np.uint64(7)
            return np.take_along_axis(vals, sel.astype(np.int64), 1)
        if kind == "bc1": out[:] = colours(blocks, False)
        elif kind == "bc3": out[:] = colours(blocks[:, 8:], True); out[:, :, 3] = alphas(blocks[:, :8], False)
        else:
            out[:, :, 0] = alphas(blocks[:, :8], kind == "bc5s"); out[:, :, 1] = alphas(blocks[:, 8:], kind == "bc5s"); out[:, :, 2] = 0; out[:, :, 3] = 255
        img = out.reshape(bh, bw, 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(bh * 4, bw * 4, 4)[:h, :w]
        levels.append(np.ascontiguousarray(img))
    return levels, srgb
