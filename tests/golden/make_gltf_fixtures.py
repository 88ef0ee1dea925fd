"""Writes the committed glTF / PNG / JPEG fixtures under tests/golden/gltf/ (row N5: asset import).

Everything here is synthetic (seeded numpy), nothing comes from /root/reference.  Re-run only when a fixture has to change:
    python tests/golden/make_gltf_fixtures.py

  courtyard.gltf + courtyard.bin + albedo.png + spec.png     JSON container: matrix / TRS / negative-scale nodes, interleaved and tightly packed
                                                             vertex streams, u8 / u16 / absent indices, normalised u16 uvs and u8 colours,
                                                             a sparse accessor, a primitive without NORMAL, a data: URI buffer and a data: URI image,
                                                             KHR_texture_transform, two scenes
  courtyard.glb                                              the same document as a binary container (BIN chunk, images through buffer views)
  png/*.png                                                  every colour type / bit depth, Adam7, every filter type, stored / fixed / dynamic deflate blocks
                                                             (written by the small encoder below so those paths are hit on purpose) + *.npy expectations
  jpg/*.jpg                                                  baseline 4:4:4 / 4:2:2 / 4:2:0, progressive, greyscale, restart intervals (PIL encoder)
"""
import base64, io, json, os, struct, zlib
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gltf")
rng = np.random.default_rng(20260922)


# ------------------------------------------------------------------ a PNG encoder that can hit every decoder path
def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)


def _paeth(a, b, c):
    p = a + b - c; pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if pa <= pb and pa <= pc else (b if pb <= pc else c)


def _filter_rows(rows, bpp, filters):
    """rows: list of bytes per scanline; filters: cyclic list of filter types"""
    out = bytearray(); prev = bytes(len(rows[0])) if rows else b""
    for y, row in enumerate(rows):
        ft = filters[y % len(filters)]; cur = bytearray(len(row))
        for i, v in enumerate(row):
            a = row[i - bpp] if i >= bpp else 0; b = prev[i]; c = prev[i - bpp] if i >= bpp else 0
            pred = [0, a, b, (a + b) >> 1, _paeth(a, b, c)][ft]
            cur[i] = (v - pred) & 255
        out.append(ft); out += cur; prev = row
    return bytes(out)


def _pack_samples(arr, depth):
    """arr: [h, w, channels] unsigned samples -> list of packed scanlines"""
    h, w, c = arr.shape; rows = []
    for y in range(h):
        flat = arr[y].reshape(-1)
        if depth == 8: rows.append(flat.astype(np.uint8).tobytes())
        elif depth == 16: rows.append(flat.astype(">u2").tobytes())
        else:
            bits = np.zeros(((w * c * depth + 7) // 8) * 8, np.uint8)
            for k in range(depth):
                bits[k:w * c * depth:depth][:len(flat)] = (flat >> (depth - 1 - k)) & 1
            rows.append(np.packbits(bits).tobytes())
    return rows


_ADAM7 = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]


def encode_png(arr, ctype, depth, interlace=False, filters=(0,), palette=None, trns=None, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, idat_split=0):
    h, w, c = arr.shape
    bpp = max(1, c * depth // 8)
    raw = b""
    passes = _ADAM7 if interlace else [(0, 0, 1, 1)]
    for x0, y0, dx, dy in passes:
        sub = arr[y0::dy, x0::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0: continue
        raw += _filter_rows(_pack_samples(sub, depth), bpp, list(filters))
    co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
    z = co.compress(raw) + co.flush()
    png = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1 if interlace else 0))
    if palette is not None: png += _chunk(b"PLTE", np.asarray(palette, np.uint8).tobytes())
    if trns is not None: png += _chunk(b"tRNS", bytes(trns))
    png += _chunk(b"tEXt", b"Comment\0synthetic fixture")
    if idat_split:
        for i in range(0, len(z), idat_split): png += _chunk(b"IDAT", z[i:i + idat_split])
    else:
        png += _chunk(b"IDAT", z)
    return png + _chunk(b"IEND", b"")


def narrow16(v):
    return ((v.astype(np.uint32) + 128) // 257).astype(np.uint8)


def make_pngs():
    d = os.path.join(OUT, "png"); os.makedirs(d, exist_ok=True)

    def save(name, data, expect):
        open(os.path.join(d, name + ".png"), "wb").write(data); np.save(os.path.join(d, name + ".npy"), expect.astype(np.uint8))

    def rgba_of(gray=None, rgb=None, alpha=None, shape=None):
        h, w = shape; o = np.full((h, w, 4), 255, np.uint8)
        if gray is not None: o[..., 0] = o[..., 1] = o[..., 2] = gray
        if rgb is not None: o[..., :3] = rgb
        if alpha is not None: o[..., 3] = alpha
        return o
    # greyscale, every bit depth, a transparent key, interlaced and not
    for depth in (1, 2, 4, 8, 16):
        h, w = 13, 19
        v = rng.integers(0, 1 << depth, (h, w, 1), dtype=np.uint32)
        key = int(v[3, 5, 0])
        g8 = narrow16(v[..., 0]) if depth == 16 else (v[..., 0] * (255 // ((1 << depth) - 1))).astype(np.uint8)
        for inter in (False, True):
            save(f"gray{depth}{'_adam7' if inter else ''}", encode_png(v, 0, depth, inter, filters=(0, 1, 2, 3, 4)), rgba_of(gray=g8, shape=(h, w)))
        save(f"gray{depth}_trns", encode_png(v, 0, depth, False, filters=(4, 3), trns=struct.pack(">H", key)), rgba_of(gray=g8, alpha=np.where(v[..., 0] == key, 0, 255), shape=(h, w)))
    # truecolour 8 / 16, with and without alpha, with a tRNS colour key
    for depth in (8, 16):
        h, w = 11, 23
        v = rng.integers(0, 1 << depth, (h, w, 4), dtype=np.uint32)
        n = (lambda a: narrow16(a)) if depth == 16 else (lambda a: a.astype(np.uint8))
        save(f"rgb{depth}", encode_png(v[..., :3], 2, depth, False, filters=(1, 4, 2)), rgba_of(rgb=n(v[..., :3]), shape=(h, w)))
        save(f"rgb{depth}_adam7", encode_png(v[..., :3], 2, depth, True, filters=(3,)), rgba_of(rgb=n(v[..., :3]), shape=(h, w)))
        save(f"rgba{depth}", encode_png(v, 6, depth, False, filters=(4,), idat_split=97), rgba_of(rgb=n(v[..., :3]), alpha=n(v[..., 3]), shape=(h, w)))
        save(f"rgba{depth}_adam7", encode_png(v, 6, depth, True, filters=(2, 0, 4)), rgba_of(rgb=n(v[..., :3]), alpha=n(v[..., 3]), shape=(h, w)))
        key = v[2, 7, :3]
        a = np.where((v[..., :3] == key).all(-1), 0, 255)
        save(f"rgb{depth}_trns", encode_png(v[..., :3], 2, depth, False, filters=(0,), trns=struct.pack(">HHH", *[int(k) for k in key])), rgba_of(rgb=n(v[..., :3]), alpha=a, shape=(h, w)))
        ga = v[..., :2]
        save(f"graya{depth}", encode_png(ga, 4, depth, False, filters=(1, 3)), rgba_of(gray=n(ga[..., 0]), alpha=n(ga[..., 1]), shape=(h, w)))
    # palette images, 1 / 2 / 4 / 8 bit, with a partial tRNS table
    for depth in (1, 2, 4, 8):
        h, w = 9, 17
        ncol = 1 << depth
        pal = rng.integers(0, 256, (ncol, 3), dtype=np.uint8)
        tr = rng.integers(0, 256, max(1, ncol // 2), dtype=np.uint8)
        idx = rng.integers(0, ncol, (h, w, 1), dtype=np.uint32)
        alpha = np.where(idx[..., 0] < len(tr), tr[np.minimum(idx[..., 0], len(tr) - 1)], 255)
        for inter in (False, True):
            save(f"pal{depth}{'_adam7' if inter else ''}", encode_png(idx, 3, depth, inter, filters=(0, 2), palette=pal, trns=tr), rgba_of(rgb=pal[idx[..., 0]], alpha=alpha, shape=(h, w)))
    # deflate block types: stored, fixed Huffman, dynamic Huffman with long matches; 1x1 and a wide flat image
    smooth = (np.add.outer(np.arange(64), np.arange(96)) % 256).astype(np.uint32)[..., None].repeat(3, -1)
    save("deflate_stored", encode_png(smooth, 2, 8, False, filters=(0,), level=0), rgba_of(rgb=smooth.astype(np.uint8), shape=smooth.shape[:2]))
    save("deflate_fixed", encode_png(smooth, 2, 8, False, filters=(1,), level=6, strategy=zlib.Z_FIXED), rgba_of(rgb=smooth.astype(np.uint8), shape=smooth.shape[:2]))
    save("deflate_dynamic", encode_png(smooth, 2, 8, False, filters=(4,), level=9), rgba_of(rgb=smooth.astype(np.uint8), shape=smooth.shape[:2]))
    one = np.array([[[200, 100, 50, 25]]], np.uint32)
    save("one_texel", encode_png(one, 6, 8), one.astype(np.uint8))
    flat = np.zeros((3, 300, 3), np.uint32); flat[..., 1] = 77
    save("flat_run", encode_png(flat, 2, 8, True, filters=(2,)), rgba_of(rgb=flat.astype(np.uint8), shape=(3, 300)))


def make_jpgs():
    from PIL import Image
    d = os.path.join(OUT, "jpg"); os.makedirs(d, exist_ok=True)
    yy, xx = np.mgrid[0:67, 0:93].astype(np.float64)
    img = np.stack([127 + 120 * np.sin(xx / 9.0) * np.cos(yy / 13.0), 127 + 110 * np.cos((xx + yy) / 17.0), 40 + 2.0 * xx + 0.3 * yy], -1)
    img = np.clip(img + rng.normal(0, 6, img.shape), 0, 255).astype(np.uint8)
    rgb = Image.fromarray(img, "RGB")
    rgb.save(os.path.join(d, "baseline_444.jpg"), quality=90, subsampling=0)
    rgb.save(os.path.join(d, "baseline_422.jpg"), quality=85, subsampling=1)
    rgb.save(os.path.join(d, "baseline_420.jpg"), quality=80, subsampling=2, optimize=True)
    rgb.save(os.path.join(d, "progressive_420.jpg"), quality=88, subsampling=2, progressive=True)
    rgb.save(os.path.join(d, "progressive_444.jpg"), quality=70, subsampling=0, progressive=True)
    rgb.save(os.path.join(d, "restart_420.jpg"), quality=75, subsampling=2, restart_marker_blocks=3)
    rgb.convert("L").save(os.path.join(d, "gray.jpg"), quality=92)
    rgb.convert("L").save(os.path.join(d, "gray_progressive.jpg"), quality=60, progressive=True)
    Image.fromarray(img[:8, :8], "RGB").save(os.path.join(d, "tiny_420.jpg"), quality=95, subsampling=2)
    Image.fromarray(img[:1, :1], "RGB").save(os.path.join(d, "one_texel.jpg"), quality=95, subsampling=2)


# ------------------------------------------------------------------ the glTF scene
def _grid(n):
    u = np.linspace(0, 1, n, dtype=np.float32)
    uu, vv = np.meshgrid(u, u, indexing="xy")
    pos = np.stack([uu * 2 - 1, 0.15 * np.sin(uu * 5) * np.cos(vv * 4), vv * 2 - 1], -1).reshape(-1, 3).astype(np.float32)
    nrm = np.tile(np.array([0.1, 1.0, -0.05], np.float32), (n * n, 1)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    uv = np.stack([uu, vv], -1).reshape(-1, 2)
    idx = []
    for y in range(n - 1):
        for x in range(n - 1):
            a = y * n + x; idx += [a, a + n, a + 1, a + 1, a + n, a + n + 1]
    return pos, nrm.astype(np.float32), uv.astype(np.float32), np.array(idx, np.uint32)


def _cube():
    P, N, I = [], [], []
    faces = [((0, 0, -1), [(0, 0, 0), (0, 1, 0), (1, 1, 0), (1, 0, 0)]), ((0, 0, 1), [(0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]),
             ((-1, 0, 0), [(0, 0, 0), (0, 0, 1), (0, 1, 1), (0, 1, 0)]), ((1, 0, 0), [(1, 0, 0), (1, 1, 0), (1, 1, 1), (1, 0, 1)]),
             ((0, -1, 0), [(0, 0, 0), (1, 0, 0), (1, 0, 1), (0, 0, 1)]), ((0, 1, 0), [(0, 1, 0), (0, 1, 1), (1, 1, 1), (1, 1, 0)])]
    for n, cs in faces:
        b = len(P)
        for c in cs: P.append(c); N.append(n)
        I += [b, b + 1, b + 2, b, b + 2, b + 3]
    return np.array(P, np.float32) - 0.5, np.array(N, np.float32), np.array(I, np.uint32)


def make_scene():
    from PIL import Image
    os.makedirs(OUT, exist_ok=True)
    # textures
    yy, xx = np.mgrid[0:21, 0:37]
    albedo = np.stack([(xx * 7) % 256, (yy * 12) % 256, ((xx // 4 + yy // 4) % 2) * 200 + 30], -1).astype(np.uint8)
    Image.fromarray(albedo, "RGB").save(os.path.join(OUT, "albedo tex.png"))
    spec = rng.integers(0, 256, (16, 16, 4), dtype=np.uint8)
    Image.fromarray(spec, "RGBA").save(os.path.join(OUT, "spec.png"))
    em = encode_png(rng.integers(0, 4, (8, 8, 1), dtype=np.uint32), 3, 2, False, palette=[[0, 0, 0], [255, 180, 40], [40, 200, 255], [255, 255, 255]])
    albedo_png = open(os.path.join(OUT, "albedo tex.png"), "rb").read(); spec_png = open(os.path.join(OUT, "spec.png"), "rb").read()

    gp, gn, guv, gi = _grid(5)
    cp, cn, ci = _cube()
    bin0 = bytearray(); views = []; accessors = []

    def add_view(data, stride=None, buffer=0):
        while len(bin0) % 4: bin0.append(0)
        views.append(dict(buffer=buffer, byteOffset=len(bin0), byteLength=len(data), **({"byteStride": stride} if stride else {})))
        bin0.extend(data); return len(views) - 1

    def add_acc(view, ct, typ, count, off=0, **kw):
        accessors.append(dict(bufferView=view, componentType=ct, type=typ, count=count, **({"byteOffset": off} if off else {}), **kw)); return len(accessors) - 1
    # grid: one interleaved stream (pos f32x3 | nrm f32x3 | uv u16x2 normalised | colour u8x4 normalised), stride 32
    inter = bytearray()
    gcol = rng.integers(60, 256, (len(gp), 4), dtype=np.uint8)
    guv16 = np.round(guv * 65535).astype(np.uint16)
    for i in range(len(gp)):
        inter += gp[i].tobytes() + gn[i].tobytes() + guv16[i].tobytes() + gcol[i].tobytes()
    v_inter = add_view(bytes(inter), stride=32)
    a_gpos = add_acc(v_inter, 5126, "VEC3", len(gp), 0, min=gp.min(0).tolist(), max=gp.max(0).tolist())
    a_gnrm = add_acc(v_inter, 5126, "VEC3", len(gp), 12)
    a_guv = add_acc(v_inter, 5123, "VEC2", len(gp), 24, normalized=True)
    a_gcol = add_acc(v_inter, 5121, "VEC4", len(gp), 28, normalized=True)
    a_gidx = add_acc(add_view(gi.astype(np.uint16).tobytes()), 5123, "SCALAR", len(gi))
    # a non-indexed fan of 2 triangles (6 vertices), f32 uvs, RGB f32 colours
    fp = np.array([[0, 0.5, 0], [1, 0.5, 0], [0, 0.5, 1], [1, 0.5, 0], [1, 0.5, 1], [0, 0.5, 1]], np.float32)
    fn = np.tile(np.array([0, 1, 0], np.float32), (6, 1)); fuv = fp[:, [0, 2]].copy(); fcol = rng.random((6, 3)).astype(np.float32)
    a_fpos = add_acc(add_view(fp.tobytes()), 5126, "VEC3", 6); a_fnrm = add_acc(add_view(fn.tobytes()), 5126, "VEC3", 6)
    a_fuv = add_acc(add_view(fuv.tobytes()), 5126, "VEC2", 6); a_fcol = add_acc(add_view(fcol.tobytes()), 5126, "VEC3", 6)
    # cube in a data: URI buffer (buffer 1), u8 indices, tangents
    bin1 = bytearray(); views1 = []

    def add_view1(data):
        while len(bin1) % 4: bin1.append(0)
        views.append(dict(buffer=1, byteOffset=len(bin1), byteLength=len(data))); bin1.extend(data); return len(views) - 1
    ctan = np.concatenate([np.roll(cn, 1, axis=1), np.where(np.arange(len(cn))[:, None] % 2 == 0, 1.0, -1.0)], 1).astype(np.float32)
    a_cpos = add_acc(add_view1(cp.tobytes()), 5126, "VEC3", len(cp)); a_cnrm = add_acc(add_view1(cn.tobytes()), 5126, "VEC3", len(cp))
    a_ctan = add_acc(add_view1(ctan.tobytes()), 5126, "VEC4", len(cp)); a_cidx = add_acc(add_view1(ci.astype(np.uint8).tobytes()), 5121, "SCALAR", len(ci))
    # sparse: the fan positions with two vertices lifted
    sp_idx = np.array([1, 4], np.uint16); sp_val = np.array([[1, 0.9, 0], [1, 0.9, 1]], np.float32)
    v_si = add_view(sp_idx.tobytes()); v_sv = add_view(sp_val.tobytes())
    accessors.append(dict(bufferView=views.index(views[accessors[a_fpos]["bufferView"]]), componentType=5126, type="VEC3", count=6,
                          sparse=dict(count=2, indices=dict(bufferView=v_si, componentType=5123), values=dict(bufferView=v_sv))))
    a_spos = len(accessors) - 1

    materials = [
        dict(name="tiled", pbrMetallicRoughness=dict(baseColorFactor=[0.9, 0.8, 0.7, 1.0], roughnessFactor=0.85, metallicFactor=0.25,
             baseColorTexture=dict(index=0, extensions=dict(KHR_texture_transform=dict(offset=[0.125, 0.25], rotation=0.4, scale=[3.0, 2.0]))),
             metallicRoughnessTexture=dict(index=1))),
        dict(name="plain", pbrMetallicRoughness=dict(baseColorFactor=[0.2, 0.6, 0.3, 1.0], roughnessFactor=0.5, metallicFactor=0.0)),
        dict(name="glow", emissiveFactor=[2.0, 1.5, 0.5], emissiveTexture=dict(index=2), normalTexture=dict(index=1),
             pbrMetallicRoughness=dict(roughnessFactor=0.3)),
        dict(name="legacy", extensions=dict(KHR_materials_pbrSpecularGlossiness=dict(diffuseTexture=dict(index=0)))),
    ]
    meshes = [
        dict(name="grid", primitives=[dict(attributes=dict(POSITION=a_gpos, NORMAL=a_gnrm, TEXCOORD_0=a_guv, COLOR_0=a_gcol), indices=a_gidx, material=0),
                                      dict(attributes=dict(POSITION=a_fpos, NORMAL=a_fnrm, TEXCOORD_0=a_fuv, COLOR_0=a_fcol), material=1, mode=4)]),
        dict(name="cube", primitives=[dict(attributes=dict(POSITION=a_cpos, NORMAL=a_cnrm, TANGENT=a_ctan), indices=a_cidx, material=2)]),
        dict(name="broken", primitives=[dict(attributes=dict(POSITION=a_fpos, NORMAL=a_fnrm), material=3),
                                        dict(attributes=dict(POSITION=a_fpos), material=1),            # no NORMAL: the node is left here
                                        dict(attributes=dict(POSITION=a_fpos, NORMAL=a_fnrm), material=1)]),   # never reached
        dict(name="sparse", primitives=[dict(attributes=dict(POSITION=a_spos, NORMAL=a_fnrm))]),        # no material: defaults
    ]
    s2 = float(np.sqrt(0.5))
    nodes = [
        dict(name="root", matrix=[1.5, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 2.0, 0, 0.25, -0.5, 0.75, 1], children=[1, 2, 6]),
        dict(name="floor", mesh=0, rotation=[0.0, 0.3826834, 0.0, 0.9238795], translation=[0.0, -0.25, 0.0], scale=[2.0, 1.0, 2.0]),
        dict(name="mirrored", mesh=1, scale=[-1.0, 1.0, 1.0], translation=[0.5, 0.25, -0.25], children=[3]),
        dict(name="broken", mesh=2, translation=[0.0, 1.0, 0.0]),
        dict(name="floor again", mesh=0, translation=[3.0, 0.0, 0.0], rotation=[s2, 0.0, 0.0, s2]),
        dict(name="sparse", mesh=3, translation=[-2.0, 0.5, 0.0]),
        dict(name="camera rig", translation=[0, 2, 5], children=[7]),
        dict(name="empty leaf"),
    ]
    doc = dict(asset=dict(version="2.0", generator="kajiya_b200 tests/golden/make_gltf_fixtures.py"), extensionsUsed=["KHR_texture_transform", "KHR_materials_pbrSpecularGlossiness"],
               scene=1, scenes=[dict(name="unused", nodes=[7]), dict(name="main", nodes=[0, 4, 5])], nodes=nodes, meshes=meshes, materials=materials,
               textures=[dict(source=0), dict(source=1), dict(source=2)],
               images=[dict(uri="albedo%20tex.png"), dict(uri="spec.png"), dict(uri="data:image/png;base64," + base64.b64encode(em).decode())],
               accessors=accessors, bufferViews=views,
               buffers=[dict(uri="courtyard.bin", byteLength=len(bin0)), dict(uri="data:application/octet-stream;base64," + base64.b64encode(bytes(bin1)).decode(), byteLength=len(bin1))])
    open(os.path.join(OUT, "courtyard.bin"), "wb").write(bytes(bin0))
    json.dump(doc, open(os.path.join(OUT, "courtyard.gltf"), "w"), indent=1)

    # GLB: one BIN chunk = bin0 | bin1 | images ; buffer views of buffer 1 are rebased
    glb_bin = bytearray(bin0)
    while len(glb_bin) % 4: glb_bin.append(0)
    base1 = len(glb_bin); glb_bin += bin1
    gviews = [dict(v) for v in views]
    for v in gviews:
        if v["buffer"] == 1: v["buffer"] = 0; v["byteOffset"] += base1
    gimages = []
    for data in (albedo_png, spec_png, em):
        while len(glb_bin) % 4: glb_bin.append(0)
        gviews.append(dict(buffer=0, byteOffset=len(glb_bin), byteLength=len(data))); glb_bin += data
        gimages.append(dict(bufferView=len(gviews) - 1, mimeType="image/png"))
    gdoc = dict(doc); gdoc["bufferViews"] = gviews; gdoc["images"] = gimages; gdoc["buffers"] = [dict(byteLength=len(glb_bin))]
    js = json.dumps(gdoc, separators=(",", ":")).encode()
    js += b" " * (-len(js) % 4)
    while len(glb_bin) % 4: glb_bin.append(0)
    total = 12 + 8 + len(js) + 8 + len(glb_bin)
    open(os.path.join(OUT, "courtyard.glb"), "wb").write(struct.pack("<4sII", b"glTF", 2, total) + struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(glb_bin), 0x004E4942) + bytes(glb_bin))


if __name__ == "__main__":
    make_pngs(); make_jpgs(); make_scene()
    n = sum(os.path.getsize(os.path.join(b, f)) for b, _, fs in os.walk(OUT) for f in fs)
    print("wrote", OUT, n, "bytes")
