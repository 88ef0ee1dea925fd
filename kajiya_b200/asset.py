"""ctypes view of libkjb_asset.so (include/kjb_asset.h): glTF scene -> TriangleMesh, encoded image -> RGBA8 mip chain.

Host-only library; mirrors kajiya-asset's LoadGltfScene / LoadImage / CreateGpuImage (see the header for file:line)."""
import ctypes as C, os
import numpy as np
from ._abi import MeshDesc, MeshMaterial, TextureDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
ASSET_SO = os.environ.get("KJB_ASSET_SO") or os.path.join(_HERE, "csrc", "libkjb_asset.so")   # KJB_ASSET_SO: the sanitizer build tests/test_asset.py makes of the same sources
_dll = None


def _lib():
    global _dll
    if _dll is None:
        if not os.path.exists(ASSET_SO):
            raise RuntimeError(f"{ASSET_SO} missing: run __graft_entry__.build()")
        d = C.CDLL(ASSET_SO, mode=C.RTLD_LOCAL)
        P = C.c_void_p
        d.kjb_asset_load_gltf.restype = C.c_int; d.kjb_asset_load_gltf.argtypes = [C.c_char_p, C.c_float, C.POINTER(C.c_float * 4), C.POINTER(P)]
        d.kjb_asset_destroy.restype = None; d.kjb_asset_destroy.argtypes = [P]
        d.kjb_asset_last_error.restype = C.c_char_p; d.kjb_asset_last_error.argtypes = []
        d.kjb_asset_get_mesh.restype = C.c_int; d.kjb_asset_get_mesh.argtypes = [P, C.POINTER(MeshDesc)]
        d.kjb_asset_tangents.restype = C.POINTER(C.c_float); d.kjb_asset_tangents.argtypes = [P]
        d.kjb_asset_stats.restype = C.c_int; d.kjb_asset_stats.argtypes = [P, C.POINTER(C.c_uint32 * 4)]
        d.kjb_asset_decode_image.restype = C.c_int
        d.kjb_asset_decode_image.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(P), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        d.kjb_asset_build_mips.restype = C.c_int
        d.kjb_asset_build_mips.argtypes = [P, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32 * 4), C.POINTER(P), C.POINTER(C.c_uint64),
                                           C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        d.kjb_asset_free_buffer.restype = None; d.kjb_asset_free_buffer.argtypes = [P]
        _dll = d
    return _dll


class AssetError(RuntimeError):
    pass


def _fail():
    raise AssetError(_lib().kjb_asset_last_error().decode(errors="replace"))


def _np(ptr, count, dtype):
    if not ptr or count == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count * np.dtype(dtype).itemsize,)).view(dtype).copy()


class GltfScene:
    """LoadGltfScene { path, scale, rotation } -> TriangleMesh.  `.desc` is the kjb_mesh_desc to hand to World.add_mesh_desc."""

    def __init__(self, path, scale=1.0, rotation=(0.0, 0.0, 0.0, 1.0)):
        d = _lib()
        self._h = C.c_void_p()
        rot = (C.c_float * 4)(*rotation)
        if d.kjb_asset_load_gltf(os.fsencode(path), scale, C.byref(rot), C.byref(self._h)) != 0:
            _fail()
        self.desc = MeshDesc()
        if d.kjb_asset_get_mesh(self._h, C.byref(self.desc)) != 0:
            _fail()

    def close(self):
        if self._h:
            _lib().kjb_asset_destroy(self._h); self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stats(self):
        s = (C.c_uint32 * 4)(); _lib().kjb_asset_stats(self._h, C.byref(s))
        return dict(nodes=s[0], primitives=s[1], skipped=s[2], images=s[3])

    def arrays(self):
        """numpy copies of the TriangleMesh streams + materials + maps (for tests and tools)"""
        m = self.desc; nv, ni = m.vertex_count, m.index_count
        out = dict(positions=_np(m.positions, nv * 3, np.float32).reshape(-1, 3), normals=_np(m.normals, nv * 3, np.float32).reshape(-1, 3),
                   uvs=_np(m.uvs, nv * 2, np.float32).reshape(-1, 2), colors=_np(m.colors, nv * 4, np.float32).reshape(-1, 4),
                   tangents=_np(_lib().kjb_asset_tangents(self._h), nv * 4, np.float32).reshape(-1, 4),
                   material_ids=_np(m.material_ids, nv, np.uint32), indices=_np(m.indices, ni, np.uint32))
        mats = []
        for i in range(m.material_count):
            mm = m.materials[i]
            mats.append(dict(base_color=list(mm.base_color_mult), maps=list(mm.maps), roughness=mm.roughness_mult, metallic=mm.metalness_factor,
                             emissive=list(mm.emissive), flags=mm.flags, map_transforms=np.array(list(mm.map_transforms), np.float32).reshape(4, 6)))
        maps = []
        for i in range(m.map_count):
            t = m.maps[i]
            n = sum(max(1, t.width >> l) * max(1, t.height >> l) * 4 for l in range(t.mip_count))
            maps.append(dict(texels=_np(t.texels, n, np.uint8), width=t.width, height=t.height, mips=t.mip_count, srgb=t.srgb))
        out["materials"] = mats; out["maps"] = maps
        return out


def decode_image(data: bytes):
    """LoadImage: PNG / JPEG bytes -> uint8[h, w, 4]"""
    d = _lib(); p = C.c_void_p(); w = C.c_uint32(); h = C.c_uint32()
    if d.kjb_asset_decode_image(data, len(data), C.byref(p), C.byref(w), C.byref(h)) != 0:
        _fail()
    try:
        return _np(p, w.value * h.value * 4, np.uint8).reshape(h.value, w.value, 4)
    finally:
        d.kjb_asset_free_buffer(p)


def build_mips(rgba8, use_mips=True, swizzle=None):
    """CreateGpuImage (uncompressed): uint8[h, w, 4] -> (list of uint8[h_l, w_l, 4], width, height)"""
    d = _lib(); a = np.ascontiguousarray(rgba8, np.uint8); h, w = a.shape[:2]
    p = C.c_void_p(); nb = C.c_uint64(); ow = C.c_uint32(); oh = C.c_uint32(); nl = C.c_uint32()
    sw = (C.c_uint32 * 4)(*swizzle) if swizzle is not None else None
    if d.kjb_asset_build_mips(a.ctypes.data, w, h, int(use_mips), C.byref(sw) if sw is not None else None, C.byref(p), C.byref(nb), C.byref(ow), C.byref(oh), C.byref(nl)) != 0:
        _fail()
    try:
        flat = _np(p, nb.value, np.uint8)
    finally:
        d.kjb_asset_free_buffer(p)
    levels, off = [], 0
    for l in range(nl.value):
        lw, lh = max(1, ow.value >> l), max(1, oh.value >> l)
        levels.append(flat[off:off + lw * lh * 4].reshape(lh, lw, 4)); off += lw * lh * 4
    return levels, ow.value, oh.value
