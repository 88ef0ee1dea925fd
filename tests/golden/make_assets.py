"""Regenerates the small DATA assets the repo needs from the reference checkout (run in the build container only;
/root/reference does not exist on the GPU box, so the outputs are committed).

  kajiya_b200/assets/cornell_box.json    <- /root/reference/assets/meshes/cornell_box/scene.{gltf,bin}
        (32 triangles; node transforms baked into the vertices exactly like kajiya-asset's
         iter_gltf_node_tree, crates/lib/kajiya-asset/src/mesh.rs:100-113,282-290)
  kajiya_b200/assets/bluenoise_256_rgba8.bin <- /root/reference/assets/images/bluenoise/256_256/LDR_RGBA_0.png
        (the 256x256 RGBA8 blue-noise LUT bound at bindless slot 1, inc/bindless_textures.hlsl:11; raw texels)

  kajiya_b200/assets/spatial_resolve_offsets_i16.bin <- the SPATIAL_RESOLVE_OFFSETS constant table of
        /root/reference/crates/lib/kajiya/src/renderers/rtr.rs:402-915 (512 integer (x, y) sample offsets that the host pushes to the
        reflection passes as shader constants; stored as int16[512, 2], the zero z/w lanes are re-added on load)

No reference SOURCE code is copied: all outputs are data assets, re-encoded.
"""
import json, struct, sys, os
import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(__file__), "..", "..", "kajiya_b200", "assets")


def quat_to_mat(q):
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=np.float64)


def node_matrix(n):
    m = np.eye(4)
    if "matrix" in n:
        return np.array(n["matrix"], dtype=np.float64).reshape(4, 4).T
    if "scale" in n:
        m[:3, :3] = np.diag(n["scale"])
    if "rotation" in n:
        m[:3, :3] = quat_to_mat(n["rotation"]) @ m[:3, :3]
    if "translation" in n:
        m[:3, 3] = n["translation"]
    return m


def main():
    g = json.load(open(f"{REF}/assets/meshes/cornell_box/scene.gltf"))
    buf = open(f"{REF}/assets/meshes/cornell_box/" + g["buffers"][0]["uri"], "rb").read()

    def accessor(i):
        a = g["accessors"][i]; bv = g["bufferViews"][a["bufferView"]]
        off = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
        ncomp = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4}[a["type"]]
        dt = {5126: np.float32, 5125: np.uint32, 5123: np.uint16, 5121: np.uint8}[a["componentType"]]
        arr = np.frombuffer(buf, dtype=dt, count=a["count"] * ncomp, offset=off)
        return arr.reshape(a["count"], ncomp) if ncomp > 1 else arr

    positions, normals, indices, material_ids = [], [], [], []
    base = 0

    def walk(ni, xf):
        nonlocal base
        n = g["nodes"][ni]; xf = xf @ node_matrix(n)
        if "mesh" in n:
            for prim in g["meshes"][n["mesh"]]["primitives"]:
                p = accessor(prim["attributes"]["POSITION"]).astype(np.float64)
                nn = accessor(prim["attributes"]["NORMAL"]).astype(np.float64)
                idx = accessor(prim["indices"]).astype(np.uint32)
                pw = (xf[:3, :3] @ p.T).T + xf[:3, 3]
                nw = (xf[:3, :3] @ nn.T).T
                nw /= np.linalg.norm(nw, axis=1, keepdims=True)
                positions.append(pw.astype(np.float32)); normals.append(nw.astype(np.float32))
                indices.append(idx + base); material_ids.append(np.full(len(p), prim["material"], np.uint32))
                base += len(p)
        for c in n.get("children", []):
            walk(c, xf)

    for ni in g["scenes"][g.get("scene", 0)]["nodes"]:
        walk(ni, np.eye(4))
    mats = []
    for m in g["materials"]:
        pbr = m.get("pbrMetallicRoughness", {})
        mats.append({"name": m["name"], "base_color": pbr.get("baseColorFactor", [1, 1, 1, 1]), "roughness": pbr.get("roughnessFactor", 1.0),
                     "metallic": pbr.get("metallicFactor", 1.0), "emissive": m.get("emissiveFactor", [0, 0, 0])})
    out = {"source": "assets/meshes/cornell_box/scene.gltf (node transforms baked)",
           "positions": np.concatenate(positions).round(7).tolist(), "normals": np.concatenate(normals).round(7).tolist(),
           "indices": np.concatenate(indices).tolist(), "material_ids": np.concatenate(material_ids).tolist(), "materials": mats}
    os.makedirs(OUT, exist_ok=True)
    json.dump(out, open(f"{OUT}/cornell_box.json", "w"))
    print("cornell:", len(out["positions"]), "verts", len(out["indices"]) // 3, "tris",
          "aabb", np.concatenate(positions).min(0), np.concatenate(positions).max(0))

    from PIL import Image
    im = np.array(Image.open(f"{REF}/assets/images/bluenoise/256_256/LDR_RGBA_0.png").convert("RGBA"), dtype=np.uint8)
    assert im.shape == (256, 256, 4)
    im.tofile(f"{OUT}/bluenoise_256_rgba8.bin")
    print("blue noise:", im.shape, im.mean())

    import re
    src = open(f"{REF}/crates/lib/kajiya/src/renderers/rtr.rs").read()
    body = src[src.index("pub const SPATIAL_RESOLVE_OFFSETS"):]
    body = body[body.index("= [") + 3:body.index("];")]
    tup = re.findall(r"\(\s*(-?\d+)i32,\s*(-?\d+)i32,\s*0,\s*0\)", body)
    assert len(tup) == 16 * 4 * 8, len(tup)
    offs = np.array(tup, dtype=np.int16)
    offs.tofile(f"{OUT}/spatial_resolve_offsets_i16.bin")
    print("spatial resolve offsets:", offs.shape, offs.min(), offs.max())


if __name__ == "__main__":
    main()
