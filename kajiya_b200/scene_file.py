"""kajiya's scene files (`assets/scenes/*.ron`) -> meshes + instances on a World.

Mirrors the `view` application's loader: `SceneDesc` / `SceneInstanceDesc` (crates/bin/view/src/scene.rs:1-18: position, scale = (1,1,1),
rotation = (0,0,0) Euler degrees, mesh path in the `/meshes/...` virtual file system), `RuntimeState::load_scene`
(crates/bin/view/src/runtime.rs:150-190) and `SceneElementTransform::affine_transform` (crates/bin/view/src/persisted.rs:271-283:
`Affine3A::from_scale_rotation_translation(scale, Quat::from_euler(YXZ, y, x, z), position)`).  Meshes go through the glTF importer
(`kajiya_b200.asset`), each distinct file once."""
import math, os, re
import numpy as np

F = np.float32


class RonError(ValueError):
    pass


def parse_ron(text):
    """the subset of RON the scene files use: structs `(name: value, ...)` / `Name(...)`, tuples `(a, b, c)`, lists, strings, numbers,
    bools, `//` and `/* */` comments, trailing commas.  Structs become dicts, tuples and lists become lists."""
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    pos = 0

    def ws():
        nonlocal pos
        while pos < len(text) and text[pos].isspace():
            pos += 1

    def fail(msg):
        raise RonError(f"{msg} at offset {pos}")

    def value():
        nonlocal pos
        ws()
        if pos >= len(text): fail("unexpected end of input")
        c = text[pos]
        if c == '"':
            m = re.compile(r'"((?:[^"\\]|\\.)*)"').match(text, pos)
            if not m: fail("unterminated string")
            pos = m.end()
            return bytes(m.group(1), "utf-8").decode("unicode_escape")
        if c == "[":
            pos += 1; out = []
            while True:
                ws()
                if pos < len(text) and text[pos] == "]": pos += 1; return out
                out.append(value()); ws()
                if pos < len(text) and text[pos] == ",": pos += 1
                elif pos < len(text) and text[pos] == "]": pos += 1; return out
                else: fail("expected ',' or ']'")
        m = re.compile(r"[A-Za-z_][A-Za-z0-9_]*").match(text, pos)
        if m and m.group(0) in ("true", "false"):
            pos = m.end(); return m.group(0) == "true"
        if m:   # struct name before '(' (or a unit / enum variant)
            pos = m.end(); ws()
            if pos < len(text) and text[pos] == "(":
                return paren()
            return m.group(0)
        if c == "(":
            return paren()
        m = re.compile(r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)").match(text, pos)
        if not m: fail(f"unexpected character {c!r}")
        pos = m.end()
        s = m.group(0)
        return float(s) if any(ch in s for ch in ".eE") else int(s)

    def paren():
        nonlocal pos
        pos += 1; ws()
        named = re.compile(r"[A-Za-z_][A-Za-z0-9_]*\s*:").match(text, pos) is not None
        out = {} if named else []
        while True:
            ws()
            if pos < len(text) and text[pos] == ")": pos += 1; return out
            if named:
                m = re.compile(r"([A-Za-z_][A-Za-z0-9_]*)\s*:").match(text, pos)
                if not m: fail("expected a field name")
                pos = m.end(); out[m.group(1)] = value()
            else:
                out.append(value())
            ws()
            if pos < len(text) and text[pos] == ",": pos += 1
            elif pos < len(text) and text[pos] == ")": pos += 1; return out
            else: fail("expected ',' or ')'")

    v = value(); ws()
    if pos != len(text): fail("trailing characters")
    return v


def instance_transform(position, rotation_euler_degrees=(0.0, 0.0, 0.0), scale=(1.0, 1.0, 1.0)):
    """SceneElementTransform::affine_transform as a row-major 3x4 float32 matrix (what World.add_instance takes)"""
    rx, ry, rz = (F(math.radians(float(a))) for a in rotation_euler_degrees)
    # glam Quat::from_euler(EulerRot::YXZ, a = y, b = x, c = z): q = qy(a) * qx(b) * qz(c)
    sy, cy = F(math.sin(ry * F(0.5))), F(math.cos(ry * F(0.5)))
    sx, cx = F(math.sin(rx * F(0.5))), F(math.cos(rx * F(0.5)))
    sz, cz = F(math.sin(rz * F(0.5))), F(math.cos(rz * F(0.5)))
    x, y, z, w = cy * sx * cz + sy * cx * sz, sy * cx * cz - cy * sx * sz, cy * cx * sz - sy * sx * cz, cy * cx * cz + sy * sx * sz
    x2, y2, z2 = x + x, y + y, z + z
    xx, xy, xz, yy, yz, zz, wx, wy, wz = x * x2, x * y2, x * z2, y * y2, y * z2, z * z2, w * x2, w * y2, w * z2
    cols = np.array([[F(1) - (yy + zz), xy + wz, xz - wy], [xy - wz, F(1) - (xx + zz), yz + wx], [xz + wy, yz - wx, F(1) - (xx + yy)]], F)   # glam quat_to_axes
    cols = (cols * np.asarray(scale, F)[:, None]).astype(F)
    m = np.zeros((3, 4), F)
    m[:, :3] = cols.T
    m[:, 3] = np.asarray(position, F)
    return m


def read_scene(path):
    """-> list of dict(mesh=str, position, rotation, scale) in file order"""
    desc = parse_ron(open(path, encoding="utf-8").read())
    if not isinstance(desc, dict) or not isinstance(desc.get("instances"), list):
        raise RonError(f"{path}: not a SceneDesc (no `instances` list)")
    out = []
    for i, inst in enumerate(desc["instances"]):
        if not isinstance(inst, dict) or "position" not in inst or "mesh" not in inst:
            raise RonError(f"{path}: instance {i} needs `position` and `mesh`")
        out.append(dict(mesh=str(inst["mesh"]), position=[float(v) for v in inst["position"]],
                        rotation=[float(v) for v in inst.get("rotation", (0.0, 0.0, 0.0))], scale=[float(v) for v in inst.get("scale", (1.0, 1.0, 1.0))]))
    return out


def load_scene(world, path, vfs_root, use_lights=False):
    """RuntimeState::load_scene: every instance's mesh through the glTF importer (one add_mesh per distinct file), then add_instance.
    `vfs_root` is the directory the `/meshes/...` paths are relative to (kajiya's `assets/`).  Returns the instance handles."""
    from . import asset
    meshes, handles = {}, []
    for inst in read_scene(path):
        rel = inst["mesh"].lstrip("/")
        full = os.path.join(vfs_root, rel)
        if full not in meshes:
            sc = asset.GltfScene(full)
            meshes[full] = world.add_mesh_desc(sc.desc, use_lights=use_lights)
            sc.close()
        handles.append(world.add_instance(meshes[full], instance_transform(inst["position"], inst["rotation"], inst["scale"])))
    return handles
