"""Scene sources for the benchmark configurations (BASELINE.json `configs`).

cornell_box(): the reference's bundled Cornell box (assets/meshes/cornell_box, 32 triangles, 8 untextured materials,
no emissive) placed like assets/scenes/cornell_box.ron: position (0,-1,0), scale 2.  The geometry is a committed data
fixture (kajiya_b200/assets/cornell_box.json, made by tests/golden/make_assets.py).
atrium()/ruins(): procedural "Sponza-class"/"Ruins-class" stand-ins (the reference bundles neither).
"""
import json, os
import numpy as np

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def blue_noise():
    return np.fromfile(os.path.join(_ASSETS, "bluenoise_256_rgba8.bin"), np.uint8).reshape(256, 256, 4)


def spatial_resolve_offsets():
    """SPATIAL_RESOLVE_OFFSETS (rtr.rs:402-915) as int32[512, 4]"""
    xy = np.fromfile(os.path.join(_ASSETS, "spatial_resolve_offsets_i16.bin"), np.int16).reshape(512, 2)
    out = np.zeros((512, 4), np.int32); out[:, :2] = xy
    return out


def cornell_box():
    j = json.load(open(os.path.join(_ASSETS, "cornell_box.json")))
    mesh = dict(positions=np.array(j["positions"], np.float32), normals=np.array(j["normals"], np.float32),
                indices=np.array(j["indices"], np.uint32), material_ids=np.array(j["material_ids"], np.uint32),
                materials=[dict(base_color=m["base_color"], roughness=m["roughness"], metallic=m["metallic"], emissive=m["emissive"]) for m in j["materials"]])
    transform = np.array([[2, 0, 0, 0], [0, 2, 0, -1], [0, 0, 2, 0]], np.float32)   # cornell_box.ron
    # camera looks down -Z into the open face; the bundled box has no emissive so the sun must shine into the opening
    view = dict(camera_position=(0.0, 1.0, 7.0), camera_rotation=(0.0, 0.0, 0.0, 1.0), sun_direction=_norm((0.3, 0.6, 1.0)))
    return [(mesh, [transform])], view


def _norm(v):
    v = np.asarray(v, np.float64); v = v / np.linalg.norm(v)
    return tuple(float(x) for x in v)


def _box(lo, hi, mat, P, N, I, M):
    lo = np.asarray(lo, np.float32); hi = np.asarray(hi, np.float32)
    faces = [((0, 0, -1), [(0, 0, 0), (0, 1, 0), (1, 1, 0), (1, 0, 0)]), ((0, 0, 1), [(0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]),
             ((-1, 0, 0), [(0, 0, 0), (0, 0, 1), (0, 1, 1), (0, 1, 0)]), ((1, 0, 0), [(1, 0, 0), (1, 1, 0), (1, 1, 1), (1, 0, 1)]),
             ((0, -1, 0), [(0, 0, 0), (1, 0, 0), (1, 0, 1), (0, 0, 1)]), ((0, 1, 0), [(0, 1, 0), (0, 1, 1), (1, 1, 1), (1, 1, 0)])]
    for n, corners in faces:
        base = len(P)
        for c in corners:
            P.append(lo + (hi - lo) * np.asarray(c, np.float32)); N.append(n); M.append(mat)
        I += [base, base + 1, base + 2, base, base + 2, base + 3]


def _grid_sheet(origin, du, dv, nu, nv, height_fn, mat, P, N, I, M):
    """tessellated, displaced sheet: the bulk of the triangle budget"""
    u = np.arange(nu + 1, dtype=np.float32) / nu; v = np.arange(nv + 1, dtype=np.float32) / nv
    uu, vv = np.meshgrid(u, v, indexing="ij")
    normal = np.cross(du, dv); normal = normal / np.linalg.norm(normal)
    h = height_fn(uu, vv).astype(np.float32)
    pts = origin[None, None, :] + uu[..., None] * du[None, None, :] + vv[..., None] * dv[None, None, :] + h[..., None] * normal[None, None, :]
    # finite-difference normals
    gu = np.gradient(pts, axis=0); gv = np.gradient(pts, axis=1)
    nn = np.cross(gu, gv); nn /= np.maximum(np.linalg.norm(nn, axis=-1, keepdims=True), 1e-20)
    base = len(P)
    P.extend(pts.reshape(-1, 3)); N.extend(nn.reshape(-1, 3)); M.extend([mat] * pts.shape[0] * pts.shape[1])
    ii, jj = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a = (ii * (nv + 1) + jj).ravel() + base; b = a + (nv + 1); c = b + 1; d = a + 1
    I.extend(np.stack([a, b, c, a, c, d], 1).ravel().tolist())


def atrium(seed=0xC0FFEE, target_tris=260_000):
    """Sponza-class: a colonnaded two-storey atrium, ~260 k triangles, 25 untextured materials (1x1 placeholder maps)."""
    rng = np.random.RandomState(seed & 0x7fffffff)
    mats = [dict(base_color=[*(0.25 + 0.6 * rng.rand(3)), 1.0], roughness=float(0.3 + 0.65 * rng.rand()), metallic=0.0, emissive=[0, 0, 0]) for _ in range(25)]
    P, N, I, M = [], [], [], []
    L, Wd, Ht = 24.0, 10.0, 9.0
    per_sheet = max(8, int(np.sqrt(target_tris / 2 / 6)))
    bump = lambda s, f: (lambda u, v: s * np.sin(u * f) * np.cos(v * f * 0.7))
    X, Y, Z = np.eye(3, dtype=np.float32)
    _grid_sheet(np.array([-L / 2, 0, -Wd / 2], np.float32), Z * Wd, X * L, per_sheet, per_sheet, bump(0.02, 40), 0, P, N, I, M)      # floor (normal +Y)
    _grid_sheet(np.array([-L / 2, 0, -Wd / 2], np.float32), Y * Ht, Z * Wd, per_sheet, per_sheet, bump(0.03, 25), 1, P, N, I, M)     # -X wall (normal +X)
    _grid_sheet(np.array([L / 2, 0, -Wd / 2], np.float32), Z * Wd, Y * Ht, per_sheet, per_sheet, bump(0.03, 25), 2, P, N, I, M)      # +X wall (normal -X)
    _grid_sheet(np.array([-L / 2, 0, -Wd / 2], np.float32), X * L, Y * Ht, per_sheet, per_sheet, bump(0.05, 30), 3, P, N, I, M)      # -Z wall (normal +Z)
    _grid_sheet(np.array([-L / 2, 0, Wd / 2], np.float32), Y * Ht, X * L, per_sheet, per_sheet, bump(0.05, 30), 4, P, N, I, M)       # +Z wall (normal -Z)
    # roof: two strips leaving a skylight in the middle so the sun gets in
    _grid_sheet(np.array([-L / 2, Ht, -Wd / 2], np.float32), X * L, Z * (Wd * 0.3), per_sheet, per_sheet // 3, bump(0.02, 20), 5, P, N, I, M)
    _grid_sheet(np.array([-L / 2, Ht, Wd * 0.2], np.float32), X * L, Z * (Wd * 0.3), per_sheet, per_sheet // 3, bump(0.02, 20), 5, P, N, I, M)
    # colonnade + gallery slabs + drapes
    for side in (-1, 1):
        for k in range(10):
            x = -L / 2 + 1.5 + k * (L - 3) / 9
            z = side * (Wd / 2 - 1.6)
            _box((x - 0.3, 0, z - 0.3), (x + 0.3, 4.0, z + 0.3), 6 + (k % 9), P, N, I, M)
            _box((x - 0.25, 4.4, z - 0.25), (x + 0.25, Ht, z + 0.25), 15 + (k % 9), P, N, I, M)
        _box((-L / 2, 4.0, side * (Wd / 2 - 2.2) - 0.9 * (side > 0) + 0.0), (L / 2, 4.4, side * (Wd / 2 - 2.2) + 0.9 * (side < 0) + 0.9 * (side > 0)), 24, P, N, I, M)
    mesh = dict(positions=np.asarray(P, np.float32), normals=np.asarray(N, np.float32), indices=np.asarray(I, np.uint32),
                material_ids=np.asarray(M, np.uint32), materials=mats)
    transform = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]], np.float32)
    view = dict(camera_position=(-9.0, 2.0, 0.0), camera_rotation=(0.0, -0.7071068, 0.0, 0.7071068), sun_direction=_norm((0.25, 1.0, 0.12)))
    return [(mesh, [transform])], view


def ruins(seed=0x5EED, grid=1024):
    """Ruins-class: 2*grid^2 = 2 097 152-triangle heightfield with broken pillars."""
    rng = np.random.RandomState(seed & 0x7fffffff)
    mats = [dict(base_color=[*(0.3 + 0.5 * rng.rand(3)), 1.0], roughness=float(0.5 + 0.45 * rng.rand()), metallic=0.0, emissive=[0, 0, 0]) for _ in range(8)]
    P, N, I, M = [], [], [], []
    S = 64.0
    ph = rng.rand(6) * 6.28
    def hf(u, v):
        h = 0.8 * np.sin(u * 9 + ph[0]) * np.cos(v * 7 + ph[1]) + 0.35 * np.sin(u * 31 + ph[2]) * np.sin(v * 27 + ph[3]) + 0.08 * np.sin(u * 140 + ph[4]) * np.cos(v * 160 + ph[5])
        return h
    X, Y, Z = np.eye(3, dtype=np.float32)
    _grid_sheet(np.array([-S / 2, 0, -S / 2], np.float32), Z * S, X * S, grid, grid, hf, 0, P, N, I, M)
    for k in range(40):
        x, z = (rng.rand(2) - 0.5) * (S * 0.8); hgt = 2 + 6 * rng.rand()
        _box((x - 0.5, -1, z - 0.5), (x + 0.5, hgt, z + 0.5), 1 + k % 7, P, N, I, M)
    mesh = dict(positions=np.asarray(P, np.float32), normals=np.asarray(N, np.float32), indices=np.asarray(I, np.uint32),
                material_ids=np.asarray(M, np.uint32), materials=mats)
    transform = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]], np.float32)
    view = dict(camera_position=(0.0, 4.0, 20.0), camera_rotation=(-0.0871557, 0.0, 0.0, 0.9961947), sun_direction=_norm((0.4, 0.7, 0.3)))
    return [(mesh, [transform])], view


def populate(world, scene):
    """add_mesh + add_instance for every (mesh, [transforms]) of a scene; uploads the blue-noise LUT."""
    for mesh, transforms in scene:
        h = world.add_mesh(mesh)
        for t in transforms:
            world.add_instance(h, t)
    world.set_blue_noise(blue_noise())
    world.set_spatial_resolve_offsets(spatial_resolve_offsets())
