"""The intersection contract is ours (the reference delegates to the Vulkan driver): validate the oracle's BVH walker
against brute force over all triangles — same t and same triangle id, including tie-breaks."""
import ctypes as C, numpy as np
import parity
from kajiya_b200 import scenes


def _check(lib, scene, n=4000, seed=0):
    w = parity.make_world(lib, scene[0], 16, 16)
    w.render_frame(**scene[1])   # uploads geometry + builds the BVH
    rs = np.random.RandomState(seed)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rs.uniform(-3, 3, (n, 3)); rays[:, 3] = 0.0
    d = rs.randn(n, 3); rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True); rays[:, 7] = 1e4
    rays[: n // 8, 4:7] = np.round(rays[: n // 8, 4:7])   # axis-aligned rays hit shared edges/corners: exercises ties
    rays[: n // 8, 4] += (np.abs(rays[: n // 8, 4:7]).sum(1) == 0)
    t0, t1 = np.empty(n, np.float32), np.empty(n, np.float32); i0, i1 = np.empty(n, np.uint32), np.empty(n, np.uint32)
    f = lib.dll.kjo_trace_closest
    f(w.ctx, rays.ctypes.data_as(C.c_void_p), C.c_uint32(n), C.c_int(0), t0.ctypes.data_as(C.c_void_p), i0.ctypes.data_as(C.c_void_p))
    f(w.ctx, rays.ctypes.data_as(C.c_void_p), C.c_uint32(n), C.c_int(1), t1.ctypes.data_as(C.c_void_p), i1.ctypes.data_as(C.c_void_p))
    assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32)) and np.array_equal(i0, i1)
    assert (t0 > 0).mean() > 0.2


def test_bvh_vs_brute_force_cornell(oracle_lib):
    _check(oracle_lib, scenes.cornell_box())


def test_bvh_vs_brute_force_atrium(oracle_lib):
    scene, view = scenes.atrium(target_tris=6000)
    _check(oracle_lib, (scene, view), n=1500, seed=1)
