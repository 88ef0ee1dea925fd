"""Pins the oracle's integer/bit-level core against tests/golden/kat_vectors.json (independent pure-Python transcription of
the reference formulas, tests/golden/make_kat.py).  The reference ships no golden vectors of its own."""
import ctypes as C, json, os
import numpy as np

V = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat_vectors.json")))


def test_hashes(oracle_lib):
    d = oracle_lib.dll
    for f in (d.kjo_hash1, d.kjo_hash3, d.kjo_hash_combine2):
        f.restype = C.c_uint32
    d.kjo_hash1.argtypes = [C.c_uint32]; d.kjo_hash3.argtypes = [C.c_uint32] * 3; d.kjo_hash_combine2.argtypes = [C.c_uint32] * 2
    for x, h in V["hash1"]:
        assert d.kjo_hash1(x) == h
    for a, b, c, h in V["hash3"]:
        assert d.kjo_hash3(a, b, c) == h
    for a, b, h in V["hash_combine2"]:
        assert d.kjo_hash_combine2(a, b) == h
    d.kjo_u01.restype = C.c_float; d.kjo_u01.argtypes = [C.c_uint32]
    for x, u in V["u01"]:
        assert d.kjo_u01(x) == np.float32(u)


def test_packing(oracle_lib):
    d = oracle_lib.dll
    d.kjo_pack_normal_11_10_11.restype = C.c_uint32; d.kjo_pack_normal_11_10_11.argtypes = [C.c_float] * 3
    for x, y, z, p in V["pack_normal_11_10_11"]:
        assert d.kjo_pack_normal_11_10_11(x, y, z) == p
    out = (C.c_float * 3)()
    d.kjo_unpack_normal_11_10_11_no_normalize.argtypes = [C.c_uint32, C.POINTER(C.c_float * 3)]
    for p, x, y, z in V["unpack_normal_11_10_11"]:
        d.kjo_unpack_normal_11_10_11_no_normalize(p, C.byref(out))
        assert list(out) == [np.float32(x), np.float32(y), np.float32(z)]
    d.kjo_float3_to_rgb9e5.restype = C.c_uint32; d.kjo_float3_to_rgb9e5.argtypes = [C.c_float] * 3
    for x, y, z, p in V["rgb9e5"]:
        assert d.kjo_float3_to_rgb9e5(x, y, z) == p
    d.kjo_rgb9e5_to_float3.argtypes = [C.c_uint32, C.POINTER(C.c_float * 3)]
    for p, x, y, z in V["rgb9e5_dec"]:
        d.kjo_rgb9e5_to_float3(p, C.byref(out))
        assert list(out) == [np.float32(x), np.float32(y), np.float32(z)]
    d.kjo_pack_2x16f.restype = C.c_uint32; d.kjo_pack_2x16f.argtypes = [C.c_float] * 2
    for a, b, p in V["pack_2x16f"]:
        assert d.kjo_pack_2x16f(a, b) == p
    d.kjo_pack_color_888.restype = C.c_uint32; d.kjo_pack_color_888.argtypes = [C.c_float] * 3
    for x, y, z, p in V["color_888"]:
        assert d.kjo_pack_color_888(x, y, z) == p


def test_reservoir_stream(oracle_lib):
    d = oracle_lib.dll
    d.kjo_reservoir_stream.restype = C.c_uint32
    for s in V["reservoir_stream"]:
        n = len(s["w"])
        w = (C.c_float * n)(*s["w"]); p = (C.c_uint32 * n)(*s["payload"]); mw = (C.c_float * 2)(); rng = C.c_uint32()
        sel = d.kjo_reservoir_stream(C.c_uint32(s["seed"]), w, p, C.c_uint32(n), mw, C.byref(rng))
        assert sel == s["sel"] and rng.value == s["rng"]
        assert mw[0] == np.float32(s["M"]) and mw[1] == np.float32(s["w_sum"])


def test_halfres_offsets(oracle_lib):
    d = oracle_lib.dll
    o = (C.c_int * 2)()
    for f, x, y in V["halfres_offset"]:
        d.kjo_halfres_offset(C.c_uint32(f), o)
        assert (o[0], o[1]) == (x, y)
