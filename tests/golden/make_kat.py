"""Generates tests/golden/kat_vectors.json: known answers for the integer/bit-level core of the path, computed by an
INDEPENDENT pure-Python transcription of the formulas in the reference's HLSL (hash.hlsl:7-55, pack_unpack.hlsl:1-164,
reservoir.hlsl:47-59, frame_constants.hlsl:235-250) using numpy float32 arithmetic.  Neither the oracle nor the kernels
are involved, so both can be checked against these vectors.  The reference itself ships no golden data."""
import json, os
import numpy as np

f32 = np.float32
M32 = 0xffffffff


def hash1(x):
    x = (x + (x << 10)) & M32; x ^= x >> 6; x = (x + (x << 3)) & M32; x ^= x >> 11; x = (x + (x << 15)) & M32
    return x


def hash_combine2(x, y):
    seed = ((x * 1664525 + y + 1013904223) & M32) * 1664525 & M32
    seed ^= seed >> 11; seed ^= (seed << 7) & 0x9d2c5680; seed ^= (seed << 15) & 0xefc60000; seed ^= seed >> 18
    return seed & M32


def hash3(x, y, z):
    return hash_combine2(x, hash_combine2(y, hash1(z)))


def u01(h):
    return float(np.array([(h & 0x007FFFFF) | 0x3F800000], np.uint32).view(np.float32)[0] - f32(1.0))


def pack_unorm(v, bits):
    mx = (1 << bits) - 1
    return int(f32(min(max(f32(v), f32(0)), f32(1))) * f32(mx) + f32(0.5))


def pack_normal(n):
    return pack_unorm(f32(n[0]) * f32(0.5) + f32(0.5), 11) + (pack_unorm(f32(n[1]) * f32(0.5) + f32(0.5), 10) << 11) + (pack_unorm(f32(n[2]) * f32(0.5) + f32(0.5), 11) << 21)


def unpack_normal(p):
    def un(v, bits):
        mx = (1 << bits) - 1
        return f32(v & mx) / f32(mx)
    return [float(un(p, 11) * f32(2) - f32(1)), float(un(p >> 11, 10) * f32(2) - f32(1)), float(un(p >> 21, 11) * f32(2) - f32(1))]


def rgb9e5(rgb):
    MAXV = f32(511.0 / 512.0) * f32(65536.0)
    c = [min(max(f32(v), f32(0)), MAXV) for v in rgb]
    m = max(c)
    fl = ((int(np.array([m], np.float32).view(np.uint32)[0]) & 0x7F800000) >> 23) - 127
    e = max(-16, fl) + 1 + 15
    denom = f32(2.0) ** f32(e - 15 - 9)
    if int(np.floor(m / denom + f32(0.5))) == 512:
        denom = denom * f32(2); e += 1
    r, g, b = (int(np.floor(v / denom + f32(0.5))) for v in c)
    return (r << 23) | (g << 14) | (b << 5) | e


def rgb9e5_dec(v):
    s = f32(2.0) ** f32((v & 31) - 15 - 9)
    return [float(f32((v >> 23) & 511) * s), float(f32((v >> 14) & 511) * s), float(f32((v >> 5) & 511) * s)]


def f16bits(v):
    return int(np.array([v], np.float32).astype(np.float16).view(np.uint16)[0])


def reservoir_stream(seed, ws, payloads):
    rng, w_sum, M, payload = seed, f32(0), f32(0), 0
    for w, p in zip(ws, payloads):
        w_sum = f32(w_sum + f32(w)); M = f32(M + f32(1))
        dart = f32(u01(rng)); rng = hash1(rng)
        prob = f32(w) / w_sum
        if prob >= dart:
            payload = p
    return payload, float(M), float(w_sum), rng


def main():
    rs = np.random.RandomState(1234)
    out = {}
    xs = [0, 1, 2, 0xdeadbeef, 0xffffffff] + [int(v) for v in rs.randint(0, 2**32, 27, dtype=np.uint64)]
    out["hash1"] = [[x, hash1(x)] for x in xs]
    out["hash3"] = [[a, b, c, hash3(a, b, c)] for a, b, c in zip(xs, xs[3:] + xs[:3], xs[7:] + xs[:7])]
    out["hash_combine2"] = [[a, b, hash_combine2(a, b)] for a, b in zip(xs, xs[5:] + xs[:5])]
    out["u01"] = [[x, u01(x)] for x in xs]
    ns = rs.randn(32, 3).astype(np.float32); ns /= np.linalg.norm(ns, axis=1, keepdims=True)
    out["pack_normal_11_10_11"] = [[*map(float, n), pack_normal(n)] for n in ns]
    out["unpack_normal_11_10_11"] = [[p, *unpack_normal(p)] for p in (pack_normal(n) for n in ns)]
    cols = np.abs(rs.randn(32, 3).astype(np.float32)) * np.float32(3.0)
    cols[0] = [0, 0, 0]; cols[1] = [1e5, 2, 3]; cols[2] = [1e-6, 1e-7, 0]
    out["rgb9e5"] = [[*map(float, c), rgb9e5(c)] for c in cols]
    out["rgb9e5_dec"] = [[v, *rgb9e5_dec(v)] for v in (rgb9e5(c) for c in cols)]
    hv = rs.randn(32, 2).astype(np.float32) * np.float32(10)
    out["pack_2x16f"] = [[float(a), float(b), f16bits(a) | (f16bits(b) << 16)] for a, b in hv]
    out["color_888"] = [[*map(float, c), pack_unorm(np.sqrt(f32(c[0])), 8) + (pack_unorm(np.sqrt(f32(c[1])), 8) << 8) + (pack_unorm(np.sqrt(f32(c[2])), 8) << 16)]
                        for c in rs.rand(16, 3).astype(np.float32)]
    streams = []
    for k in range(8):
        n = 3 + k
        ws = [float(v) for v in rs.rand(n).astype(np.float32)]; ps = [int(v) for v in rs.randint(0, 2**32, n, dtype=np.uint64)]
        seed = int(rs.randint(0, 2**32, dtype=np.uint64))
        payload, M, wsum, rng = reservoir_stream(seed, ws, ps)
        streams.append({"seed": seed, "w": ws, "payload": ps, "sel": payload, "M": M, "w_sum": wsum, "rng": rng})
    out["reservoir_stream"] = streams
    out["halfres_offset"] = [[f, *[(1, 1), (1, 0), (0, 0), (0, 1)][f & 3]] for f in range(8)]
    json.dump(out, open(os.path.join(os.path.dirname(__file__), "kat_vectors.json"), "w"))
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
