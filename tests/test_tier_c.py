"""Tier C: converged real-time GI against the reference path tracer on the same scene and camera.
kajiya's real-time path is biased by design (reservoir M clamps, the irradiance cache's self-lighting limiter, half-res resolve),
so this is a coarse energy/shape check with a stated tolerance — the bit-level gates are the lockstep tests."""
import numpy as np, pytest
import parity
from kajiya_b200 import scenes

W, H = 112, 72


def _converged(lib, ircache, frames=56, tail=24):
    scene, view = scenes.cornell_box()
    w = parity.make_world(lib, scene, W, H, enable_ircache=ircache)
    acc = np.zeros((H, W, 3)); n = 0
    for f in range(frames):
        w.render_frame(**view)
        if f >= frames - tail:
            acc += w.image("rtdgi.spatial_filtered")[..., :3].astype(np.float64); n += 1
    return acc / n, w.image("depth")[..., 0].copy()


def _path_traced(lib, frames=192):
    scene, view = scenes.cornell_box()
    w = parity.make_world(lib, scene, W, H)
    for _ in range(frames):
        w.render_reference(indirect_only=True, **view)
    return w.image("refpt.accum")[..., :3].astype(np.float64)


def _check(gi0, gi1, pt, depth):
    m = (depth > 0) & (pt.max(-1) < 5.0)          # geometry, minus the emitter itself (the path tracer adds its emission at the primary hit)
    r0, r1 = gi0[m].mean() / pt[m].mean(), gi1[m].mean() / pt[m].mean()
    l2 = np.sqrt(((gi1[m] - pt[m]) ** 2).mean()) / np.sqrt((pt[m] ** 2).mean())
    # stated tolerances: with the cache the mean irradiance is within 30 % of the path tracer's multi-bounce result and the relative
    # per-pixel L2 error below 0.4; without it (single bounce + screen-space feedback only) markedly more energy is missing
    assert 0.70 < r1 < 1.10, (r0, r1)
    assert r1 > r0 + 0.12, (r0, r1)
    assert l2 < 0.40, l2


def test_converged_gi_vs_reference_path_tracer_oracle(oracle_lib):
    pt = _path_traced(oracle_lib)
    gi0, depth = _converged(oracle_lib, False)
    gi1, _ = _converged(oracle_lib, True)
    _check(gi0, gi1, pt, depth)


@pytest.mark.gpu
def test_converged_gi_vs_reference_path_tracer_cuda(cuda_lib):
    pt = _path_traced(cuda_lib)
    gi0, depth = _converged(cuda_lib, False)
    gi1, _ = _converged(cuda_lib, True)     # parallel (racy) cache schedule
    _check(gi0, gi1, pt, depth)


def _lit(lib, frames=52, tail=20, **kw):
    scene, view = scenes.cornell_box()
    w = parity.make_world(lib, scene, W, H, enable_lighting=True, hard_sun=True, **kw)
    acc = np.zeros((H, W, 3)); n = 0
    for f in range(frames):
        w.render_frame(**view)
        if f >= frames - tail:
            acc += w.image("debug_out")[..., :3].astype(np.float64); n += 1
    return acc / n, w.image("depth")[..., 0].copy()


def _check_lit(lib):
    """The complete lit image (direct sun + emissive + rtdgi * albedo + rtr * FG, light_gbuffer.hlsl) against the reference path
    tracer on the same scene/camera, hard sun (the configuration without a shadow denoiser upstream).  Stated tolerance: mean radiance
    within 12 %, relative per-pixel L2 <= 0.15 with the full path; without cache and reflections the error is about twice that."""
    scene, view = scenes.cornell_box()
    wp = parity.make_world(lib, scene, W, H, hard_sun=True)
    for _ in range(192):
        wp.render_reference(**view)
    pt = wp.image("refpt.accum")[..., :3].astype(np.float64)
    lit0, depth = _lit(lib)
    lit1, _ = _lit(lib, enable_ircache=True, enable_rtr=True)
    m = (depth > 0) & (pt.max(-1) < 5.0)
    def err(img): return np.sqrt(((img[m] - pt[m]) ** 2).mean()) / np.sqrt((pt[m] ** 2).mean())
    r1 = lit1[m].mean() / pt[m].mean()
    assert 0.88 < r1 < 1.12, r1
    assert err(lit1) <= 0.15, err(lit1)
    assert err(lit0) > err(lit1) + 0.03, (err(lit0), err(lit1))


def test_lit_image_vs_reference_path_tracer_oracle(oracle_lib):
    _check_lit(oracle_lib)


@pytest.mark.gpu
def test_lit_image_vs_reference_path_tracer_cuda(cuda_lib):
    _check_lit(cuda_lib)


def _check_lit_soft_sun(lib):
    """The default area sun: 1-spp shadow mask -> shadow denoiser -> light_gbuffer, against the path tracer sampling the same sun disk.
    Measured on the oracle: mean ratio 0.93, relative per-pixel L2 0.116; gate: mean within 12 %, L2 <= 0.15."""
    scene, view = scenes.cornell_box()
    wp = parity.make_world(lib, scene, W, H)
    for _ in range(192):
        wp.render_reference(**view)
    pt = wp.image("refpt.accum")[..., :3].astype(np.float64)
    w = parity.make_world(lib, scene, W, H, enable_lighting=True, enable_ircache=True, enable_rtr=True)
    acc = np.zeros((H, W, 3)); n = 0
    for f in range(52):
        w.render_frame(**view)
        if f >= 32:
            acc += w.image("debug_out")[..., :3].astype(np.float64); n += 1
    lit = acc / n; depth = w.image("depth")[..., 0]
    m = (depth > 0) & (pt.max(-1) < 5.0)
    r = lit[m].mean() / pt[m].mean()
    err = np.sqrt(((lit[m] - pt[m]) ** 2).mean()) / np.sqrt((pt[m] ** 2).mean())
    assert 0.88 < r < 1.12, r
    assert err <= 0.15, err
    assert "shadow_denoise.spatial_input" in w.image_names()


def test_soft_sun_lit_image_vs_reference_path_tracer_oracle(oracle_lib):
    _check_lit_soft_sun(oracle_lib)


@pytest.mark.gpu
def test_soft_sun_lit_image_vs_reference_path_tracer_cuda(cuda_lib):
    _check_lit_soft_sun(cuda_lib)
