"""Kernel LOGIC parity without a GPU: the unmodified .cu sources compiled for the CPU by the launch emulator
(tests/emu) against the oracle, frame by frame, every image bit-for-bit.  (The real CUDA build is checked by test_gpu_parity.py.)"""
import parity
from kajiya_b200 import scenes


def _clean(report):
    bad = [(f, b) for f, frame in enumerate(report) for b in frame]
    assert not bad, bad[:10]


def test_cornell_lockstep(oracle_lib, emu_lib):
    scene, view = scenes.cornell_box()
    wa, wb, report = parity.run_lockstep(oracle_lib, emu_lib, scene, view, 96, 64, 7)
    _clean(report)
    assert wb.stats()["launches"] == 15 and len(wb.image_names()) >= 40


def test_cornell_odd_extent_single_spatial_pass(oracle_lib, emu_lib):
    scene, view = scenes.cornell_box()
    _, _, report = parity.run_lockstep(oracle_lib, emu_lib, scene, view, 53, 37, 5, spatial_reuse_pass_count=1)
    _clean(report)


def test_atrium_lockstep(oracle_lib, emu_lib):
    scene, view = scenes.atrium(target_tris=8000)
    _, _, report = parity.run_lockstep(oracle_lib, emu_lib, scene, view, 80, 48, 4)
    _clean(report)


def test_reference_path_tracer(oracle_lib, emu_lib):
    import numpy as np
    scene, view = scenes.cornell_box()
    wa, wb = parity.make_world(oracle_lib, scene, 40, 40), parity.make_world(emu_lib, scene, 40, 40)
    for _ in range(2):
        wa.render_reference(**view); wb.render_reference(**view)
    assert np.array_equal(wa.image("refpt.accum").view(np.uint32), wb.image("refpt.accum").view(np.uint32))


def test_gbuffer_ring_replay_and_host_upload_match(emu_lib):
    """The two bench legs feed the hot path differently — device-resident ring (`value`) vs host buffers uploaded inside the
    call (`e2e`) — and must produce identical frames."""
    import numpy as np
    scene, view = scenes.cornell_box()
    wc, wd = parity.make_world(emu_lib, scene, 64, 40), parity.make_world(emu_lib, scene, 64, 40)
    for i in range(3):
        wc.render_frame(capture_slot=i + 1, **view)
        wd.render_frame(**view)
    host = [[np.ascontiguousarray(wc.image(f"slot{i + 1}.{n}")) for n in ("gbuffer", "depth", "geometric_normal", "velocity")] for i in range(3)]
    result = np.zeros((40, 64, 4), np.float16)
    for i in range(3):
        wc.render_frame(replay_slot=i + 1, **view)
        wd.render_frame(host_inputs=tuple(a.ctypes.data for a in host[i]), host_result=result.ctypes.data, **view)
    names = [n for n in wc.image_names() if n.startswith("rtdgi.")]
    assert len(names) > 25 and not parity.compare_images(wc, wd, names)
    assert np.array_equal(result.view(np.uint16), wd.image("rtdgi.spatial_filtered").view(np.uint16))


def test_taa_native_and_upscaled(oracle_lib, emu_lib):
    """T1-T7 (taa.rs:41-185) at native resolution and with 1.5x temporal super-resolution"""
    scene, view = scenes.cornell_box()
    for kw in (dict(enable_taa=True), dict(enable_taa=True, upscale=(150, 96))):
        wa, wb, report = parity.run_lockstep(oracle_lib, emu_lib, scene, view, 100, 64, 5, **kw)
        _clean(report)
        assert "taa.this_frame_out" in wb.image_names() and wb.stats()["launches"] == 22
