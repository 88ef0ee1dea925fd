"""bench.py pieces that can be checked without a GPU (the driver runs the script itself on the B200)."""
import importlib.util, os
import conftest


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(conftest.ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def test_pass_byte_table_covers_every_label_the_frame_driver_emits():
    b = _bench()
    src = open(os.path.join(conftest.ROOT, "kajiya_b200", "csrc", "host", "kjb_world.cpp")).read()
    import re
    labels = set(re.findall(r'RUN(?:_TOP)?\("([^"]+)"', src)) - {"raster simple", "sky cube", "convolve sky", "brdf fg lut", "reference pt"}
    missing = sorted(l for l in labels if l not in b.PASS_BYTES and not l.startswith("_") and l != "tile border all-gather")
    assert missing == [], missing
    F, Hh = 1920 * 1080, 960 * 540
    for l in labels & set(b.PASS_BYTES):
        assert b.pass_bytes(l, F, Hh) >= 0


def test_issue_slot_roofline_helper():
    b = _bench()
    r = b.issue_slot_roofline("restir spatial", 0.134, {"sm_mhz": 1965, "sm_max_mhz": 1965})
    assert abs(r["peak_ginst_s"] - 148 * 4 * 1.965) < 1e-6 and 0.4 < r["frac"] < 0.8
    assert b.issue_slot_roofline("restir spatial", 0.134, None)["frac"] == r["frac"]      # no clock sample: nominal clock
    assert b.issue_slot_roofline("unknown pass", 0.1, {}) is None and b.issue_slot_roofline("restir spatial", 0.0, {}) is None
