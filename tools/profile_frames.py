"""Render a few frames of a bench workload on cuda:0 — the target for `ncu` captures (profiles/README.md).
   python tools/profile_frames.py [--workload cornell_1080p_rtdgi_1s1t] [--frames 6] [--ircache] [--taa]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kajiya_b200
from kajiya_b200.world import World
from kajiya_b200 import scenes

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="cornell_box"); ap.add_argument("--width", type=int, default=1920); ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--frames", type=int, default=6); ap.add_argument("--spatial", type=int, default=1)
ap.add_argument("--ircache", action="store_true"); ap.add_argument("--taa", action="store_true"); ap.add_argument("--rtr", action="store_true")
a = ap.parse_args()
scene, view = getattr(scenes, a.scene)()
w = World(kajiya_b200.lib(), a.width, a.height, spatial_reuse_pass_count=a.spatial, enable_ircache=a.ircache, enable_taa=a.taa, enable_rtr=a.rtr)
scenes.populate(w, scene)
for f in range(a.frames):
    w.render_frame(**view)
w.sync() if hasattr(w, "sync") else None
print("frames", a.frames, w.stats())
