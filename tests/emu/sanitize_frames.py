"""Frames for tests/emu/sanitize.sh: the all-features path (imported glTF scene with triangle lights, 2 spatial passes + restir check, irradiance cache,
reflections + light specular, SSAO, soft sun through the shadow denoiser, lit composite, TAA upsampling) at extents that are not multiples of 8, and the
headline rtdgi path + reference path tracer at an odd extent.  Usage: python sanitize_frames.py <path to a sanitizer build of libkjb_emu>"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from kajiya_b200._abi import KjbLib
from kajiya_b200 import scenes, asset
from kajiya_b200.world import World
lib=KjbLib(sys.argv[1])
path=os.path.join(ROOT, 'tests', 'golden', 'gltf', 'courtyard.gltf')
view=dict(camera_position=(0.5,2.5,7.0), camera_rotation=(float(np.sin(-0.15)),0.0,0.0,float(np.cos(-0.15))), sun_direction=(0.35,0.8,0.45))
kw=dict(spatial_reuse_pass_count=2, use_raytraced_reservoir_visibility=True, enable_ircache=True, enable_rtr=True, enable_ssao=True, enable_lighting=True, enable_taa=True, upscale=(105, 69))
sc=asset.GltfScene(path); w=World(lib, 70, 46, **kw)   # extents that are not multiples of 8: partial denoiser tiles
w.add_instance(w.add_mesh_desc(sc.desc, use_lights=True), np.array([[1,0,0,0],[0,1,0,0],[0,0,1,0]],np.float32))
w.set_blue_noise(scenes.blue_noise()); w.set_spatial_resolve_offsets(scenes.spatial_resolve_offsets()); sc.close()
for f in range(4):
    v=dict(view); px,py,pz=view['camera_position']; v['camera_position']=(px+0.25*np.sin(0.7*f), py+0.05*f, pz-0.1*f)
    w.render_frame(**v)
print("frames ok", w.stats())
# headline path at an odd extent too
scene, view2 = scenes.cornell_box()
import parity
w2=parity.make_world(lib, scene, 53, 37)
for f in range(4): w2.render_frame(**view2)
for f in range(2): w2.render_reference(**view2)
print("cornell ok")
# native-resolution full path at a 16-byte-aligned extent (the tiled TAA / temporal kernels with both footprints), then the multi-GPU cache-request exchange
# between two caches (kjb_pass_ircache_export_requests / _merge_requests)
import ctypes as C
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_edge_cases as T
scene3, view3 = scenes.atrium()
kw3 = dict(enable_ircache=True, enable_rtr=True, enable_taa=True, spatial_reuse_pass_count=2)
wa, wb = parity.make_world(lib, scene3, 64, 40, **kw3), parity.make_world(lib, scene3, 64, 40, **kw3)
vb = dict(view3); px, py, pz = view3['camera_position']; vb['camera_position'] = (px + 0.6, py, pz - 0.4)
for f in range(3): wa.render_frame(**view3); wb.render_frame(**vb)
for w_ in (wa, wb): w_._cache_parity = False      # three frames: the first grid_meta buffer is current
_, Buf, ba = T._cache_bindings(wa); _, _, bb = T._cache_bindings(wb)
class ShareArgs(C.Structure):
    _fields_ = [("ircache", T._CacheBindings), ("block", Buf), ("max_records", C.c_uint32), ("seed", C.c_uint32)]
d = lib.dll
for fn in ("kjb_pass_ircache_export_requests", "kjb_pass_ircache_merge_requests", "kjb_buffer_alloc"): getattr(d, fn).restype = C.c_int
d.kjb_pass_ircache_export_requests.argtypes = [C.c_void_p, C.c_void_p]; d.kjb_pass_ircache_merge_requests.argtypes = [C.c_void_p, C.c_void_p]; d.kjb_buffer_alloc.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
R = 512   # fewer than the live entries: the export must stop at the cap
block = Buf(); assert d.kjb_buffer_alloc(wa.ctx, 16 + R * 32, C.byref(block)) == 0
assert d.kjb_pass_ircache_export_requests(wa.ctx, C.byref(ShareArgs(ba, block, R, 0))) == 0
wa.sync()
assert d.kjb_pass_ircache_merge_requests(wb.ctx, C.byref(ShareArgs(bb, block, R, 3))) == 0
wb.sync(); wb.render_frame(**vb); wb.render_frame(**vb)
print("native full path + cache exchange ok", wb.stats())
