"""Tiled-vs-untiled parity of one workload on N ranks with details (diagnostic): torchrun ... tools/parity_probe.py <workload> <frames>"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench, kajiya_b200
rank = int(os.environ["RANK"]); ws = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
lib = kajiya_b200.lib()
wl = sys.argv[1]; frames = int(sys.argv[2])
orig = bench.WORKLOADS[wl]
flags = dict(orig[5]); flags.pop("enable_ircache", None) if os.environ.get("PROBE_FORCE_FRAMES") else None
import types
def pc(frames):
    f = bench.parity_check.__wrapped__ if hasattr(bench.parity_check, "__wrapped__") else bench.parity_check
    return f(lib, torch, dist, wl, rank, ws, lr, frames=frames)
# parity_check overrides `frames` for cache workloads: call the pieces directly
import ctypes as C
wt, view, W, H = bench.build_world(lib, wl, device=lr, tile=(rank, ws))
uid = [None]
if rank == 0:
    buf = C.create_string_buffer(128); assert lib.dll.kjb_comm_nccl_unique_id(buf) == 0; uid[0] = buf.raw
dist.broadcast_object_list(uid, src=0)
wt.comm_init_nccl(uid[0], rank, ws)
wu, _, _, _ = bench.build_world(lib, wl, device=lr, tile=None)
name = bench.result_image_name(wl)
for f in range(frames):
    wt.render_frame(**view); wu.render_frame(**view)
    if f + 1 in (6, 12, 24, 48):
        ok, exact, rm, rr, sha = bench.band_compare(wu.image(name), wt.image(name), H, rank, ws, True)
        print(json.dumps({"rank": rank, "frames": f + 1, "rel_mean": round(rm, 4), "rel_rms": round(rr, 4), **bench.band_compare.last_detail}), flush=True)
wt.close(); wu.close()
dist.barrier(); dist.destroy_process_group()
