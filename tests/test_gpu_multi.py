"""Tile-sharded frames on real GPUs over NCCL (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` — plain pytest, NOT under
torchrun: the test spawns its own ranks, and torchrun's agent-store environment would send their rendezvous to a store that does not exist).
Each rank's band must equal, bit for bit, the same rows of a single-GPU full-frame render done by the same process."""
import ctypes as C, os, socket, sys
import numpy as np, pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
W, H, FRAMES = 640, 720, 6


def _worker(rank, world_size, port, ret):
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ.pop("TORCHELASTIC_USE_AGENT_STORE", None)     # started from inside a torchrun worker: rendezvous through OUR store, not the agent's
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=torch.device("cuda", rank))
    import kajiya_b200
    from kajiya_b200 import scenes
    import parity
    lib = kajiya_b200.lib()
    scene, view = scenes.cornell_box()
    tiled = parity.make_world(lib, scene, W, H, device=rank, tile=(rank, world_size), enable_taa=True)
    uid = [None]
    if rank == 0:
        buf = C.create_string_buffer(128); assert lib.dll.kjb_comm_nccl_unique_id(buf) == 0; uid[0] = buf.raw
    dist.broadcast_object_list(uid, src=0)
    tiled.comm_init_nccl(uid[0], rank, world_size)
    full = parity.make_world(lib, scene, W, H, device=rank, enable_taa=True)
    for _ in range(FRAMES):
        tiled.render_frame(**view); full.render_frame(**view)
    hh = (H + 1) // 2; y0, y1 = hh * rank // world_size, hh * (rank + 1) // world_size
    bad = []
    for n in ["rtdgi.spatial_filtered", "taa.this_frame_out"] + [n for n in full.image_names() if n.endswith(":0") or n.endswith(":1")]:
        a, b = tiled.image(n), full.image(n); s = a.shape[0] // hh
        if not np.array_equal(a[y0 * s:y1 * s].view(np.uint8), b[y0 * s:y1 * s].view(np.uint8)):
            bad.append(n)
    ret[rank] = bad
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_tile_sharded_frames_over_nccl():
    import torch.multiprocessing as mp
    n = min(torch.cuda.device_count(), 4)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(n, port, ret), nprocs=n, join=True)
    for rank in range(n):
        assert ret[rank] == [], f"rank {rank}: {ret[rank]}"
