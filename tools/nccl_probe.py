"""All-gather latency / bandwidth between the GPUs of one box, as torch's NCCL sees it (diagnostic for the tile exchanges).
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 tools/nccl_probe.py"""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); ws = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
for nbytes in (64 << 10, 1 << 20, 4 << 20, 8 << 20, 16 << 20, 64 << 20):
    send = torch.empty(nbytes, dtype=torch.uint8, device="cuda"); recv = torch.empty(nbytes * ws, dtype=torch.uint8, device="cuda")
    for _ in range(5): dist.all_gather_into_tensor(recv, send)
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): dist.all_gather_into_tensor(recv, send)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 20 * 1e3
    if rank == 0: print(f"all_gather {nbytes >> 10:7d} KiB per rank x {ws}: {us:8.1f} us  ({nbytes * (ws - 1) / us / 1e3:7.1f} GB/s received per rank)")
if ws >= 2 and rank == 0:
    x = torch.empty(64 << 20, dtype=torch.uint8, device="cuda:0"); y = torch.empty(64 << 20, dtype=torch.uint8, device="cuda:1")
    for _ in range(3): y.copy_(x)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): y.copy_(x)
    b.record(); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    print(f"cudaMemcpyPeer 64 MiB gpu0 -> gpu1: {64 * 1.048576 / (a.elapsed_time(b) / 10):.1f} GB/s; can_device_access_peer = {torch.cuda.can_device_access_peer(0, 1)}")
dist.barrier(); dist.destroy_process_group()
