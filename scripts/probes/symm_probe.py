"""Does torch symmetric memory work on this box?  torchrun --nproc-per-node 2 scripts/probes/symm_probe.py"""
import os, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
t = symm.empty(1024, dtype=torch.float32, device="cuda")
t.fill_(rank)
hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "handle", type(hdl).__name__, [n for n in dir(hdl) if not n.startswith("_")], flush=True)
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
hdl.barrier()
peer[:8].copy_(torch.full((8,), 100.0 + rank, device="cuda"))
hdl.barrier()
torch.cuda.synchronize()
print(rank, "my buffer after peer store:", t[:10].tolist(), flush=True)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs], flush=True)
# barrier latency
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    hdl.barrier()
e1.record(); torch.cuda.synchronize()
print(rank, "barrier us", e0.elapsed_time(e1) * 10, flush=True)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        hdl.barrier()
    g.replay(); torch.cuda.synchronize()
    print(rank, "barrier is graph-capturable", flush=True)
except Exception as exc:
    print(rank, "graph capture failed:", repr(exc)[:200], flush=True)
dist.destroy_process_group()
