"""Probe torch symmetric memory on this box: peer pointers, multicast (NVLS) support, barrier latency."""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty((1 << 20,), dtype=torch.float32, device=dev)
h = symm_mem.rendezvous(t, dist.group.WORLD)
info = {"rank": rank, "multicast_ptr": int(h.multicast_ptr), "buffer_ptrs": [hex(p) for p in h.buffer_ptrs][:3], "signal_pad_size": h.signal_pad_size}
try:
    info["has_multicast_support"] = bool(type(h).has_multicast_support(torch._C._autograd.DeviceType.CUDA if False else dev.type and 1, local)) if False else "n/a"
except Exception as e:
    info["has_multicast_support"] = repr(e)
for _ in range(5):
    h.barrier(channel=0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    h.barrier(channel=0)
e1.record(); torch.cuda.synchronize()
info["barrier_us"] = e0.elapsed_time(e1) / 50 * 1e3
x = torch.ones(1, device=dev)
for _ in range(5):
    dist.all_reduce(x)
torch.cuda.synchronize()
e0.record()
for _ in range(50):
    dist.all_reduce(x)
e1.record(); torch.cuda.synchronize()
info["nccl_tiny_allreduce_us"] = e0.elapsed_time(e1) / 50 * 1e3
print("SYMM", info, flush=True)
dist.destroy_process_group()
