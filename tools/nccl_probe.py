"""2-GPU diagnostic: latency of the per-step all-gather of log-posteriors (32 KiB per rank)."""
import os
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
x = torch.randn(4096, dtype=torch.float64, device="cuda")
out = torch.empty(world * 4096, dtype=torch.float64, device="cuda")
for _ in range(20):
    dist.all_gather_into_tensor(out, x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    dist.all_gather_into_tensor(out, x)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"all_gather 32KiB x{world}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per call (back to back)")
# interleaved with a compute kernel and an L2 flush, as in bench.py
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
y = torch.randn(4096, 4096, device="cuda")
times = []
for k in range(30):
    flush.fill_(float(k))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    z = y @ y
    dist.all_gather_into_tensor(out, x)
    e.record()
    torch.cuda.synchronize()
    times.append(s.elapsed_time(e))
if rank == 0:
    print("matmul + all_gather per step (ms):", [round(t, 3) for t in times[5:15]])
dist.destroy_process_group()
