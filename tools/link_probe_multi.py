"""Host-side ceiling of the end-to-end path at N GPUs of one box: every rank copies 1 GiB pinned -> device (and, in the
second leg, 0.55 GiB device -> pinned at the same time), all ranks together; aggregate GB/s of input = N GiB / slowest rank.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/link_probe_multi.py"""
import os, time, json
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30; m = int(n * 0.55)
src = torch.empty(n, dtype=torch.uint8).pin_memory(); src.fill_(7)
out = torch.empty(m, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(m, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for leg in ("h2d", "h2d+d2h"):
    best = 1e9
    for rep in range(5):
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t = time.perf_counter()
        with torch.cuda.stream(s1): d.copy_(src, non_blocking=True)
        if leg != "h2d":
            with torch.cuda.stream(s2): out.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        el = torch.tensor([time.perf_counter() - t], device="cuda")
        if world > 1: dist.all_reduce(el, op=dist.ReduceOp.MAX)
        best = min(best, float(el.item()))
    res[leg] = {"ms": round(best * 1e3, 2), "aggregate_input_gbs": round(world * n / best / 1e9, 1)}
if rank == 0: print(json.dumps({"n_gpus": world, "bytes_in_per_gpu": n, "bytes_out_per_gpu": m, **res}))
if world > 1: dist.destroy_process_group()
