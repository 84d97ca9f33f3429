"""Feasibility probe: torch.distributed NCCL all_reduce captured in a CUDA graph on a side stream (torchrun, 2 GPUs)."""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", init_method="env://")
buf = torch.full((32 << 20,), float(rank + 1), device="cuda")
comm = torch.cuda.Stream()
dist.all_reduce(torch.ones(8, device="cuda"))     # communicator warm-up outside capture
torch.cuda.synchronize()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        x = buf * 1.0
        comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(comm):
            dist.all_reduce(x)
        y = buf * 2.0                              # independent work beside the collective
        torch.cuda.current_stream().wait_stream(comm)
        out = x + y
for i in range(3):
    buf.fill_(float(rank + 1 + i))
    g.replay()
    torch.cuda.synchronize()
    want = sum(r + 1 + i for r in range(world)) + 2.0 * (rank + 1 + i)
    print(f"rank {rank} replay {i}: {out[0].item()} want {want}", flush=True)
    assert abs(out[0].item() - want) < 1e-4
dist.destroy_process_group()
print("ok")
