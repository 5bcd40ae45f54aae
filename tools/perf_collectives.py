"""Latency of the two collectives of the sharded evaluation step at its message sizes (CUDA events, max over ranks):
all-gather of the packed users (B_local x (H + 2) fp32 per rank) and all-to-all of the per-shard lists (B_local x (2 k + 1) fp32 to
every rank), the latter also as an all-gather of everything.  torchrun --nproc-per-node N tools/perf_collectives.py"""
import os, sys, json
import torch, torch.distributed as dist

def main():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    dev = torch.device("cuda", local)
    B, H, k = 1024, 128, 10
    x = torch.randn(B, H + 2, device=dev)
    out = torch.empty(world * B, H + 2, device=dev)
    p = torch.randn(world, B, 2 * k + 1, device=dev)
    q = torch.empty_like(p)
    full = torch.empty(world, world, B, 2 * k + 1, device=dev)
    def ag(): dist.all_gather_into_tensor(out, x)
    def a2a(): dist.all_to_all_single(q, p)
    def a2a_as_ag(): dist.all_gather_into_tensor(full.view(world * world * B, 2 * k + 1), p.view(world * B, 2 * k + 1))
    def both(): ag(); a2a()
    res = {"world": world}
    for name, fn in (("all_gather_users", ag), ("all_to_all_lists", a2a), ("lists_as_all_gather", a2a_as_ag), ("both", both)):
        for _ in range(10): fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        # eager
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 50], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name + "_eager_us"] = round(float(t) * 1e3, 1)
        # graph
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                for _ in range(10): fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(5): g.replay()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 50], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name + "_graph_us"] = round(float(t) * 1e3, 1)
    if rank == 0: print(json.dumps(res), flush=True)
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
