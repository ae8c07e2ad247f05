"""Host-side ceiling of the end-to-end path: every rank copies the bench's per-step bytes (201 MB pinned host -> device and 201 MB device ->
pinned host, on two streams at once, no kernels) STEPS times; prints per-rank and aggregate GB/s each way.  Run alone and under torchrun
with N ranks: when the aggregate stops growing with N, the host (memory system / root complex of the box, not the library) bounds e2e.
    python tools/lab/copy_ceiling.py                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/lab/copy_ceiling.py
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    nbytes, steps = 201326592, 20
    h_in = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
    h_out = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
    h_in.fill_(1.)
    d_in = torch.empty(nbytes // 8, dtype=torch.float64, device='cuda')
    d_out = torch.ones(nbytes // 8, dtype=torch.float64, device='cuda')
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ('h2d', 'd2h', 'both'):
        def step():
            if mode in ('h2d', 'both'):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ('d2h', 'both'):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        s1.synchronize(); s2.synchronize()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = {'ms_per_step_max_over_ranks': float(t.item()), 'GBps_each_way_per_gpu': nbytes / (float(t.item()) * 1e-3) / 1e9,
                     'GBps_each_way_aggregate': world * nbytes / (float(t.item()) * 1e-3) / 1e9}
    if rank == 0:
        print(json.dumps({'copy_ceiling': True, 'n_gpus': world, 'bytes_each_way_per_step_per_gpu': nbytes, 'steps': steps,
                          'cpus_visible': len(os.sched_getaffinity(0)), 'modes': res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
