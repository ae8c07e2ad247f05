"""
Multi-GPU plumbing for the FFTLog path: one process per GPU, rows (cosmology x redshift x multipole) split into
contiguous blocks, no data-path collective (every row is independent).  The only optional collective is a
gather of the result shards (``torch.distributed``: NCCL on GPUs, gloo in the CPU tests); it is never inside a throughput
number.
"""

import numpy as np


def shard_bounds(nrows, rank, world):
    """[start, stop) of the contiguous block of ``nrows`` rows owned by ``rank`` (sizes differ by at most one)."""
    if not 0 <= rank < world:
        raise ValueError('rank {} not in [0, {})'.format(rank, world))
    base, extra = divmod(int(nrows), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return None
    return dist


def rank_world():
    dist = _dist()
    if dist is None:
        return 0, 1
    return dist.get_rank(), dist.get_world_size()


def shard(array, rank=None, world=None):
    """Rows of ``array`` (leading axis) owned by this rank."""
    if rank is None or world is None:
        rank, world = rank_world()
    start, stop = shard_bounds(array.shape[0], rank, world)
    return array[start:stop]


def gather_rows(local, nrows_total):
    """
    All-gather row shards produced by :func:`shard` back into the full array (same on every rank).  ``local`` is a
    torch tensor (CUDA with NCCL, CPU with gloo) or a numpy array (converted through a CPU tensor).
    """
    import torch
    dist = _dist()
    is_numpy = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local)) if is_numpy else local.contiguous()
    if dist is None:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(nrows_total, r, world) for r in range(world)]
    if t.shape[0] != sizes[dist.get_rank()][1] - sizes[dist.get_rank()][0]:
        raise ValueError('local shard has {} rows, expected {}'.format(t.shape[0], sizes[dist.get_rank()][1] - sizes[dist.get_rank()][0]))
    most = max(b - a for a, b in sizes)
    if all(b - a == most for a, b in sizes):
        # equal shards (the usual case): one collective straight into the result, no staging copies
        out = t.new_empty((nrows_total,) + tuple(t.shape[1:]))
        dist.all_gather_into_tensor(out, t)
        return out.numpy() if is_numpy else out
    # all_gather wants equal shapes: pad every shard to the largest one (they differ by at most one row), trim after
    if t.shape[0] < most:
        t = torch.cat([t, t.new_zeros((most - t.shape[0],) + tuple(t.shape[1:]))], dim=0)
    parts = [torch.empty_like(t) for _ in sizes]
    dist.all_gather(parts, t)
    out = torch.cat([part[:b - a] for part, (a, b) in zip(parts, sizes)], dim=0)
    return out.numpy() if is_numpy else out


def bind_to_device_numa(device):
    """
    Bind the calling process to the CPU cores (and, when libnuma is present, the memory) of the NUMA node the GPU ``device`` hangs off, so that
    each rank's pinned staging buffers and copy threads sit next to its own PCIe root instead of all on node 0 (eight ranks
    shared one memory controller).  Linux sysfs only; returns a dict describing what was done (for the bench record), never raises.
    """
    import os
    info = {'device': int(device), 'numa_node': None, 'cpus': None, 'membind': False}
    try:
        import torch
        prop = torch.cuda.get_device_properties(device)
        bdf = '{:04x}:{:02x}:{:02x}.0'.format(prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        info['pci'] = bdf
        with open('/sys/bus/pci/devices/{}/numa_node'.format(bdf)) as f:
            node = int(f.read().strip())
        if node < 0:
            return info
        info['numa_node'] = node
        with open('/sys/devices/system/node/node{}/cpulist'.format(node)) as f:
            cpus = set()
            for part in f.read().strip().split(','):
                lo, _, hi = part.partition('-')
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info['cpus'] = len(allowed)
        try:
            import ctypes
            numa = ctypes.CDLL('libnuma.so.1')
            if numa.numa_available() >= 0:
                numa.numa_set_preferred(node)
                info['membind'] = True
        except OSError:
            pass
    except Exception as exc:      # best effort: a container without sysfs, an old torch without the PCI ids, ...
        info['error'] = repr(exc)
    return info
