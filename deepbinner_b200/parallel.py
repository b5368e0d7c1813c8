"""
Multi-GPU plumbing: one process per GPU, reads sharded in contiguous blocks, the packed weight blob
broadcast once from rank 0 (NCCL on GPUs, gloo in the CPU tests), no cross-rank reduction on the
data path.  The reference has no distributed code at all (single process,
classify.py:416-423 only sizes TensorFlow's CPU thread pools); reads are independent, so this is
data parallelism over reads.
"""

import os

import numpy as np
import torch
import torch.distributed as dist


def world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process if unset)."""
    return (int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)),
            int(os.environ.get('WORLD_SIZE', 1)))


def init(backend=None):
    rank, local_rank, world_size = world()
    if world_size > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29512')
        if backend is None:
            # NCCL needs one GPU per rank; ranks that share a GPU (tests on a one-GPU box) or have none
            # use gloo for the (tiny) control traffic
            backend = 'nccl' if torch.cuda.is_available() and torch.cuda.device_count() >= world_size else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world_size)
    return rank, local_rank, world_size


def device_for(local_rank):
    """CUDA device ordinal of a rank: its own GPU, or (fewer GPUs than ranks) local_rank modulo the
    device count."""
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return local_rank % n if n else local_rank


def _collective_device():
    return torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' \
        else torch.device('cpu')


def is_distributed():
    return dist.is_initialized() and dist.get_world_size() > 1


def broadcast_object(obj, src=0):
    """Broadcast a small picklable object (file list of a round) from `src`."""
    if not is_distributed():
        return obj
    box = [obj if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src, device=_collective_device())
    return box[0]


def gather_objects(obj, dst=0):
    """Rank-ordered list of every rank's object on `dst` (None elsewhere): the host-side result
    collection - each rank's rows stay its own slice, nothing is reduced."""
    if not is_distributed():
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out if dist.get_rank() == dst else None


def all_gather_objects(obj):
    if not is_distributed():
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def bind_to_numa_node_of_gpu(device):
    """Pin this process (its fast5-parsing and gather threads, and the pinned staging it allocates
    afterwards) to the CPUs of the NUMA node its GPU hangs off, so that eight ranks do not all stage
    through one node.  Best effort: silently does nothing without sysfs / on a single-node box."""
    try:
        props = torch.cuda.get_device_properties(device)
        bus = '{:04x}:{:02x}:{:02x}.0'.format(props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        node = int(open('/sys/bus/pci/devices/{}/numa_node'.format(bus)).read())
        if node < 0:
            return None
        cpus = set()
        for part in open('/sys/devices/system/node/node{}/cpulist'.format(node)).read().strip().split(','):
            a, _, b = part.partition('-')
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:  # noqa: BLE001
        return None


def shard_range(n_items, rank, world_size):
    """Contiguous block of items for `rank`: [rank*ceil(n/W), (rank+1)*ceil(n/W)) clipped to n."""
    per = -(-n_items // world_size)
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def broadcast_blob(blob, device=None):
    """Broadcast the packed weight blob (bytes) from rank 0; every rank returns identical bytes.
    `blob` may be None on ranks != 0."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return bytes(blob)
    dev = device if device is not None else \
        (torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl'
         else torch.device('cpu'))
    size = torch.tensor([len(blob) if dist.get_rank() == 0 else 0], dtype=torch.int64, device=dev)
    dist.broadcast(size, src=0)
    n = int(size.item())
    if dist.get_rank() == 0:
        buf = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    else:
        buf = torch.empty(n, dtype=torch.uint8, device=dev)
    dist.broadcast(buf, src=0)
    return buf.cpu().numpy().tobytes()


def gather_rows(local_rows, n_total):
    """Concatenate per-rank result rows (numpy [n_r, C]) in rank order on every rank - the host-side
    result collection of section 8e (each rank's rows go to its own slice; no reduction)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_rows
    world_size = dist.get_world_size()
    per = -(-n_total // world_size)
    cols = local_rows.shape[1:]
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' \
        else torch.device('cpu')
    padded = np.zeros((per,) + cols, dtype=local_rows.dtype)
    padded[:len(local_rows)] = local_rows
    mine = torch.from_numpy(padded).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world_size)]
    dist.all_gather(parts, mine)
    return torch.cat(parts).cpu().numpy()[:n_total]


def max_over_ranks(value):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' \
        else torch.device('cpu')
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
