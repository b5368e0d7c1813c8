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
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world_size)
    return rank, local_rank, world_size


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
