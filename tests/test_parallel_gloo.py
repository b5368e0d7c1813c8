"""N>1 host logic on CPU: world_size-2 gloo processes exercise sharding, the weight broadcast and
the rank-ordered result gather of deepbinner_b200/parallel.py."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from conftest import MODELS, model_path


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, blob_path, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world_size),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from deepbinner_b200 import parallel
    r, _, w = parallel.init(backend='gloo')
    blob = open(blob_path, 'rb').read() if r == 0 else None
    got = parallel.broadcast_blob(blob)
    n_reads = 11
    lo, hi = parallel.shard_range(n_reads, r, w)
    rows = np.arange(lo, hi, dtype=np.float32)[:, None] * np.ones((1, 13), np.float32)
    allrows = parallel.gather_rows(rows, n_reads)
    slowest = parallel.max_over_ranks(10.0 + r)
    # product-level sharding of `classify --gpus N`: rank 0's file list is broadcast, every rank formats
    # the rows of its contiguous shard, rank 0 receives them in rank order
    files = parallel.broadcast_object(['f%d.fast5' % i for i in range(7)] if r == 0 else None)
    flo, fhi = parallel.shard_range(len(files), r, w)
    text = ''.join('{}\t{}\n'.format(f, i % 3) for i, f in enumerate(files[flo:fhi], start=flo))
    parts = parallel.gather_objects((text, {f: str(i) for i, f in enumerate(files[flo:fhi], start=flo)}))
    everyone = parallel.all_gather_objects(r)
    parallel.barrier()
    out[rank] = (len(got), hash(got), (lo, hi), allrows[:, 0].tolist(), slowest, parts, everyone)


def test_two_rank_broadcast_shard_gather():
    ctx = mp.get_context('spawn')
    out = ctx.Manager().dict()
    port = _free_port()
    path = model_path(MODELS[0])
    procs = [ctx.Process(target=_worker, args=(r, 2, port, path, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    blob = open(path, 'rb').read()
    assert out[0][0] == out[1][0] == len(blob)
    assert out[0][2] == (0, 6) and out[1][2] == (6, 11)
    assert out[0][3] == out[1][3] == [float(i) for i in range(11)]
    assert out[0][4] == out[1][4] == 11.0
    assert out[1][5] is None and out[0][6] == out[1][6] == [0, 1]
    text = ''.join(t for t, _ in out[0][5])
    assert text == ''.join('f{}.fast5\t{}\n'.format(i, i % 3) for i in range(7))      # rank order == input order
    merged = {}
    for _, part in out[0][5]:
        merged.update(part)
    assert merged == {'f%d.fast5' % i: str(i) for i in range(7)}


def test_shard_range_covers_everything():
    from deepbinner_b200.parallel import shard_range
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
