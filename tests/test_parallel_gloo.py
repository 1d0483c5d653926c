"""`-m "not gpu"`: the N>1 path (frame-range sharding with halos + final gather) under world_size-2 gloo on CPU, with a
CPU stand-in for the per-rank predictor that has the same window semantics as the patch-wise network."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multipitch_architectures_b200.parallel import HALF, local_window, predict_sharded, shard_range


def window_predictor(h, lo, hi):
    """Depends on exactly the 75-frame window (zero padded beyond the array), like one patch of the network."""
    C, n, F = h.shape
    pad = torch.zeros(C, n + 2 * HALF + 1, F)
    pad[:, HALF:HALF + n] = h
    w = torch.linspace(0.5, 1.5, 75)
    out = []
    for i in range(lo, hi):
        win = pad[:, i:i + 75, :]                                   # frames i-37 .. i+37
        out.append((win * w[None, :, None]).sum(dim=(0, 1))[:72] + win[:, 0, :72].sum(0) * 3 - win[:, 74, :72].sum(0))
    return torch.stack(out) if out else torch.empty(0, 72)


def test_shard_ranges_cover_and_align():
    for n, world, mult in ((1292, 8, 1), (1292, 8, 50), (75, 4, 50), (10, 4, 1), (1, 2, 1), (0, 2, 1), (5073, 3, 50)):
        rs = [shard_range(n, world, r, mult) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert all(s % mult == 0 for s, e in rs if e > s)
        sizes = [e - s for s, e in rs]
        assert max(sizes) - min(sizes) < 2 * mult          # one block of imbalance + truncation of the last block
    assert local_window(100, 40, 60) == (3, 98, 37, 57)
    assert local_window(100, 0, 10) == (0, 48, 0, 10)


def _worker(rank, world, port, n_frames, mult, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    h = torch.rand(6, n_frames, 216, generator=g)
    full = predict_sharded(window_predictor, h, multiple=mult)
    ref = window_predictor(h, 0, n_frames)
    q.put((rank, bool(torch.equal(full, ref)), tuple(full.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_frames,mult', [(130, 1), (130, 50), (40, 1)])
def test_sharded_prediction_equals_unsharded_world2(n_frames, mult):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, mult, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (n_frames, 72) for _, _, shape in res)
