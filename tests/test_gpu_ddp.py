"""`-m gpu`, needs >= 2 GPUs (skipped otherwise): data-parallel `loop.fit` — rank-partitioned patch sampling, parameter broadcast, one flat
NCCL gradient all-reduce per step, rank-averaged validation loss — keeps the replicas bit-identical and trains."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from multipitch_architectures_b200.loop import fit
    from tests.refshapes import build_model
    from tests.test_gpu_ext import _toy_sets
    torch.manual_seed(100 + rank)                            # different initial weights per rank: fit() must broadcast rank 0's
    m = build_model('cnn_xs', precision='bf16').cuda()
    aug = {'aug:randomeq': 20, 'aug:tuning': True, 'aug:transpsemitones': 5}
    hist = fit(m, _toy_sets(3, 150, 1, aug), _toy_sets(1, 100, 2, {}), batch_size=16, val_batch_size=25, lr=2e-3, max_epochs=3,
               max_batches_per_epoch=8, seed=5, graph=True, scheduler=False, early=False, log=lambda s: None)
    flat = torch.cat([p.detach().reshape(-1).float() for p in m.parameters()])
    q.put((rank, [h['train_loss'] for h in hist], [h['val_loss'] for h in hist], float(flat.double().sum().item()), float(flat.abs().max().item())))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_fit_keeps_replicas_identical():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, tl0, vl0, sum0, mx0), (_, tl1, vl1, sum1, mx1) = res
    assert tl0 == tl1 and vl0 == vl1                        # rank-averaged losses
    assert sum0 == sum1 and mx0 == mx1                      # replicas hold bit-identical parameters after training
    assert tl0[-1] < tl0[0] and all(np.isfinite(tl0 + vl0))


def _unet_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from multipitch_architectures_b200.libdl import nn_models as M
    from multipitch_architectures_b200.training_unet import UnetTrainStep
    from tests.weights import fill_state_dict, synth_patches, synth_targets
    out = {}
    for overlap in (True, False):
        m = M.simple_u_net_doubleselfattn(n_chan_input=6, n_chan_layers=[16, 10, 8, 5], n_bins_in=216, n_bins_out=72, scalefac=8, embed_dim=64,
                                          num_heads=8, mlp_dim=128, pos_encoding='sinusoidal', precision='bf16')
        m.load_state_dict(fill_state_dict(m.state_dict(), 21, scheme='torch_default'))
        m = m.cuda().train()
        step = UnetTrainStep(m, lr=1e-3, graph=True, overlap_comm=overlap)
        assert 0 < step.n_trunk < step.flat_g.numel()
        losses = []
        for i in range(6):                                   # steps 1-2 eager, 3-6 replay the captured graph pair
            x, t = synth_patches(4, 300 + 10 * i + rank).cuda(), synth_targets(4, 300 + 10 * i + rank).cuda()
            losses.append(float(step(x, t).item()))
        flat = step.flat_p.double()
        out[overlap] = (losses, float(flat.sum().item()), float(flat.abs().max().item()), step._graph2 is not None)
        step.release()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_gradient_all_reduce_keeps_replicas_identical():
    """UnetTrainStep on 2 GPUs: the gradients behind the encoder trunk are all-reduced on NCCL's stream while the trunk's backward (a second
    CUDA graph) runs.  Replicas stay bit-identical, and the run follows the single flat all-reduce."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_unet_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for overlap in (True, False):
        (l0, s0, m0, g0), (l1, s1, m1, g1) = res[0][overlap], res[1][overlap]
        assert s0 == s1 and m0 == m1                         # identical parameters on both ranks
        assert g0 == g1 == overlap                           # the second graph exists exactly in the overlapped mode
        assert all(np.isfinite(l0 + l1))
    a, b = np.array(res[0][True][0]), np.array(res[0][False][0])
    assert np.abs(a - b).max() < 2e-2 * np.abs(b).max()      # same arithmetic up to the summation order of the atomics
