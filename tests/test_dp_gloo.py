"""World-size-2 check (gloo, CPU) of the data-parallel gradient rule that
`univer_ocr_b200.parallel.DataParallel` implements:

  * Dice / Jaccard losses SUM over the batch (losses.py:23)  -> shard gradients are summed, scale 1
  * SoftmaxCE divides by the LOCAL batch (losses.py:69-72)   -> summed shard gradients are scaled by 1/world

Each rank computes the float64 oracle gradients of its half of the batch, the ranks all-reduce
(SUM) one flat gradient buffer -- the same collective the GPU path issues through NCCL -- and the
scaled result must equal the full-batch oracle gradient.  The L2 term depends on the weights only
and is added after the collective, so it is not part of the exchanged buffer."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import np_models
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    spec, kind = np_models.net_spec(name), np_models.loss_kind(name)
    w = np_models.golden_weights(name, 11)
    rng = np.random.default_rng(5)
    shape = (4, 16, 16, 1) if name != 'char' else (4, 32, 10, 1)
    X = rng.uniform(size=shape)
    pred = np_models.forward(spec, w, X)
    if kind == 'dice':
        y = (rng.uniform(size=pred.shape) < 0.3).astype(np.float64)
    else:
        y = np.zeros(pred.shape)
        y[np.arange(y.shape[0]), rng.integers(0, y.shape[1], size=y.shape[0])] = 1

    def grads_of(Xs, ys):
        p, saved = np_models.forward(spec, w, Xs, keep=True)
        _, g = np_models.loss_and_grad(kind, p, ys)
        _, gr = np_models.backward(spec, w, saved, g)
        return np.concatenate([gr[k][n].ravel() for k in sorted(gr) for n in sorted(gr[k])])

    full = grads_of(X, y)
    per = shape[0] // world
    rows = pred.shape[0] // shape[0]                       # Char: W windows per image
    xs = X[rank * per:(rank + 1) * per]
    ys = y[rank * per * rows:(rank + 1) * per * rows] if kind != 'dice' else y[rank * per:(rank + 1) * per]
    flat = torch.from_numpy(grads_of(xs, ys).copy())
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    scale = 1.0 if kind == 'dice' else 1.0 / world
    err = float(np.max(np.abs(flat.numpy() * scale - full)) / max(np.max(np.abs(full)), 1e-30))
    if rank == 0:
        out.put((name, err))
    dist.destroy_process_group()


@pytest.mark.parametrize('name', ['monochrome', 'line', 'char'])
def test_two_rank_gradient_allreduce_matches_full_batch(name):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 29600 + (os.getpid() + hash(name)) % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got_name, err = out.get(timeout=10)
    assert got_name == name and err < 1e-10, err
