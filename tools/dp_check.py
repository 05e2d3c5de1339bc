"""Multi-GPU equivalence check (run under torchrun with >= 2 ranks):
two DataParallel steps on batch shards == two single-process steps on the whole batch.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
os.environ['UOCR_DEVICE'] = str(local)
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import univer_ocr_b200.nn as nn
from univer_ocr_b200 import my_model
from univer_ocr_b200.parallel import DataParallel

nn.CP.use_gpu(); nn.CP.set_math_mode('fp32')
ok = True
for name, shape in (('monochrome', (4 * world, 32, 48, 1)), ('line', (2 * world, 32, 64, 1)), ('char', (2 * world, 32, 24, 1))):
    rng = np.random.default_rng(3)                       # same data on every rank
    X = rng.uniform(size=shape).astype(np.float32)
    np.random.seed(7)
    model = my_model.MAKERS[name]((shape[0] // world, *shape[1:]), optimizer=nn.optimizers.Adam(lr=0.002))
    if name == 'char':
        for k, p in model.params().items():
            w = p.value.get(); p.value = (w - w.mean()) * 2.5
    w0 = {k: p.value.get().copy() for k, p in model.params().items()}
    out_rows = model.get_output_shapes([(shape[0] // world, *shape[1:])])[0]
    full_rows = (out_rows[0] * world, *out_rows[1:])
    if name == 'char':
        y = np.zeros(full_rows, dtype=np.float32); y[np.arange(full_rows[0]), rng.integers(0, full_rows[1], full_rows[0])] = 1
    else:
        y = (rng.uniform(size=full_rows) < 0.3).astype(np.float32)
    per_x, per_y = shape[0] // world, full_rows[0] // world
    dp = DataParallel(model)                              # broadcasts rank 0's weights
    for _ in range(2):
        dp.train(X[rank * per_x:(rank + 1) * per_x], y[rank * per_y:(rank + 1) * per_y])
    got = {k: p.value.get() for k, p in model.params().items()}
    if rank == 0:
        dist_backup = (dist.is_initialized,)
        ref = my_model.MAKERS[name](shape, optimizer=nn.optimizers.Adam(lr=0.002))
        for k, p in ref.params().items():
            p.value = w0[k]
        # single-process reference: plain Model.train on the whole batch (no DataParallel wrapper)
        for _ in range(2):
            ref.train(X, y)
        for k, p in ref.params().items():
            want = p.value.get()
            err = np.max(np.abs(got[k] - want)) / max(np.max(np.abs(want)), 1e-30)
            flag = 'ok' if err < 2e-4 else 'MISMATCH'
            ok &= err < 2e-4
            print(f'{name:11s} {k:36s} max rel err {err:.2e} {flag}')
dist.barrier()
if rank == 0:
    print('DP CHECK', 'PASSED' if ok else 'FAILED')
dist.destroy_process_group()
