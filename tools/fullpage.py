#!/usr/bin/env python
"""BASELINE configs[3]: full-page inference, 2048 x 2048 synthetic documents (-> (1, 2064, 2064, 1) after
make_divisible_by), Monochrome -> Paragraph forward, pages sharded round-robin over the ranks (no collective).

    python tools/fullpage.py [--pages 64] [--batch 4] [--math tf32]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/fullpage.py --pages 64
Prints pages/s (device time, max over ranks).
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pages', type=int, default=64)
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--math', default='tf32')
    ap.add_argument('--repeat', type=int, default=3)
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    import univer_ocr_b200.nn as nn
    from univer_ocr_b200 import my_model
    from univer_ocr_b200._lib import lib
    nn.CP.use_gpu()
    nn.CP.set_math_mode(args.math)
    rng = np.random.default_rng(1234 + rank)
    mine = len(range(rank, args.pages, world))                 # round-robin shard
    shape = (args.batch, 2064, 2064, 1)
    mono, para = my_model.make_monochrome(shape), my_model.make_paragraph(shape)
    X = nn.CP.copy(rng.uniform(size=shape).astype(np.float32))
    stream = nn.CP.stream()

    def event():
        e = ctypes.c_void_p()
        lib.uocr_event_create(ctypes.byref(e))
        return e.value

    def run_shard():
        done = 0
        while done < mine:
            para.predict(mono.predict(X)[0])
            done += args.batch
    run_shard()
    best = 1e30
    for _ in range(args.repeat):
        e0, e1 = event(), event()
        lib.uocr_event_record(e0, stream)
        run_shard()
        lib.uocr_event_record(e1, stream)
        lib.uocr_event_sync(e1)
        ms = ctypes.c_float(0)
        lib.uocr_event_elapsed_ms(e0, e1, ctypes.byref(ms))
        best = min(best, ms.value)
    if world > 1:
        from univer_ocr_b200 import comm as comm_
        best = comm_.init_from_env().allreduce_host([best], 'max')[0]
    if rank == 0:
        pages = -(-mine // args.batch) * args.batch * world
        print(json.dumps({'config': 'full-page 2064x2064 Monochrome->Paragraph, pages sharded round-robin',
                          'n_gpus': world, 'pages': pages, 'batch_per_launch': args.batch, 'ms': best,
                          'pages_per_s': pages / (best / 1e3), 'math': args.math}))


if __name__ == '__main__':
    main()
