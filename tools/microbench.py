#!/usr/bin/env python
"""Per-layer microbench sweep (BASELINE.json configs[4]): Conv2D fwd/bwd, pooling, upsample,
activations, losses and Adam over my_model's shapes at batch N, timed with CUDA events on the
compute stream (L2 flushed between iterations by a 256 MB memset), reported against the
algorithmic bytes / flops of SURVEY.md 8(d).

    python tools/microbench.py [--batch 64] [--iters 5] [--math fp32|tf32] [--only substr] [--json out]
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
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--math', default='fp32')
    ap.add_argument('--only', default='')
    ap.add_argument('--json', default='')
    ap.add_argument('--no-bwd', action='store_true')
    args = ap.parse_args()

    import univer_ocr_b200.nn as nn
    from univer_ocr_b200 import roofline
    from univer_ocr_b200._lib import lib
    L = nn.layers
    nn.CP.use_gpu()
    nn.CP.set_math_mode(args.math)
    stream = nn.CP.stream()
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}
    flush = nn.DeviceArray((64 * 1024 * 1024,))          # 256 MB > 126 MB L2

    def event():
        e = ctypes.c_void_p()
        lib.uocr_event_create(ctypes.byref(e))
        return e.value

    def timeit(fn):
        fn()
        fn()
        best, samples = 1e30, []
        for _ in range(args.iters):
            flush.fill(0)
            e0, e1 = event(), event()
            lib.uocr_event_record(e0, stream)
            fn()
            lib.uocr_event_record(e1, stream)
            lib.uocr_event_sync(e1)
            ms = ctypes.c_float(0)
            lib.uocr_event_elapsed_ms(e0, e1, ctypes.byref(ms))
            best = min(best, ms.value)
            samples.append(ms.value)
        return float(np.median(samples)), best

    rng = np.random.default_rng(0)
    N = args.batch
    rows = []

    def report(name, ms, best, work):
        gbs = work['bytes'] / (ms / 1e3) / 1e9
        tfs = work['flops'] / (ms / 1e3) / 1e12
        if work['bound'] == 'tensor':
            frac = tfs / (peaks['bf16_tflops'] / 2)
        else:
            frac = gbs / peaks['hbm_gbs']
        rows.append({'name': name, 'ms': round(ms, 4), 'best_ms': round(best, 4), 'bound': work['bound'],
                     'GBps': round(gbs, 1), 'TFLOPs': round(tfs, 2), 'frac': round(frac, 4)})
        print(f'{name:44s} {ms:9.4f} ms  {gbs:8.1f} GB/s  {tfs:7.2f} TFLOP/s  {work["bound"]:6s} frac {frac:.3f}',
              flush=True)

    def randn(shape):
        return nn.CP.copy(rng.standard_normal(shape).astype(np.float32))

    convs = [
        ('mono conv_1 3x3 1->16', (496, 736), 1, 16, (3, 3), 1, 1),
        ('mono conv_2 3x3 16->1', (496, 736), 16, 1, (3, 3), 1, 1),
        ('para down_1 5x5 1->1 s2', (496, 736), 1, 1, (5, 5), 2, 2),
        ('para down_2 5x5 1->1 s2', (248, 368), 1, 1, (5, 5), 2, 2),
        ('para up_2 5x5 1->1', (124, 184), 1, 1, (5, 5), 2, 1),
        ('para up_1 5x5 1->1', (248, 368), 1, 1, (5, 5), 2, 1),
        ('para end 5x5 1->1', (496, 736), 1, 1, (5, 5), 2, 1),
        ('line down_1 5x5 1->4 s2', (128, 256), 1, 4, (5, 5), 2, 2),
        ('line down_2 5x5 4->4 s2', (64, 128), 4, 4, (5, 5), 2, 2),
        ('line up_2 5x5 4->4', (32, 64), 4, 4, (5, 5), 2, 1),
        ('line up_1 5x5 4->4', (64, 128), 4, 4, (5, 5), 2, 1),
        ('line end 5x5 4->2', (128, 256), 4, 2, (5, 5), 2, 1),
        ('char conv_1 5x3 1->64', (32, 256), 1, 64, (5, 3), (0, 1), (2, 1)),
        ('char conv_2 5x3 64->64', (14, 256), 64, 64, (5, 3), (0, 1), (2, 1)),
        ('char conv_3 5x3 64->64', (5, 256), 64, 64, (5, 3), (0, 1), (2, 1)),
    ]
    for name, hw, cin, cout, ks, pad, st in convs:
        if args.only and args.only not in name:
            continue
        layer = L.Convolutional2D(ks, cin, cout, padding=pad, stride=st)
        shape = (N, *hw, cin)
        X = randn(shape)
        state = {}

        def fwd():
            state['y'] = layer.forward(X)[0]
        ms, best = timeit(fwd)
        report(f'{name} fwd', ms, best, roofline.layer_work(layer, shape, 'forward'))
        if args.no_bwd:
            continue
        dy = randn(state['y'].shape)

        def bwd():
            layer._mem[0] = X
            layer._backward(dy, 0)
        ms, best = timeit(bwd)
        report(f'{name} bwd(dgrad+wgrad)', ms, best, roofline.layer_work(layer, shape, 'backward'))
        del X, dy
        state.clear()

    fcs = [('char dense_1 512->1024', 512, 1024), ('char dense_2 1024->128', 1024, 128),
           ('char dense_3 128->162', 128, 162)]
    for name, n_in, n_out in fcs:
        if args.only and args.only not in name:
            continue
        layer = L.FullyConnected(n_in, n_out)
        X = randn((N * 256, n_in))
        state = {}

        def fwd():
            state['y'] = layer.forward(X)[0]
        ms, best = timeit(fwd)
        report(f'{name} fwd', ms, best, roofline.layer_work(layer, X.shape, 'forward'))
        if not args.no_bwd:
            dy = randn(state['y'].shape)

            def bwd():
                layer._mem[0] = X
                layer._backward(dy, 0)
            ms, best = timeit(bwd)
            report(f'{name} bwd', ms, best, roofline.layer_work(layer, X.shape, 'backward'))

    ew = [('leaky_relu (N,496,736,16)', L.LeakyRelu(0.01), (N, 496, 736, 16)),
          ('sigmoid (N,496,736,1)', L.Sigmoid(), (N, 496, 736, 1)),
          ('upsample x2 (N,248,368,1)', L.Upsample2D(2), (N, 248, 368, 1)),
          ('upsample x2 (N,64,128,4)', L.Upsample2D(2), (N, 64, 128, 4)),
          ('maxpool k2 (N,496,736,16)', L.MaxPool2D(2), (N, 496, 736, 16)),
          ('maxpool k3 (N,240,320,6)', L.MaxPool2D(3), (N, 240, 320, 6)),
          ('window8 (N,1,256,64)', L.Conv2DToBatchedFixedWidthed(8), (N, 1, 256, 64))]
    for name, layer, shape in ew:
        if args.only and args.only not in name:
            continue
        X = randn(shape)
        state = {}

        def fwd():
            state['y'] = layer.forward(X)[0]
        ms, best = timeit(fwd)
        report(f'{name} fwd', ms, best, roofline.layer_work(layer, shape, 'forward'))
        if not args.no_bwd:
            dy = randn(state['y'].shape)
            saved = dict(layer._mem)

            def bwd():
                layer._mem = dict(saved)
                layer._backward(dy, 0)
            ms, best = timeit(bwd)
            w_ = roofline.layer_work(layer, shape, 'backward')
            if w_['bytes'] == roofline.layer_work(layer, shape, 'forward')['bytes']:
                w_ = dict(w_)
            report(f'{name} bwd', ms, best, w_)
        del X
        state.clear()

    if not args.only or 'dice' in args.only:
        pred = nn.CP.copy(rng.uniform(0.01, 0.99, size=(N, 496, 736, 1)).astype(np.float32))
        gt = nn.CP.copy((rng.uniform(size=(N, 496, 736, 1)) < 0.2).astype(np.float32))
        loss = nn.losses.SegmentationDice2D()
        ms, best = timeit(lambda: loss(pred, gt)[1].materialize())       # loss + the plain gradient tensor
        report('dice fwd+bwd (N,496,736,1)', ms, best, {'bound': 'hbm', 'bytes': 20 * pred.size, 'flops': 0})
        xpre = nn.CP.copy(rng.standard_normal((N, 496, 736, 1)).astype(np.float32))
        ms, best = timeit(lambda: loss(pred, gt)[1].through_sigmoid(xpre))
        report('dice fwd + (dice, sigmoid) bwd fused', ms, best, {'bound': 'hbm', 'bytes': 24 * pred.size, 'flops': 0})
    if not args.only or 'adam' in args.only:
        n = 803395
        p = L.Param(rng.standard_normal(n).astype(np.float32), optimizer=nn.optimizers.Adam())
        p.grad = randn((n,))
        ms, best = timeit(lambda: p.update_grad())
        report('adam 803395 params', ms, best, {'bound': 'hbm', 'bytes': 28 * n, 'flops': 0})

    if args.json:
        json.dump(rows, open(args.json, 'w'), indent=1)


if __name__ == '__main__':
    main()
