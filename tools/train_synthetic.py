"""`run train` on synthetic data through univer_ocr_b200.trainer.Trainer, single GPU or
data-parallel:

    python tools/train_synthetic.py --epochs 3 --batch 64
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 \\
        tools/train_synthetic.py --epochs 3 --batch 64

Mirrors `my_model/train.py:100-289` for the four single-network modes (TRAIN_MONOCHROME / _PARAGRAPH / _LINE / _CHAR:
lr 0.0015, decay 0.995 / 0.9, one Adam per mode, weights merged into model_weights.json when the validation loss
improves) with seeded synthetic samples of the BASELINE tile shapes instead of the font renderer."""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = {'monochrome': (496, 736), 'paragraph': (496, 736), 'line': (128, 256), 'char': (32, 256)}
MODES = [('monochrome', 0.0015, 0.995), ('paragraph', 0.0015, 0.995), ('line', 0.0015, 0.995), ('char', 0.0015, 0.9)]


class SyntheticDataset:
    """Seeded samples of one network, generated once and kept in HBM as two stacked tensors; `get_batch` gathers the
    requested samples with device-to-device copies, `get` serves single host samples (the reference's protocol)."""

    def __init__(self, name, n, seed, out_shape_of, nn, lib):
        self.name, self.n, self.nn, self.lib = name, n, nn, lib
        h, w = SHAPES[name]
        out = out_shape_of((1, h, w, 1))
        xs, ys = [], []
        for i in range(n):
            rng = np.random.default_rng(seed * 100003 + i)
            x = rng.random((1, h, w, 1), dtype=np.float32)
            if name == 'char':
                y = np.zeros(out, dtype=np.float32)
                y[np.arange(out[0]), rng.integers(1, out[1], size=out[0])] = 1
            elif out[-1] == 1:
                y = (x > 0.7).astype(np.float32)
            else:
                y = np.concatenate([x > 0.7, x < 0.2], axis=-1).astype(np.float32)
            xs.append(x)
            ys.append(y)
        self.x_shape, self.y_shape = xs[0].shape, ys[0].shape       # leading axis: rows per sample
        self.X = nn.CP.copy(np.concatenate(xs, axis=0))
        self.Y = nn.CP.copy(np.concatenate(ys, axis=0))

    def __len__(self):
        return self.n

    def get(self, i):
        X, Y = self.get_batch([i])[self.name]
        return {self.name: (X.get(), Y.get())}

    def _gather(self, src, sample_shape, indices):
        rows = sample_shape[0]
        out = self.nn.DeviceArray((rows * len(indices),) + tuple(sample_shape[1:]))
        per = int(np.prod(sample_shape)) * 4
        for k, i in enumerate(indices):
            self.lib.uocr_memcpy_d2d(out.ptr + k * per, src.ptr + i * per, per, self.nn.CP.stream())
        return out

    def get_batch(self, indices):
        return {self.name: (self._gather(self.X, self.x_shape, indices), self._gather(self.Y, self.y_shape, indices))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--epochs', type=int, default=2)
    ap.add_argument('--batch', type=int, default=64, help='global batch per step')
    ap.add_argument('--train-samples', type=int, default=256)
    ap.add_argument('--val-samples', type=int, default=64)
    ap.add_argument('--modes', default='monochrome,paragraph,line,char')
    ap.add_argument('--weights', default=None, help='model_weights.json to merge into (default: a temp file)')
    args = ap.parse_args()

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    import univer_ocr_b200.nn as nn
    from univer_ocr_b200 import my_model, weights_io
    from univer_ocr_b200._lib import lib
    from univer_ocr_b200.trainer import Trainer
    lib.uocr_set_device(local)
    nn.CP.use_gpu()
    nn.CP.set_math_mode('tf32')
    path = args.weights or os.path.join(tempfile.mkdtemp(), 'model_weights.json')

    for name, lr, lr_step in MODES:
        if name not in args.modes.split(','):
            continue
        h, w = SHAPES[name]
        opt = nn.optimizers.Adam(lr=lr)
        model = my_model.MAKERS[name]((1, h, w, 1), optimizer=opt)
        weights_io.load_weights(model, path) if os.path.exists(path) else None
        out_shape_of = lambda shape, m=model: tuple(m.get_output_shapes([shape])[0])
        train = SyntheticDataset(name, args.train_samples, 1, out_shape_of, nn, lib)
        val = SyntheticDataset(name, args.val_samples, 2, out_shape_of, nn, lib)
        trainer = Trainer({name: model}, train, val, optimizer=opt, learning_rate_step=lr_step, batch_size=args.batch,
                          save_weights_func=lambda names, m=model: weights_io.save_weights(m, path),
                          log=(print if rank == 0 else (lambda *a, **k: None)))
        t0 = time.perf_counter()
        best, best_epoch = trainer.train(args.epochs)
        dt = time.perf_counter() - t0
        if rank == 0:
            n_img = (args.epochs * (args.train_samples + args.val_samples) + args.val_samples)
            print(f'== {name}: best validation loss {best[name]} at epoch {best_epoch[name]}; '
                  f'{n_img / dt:.0f} images/s through the epoch driver (samples resident in HBM), world {world}', flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
