"""Scenarios that drive BOTH the reference's epoch driver (`my_model/trainer.py`) and
`univer_ocr_b200.trainer.Trainer` with the same scripted scalar "networks", so that their decisions
(sample order, learning-rate decay, NaN roll-backs, best-weights reports, returned losses) can be
compared event by event.  Used by tests/golden/make_trainer_golden.py and tests/test_trainer.py."""
import math
import random

import numpy as np


class ScriptedOptimizer:
    def __init__(self, lr):
        self.lr = lr


class ScriptedModel:
    """w <- w - lr * dL/dw on L = sum_k (w - x)^2 * (k + 1); turns NaN at scripted train steps."""

    def __init__(self, name, w0, optimizer, outputs=1, nan_steps=(), trace=None):
        self.name, self.w, self.optimizer = name, float(w0), optimizer
        self.outputs, self.nan_steps, self.trace = outputs, set(nan_steps), trace
        self.steps = 0

    def get_outputs_count(self):
        return self.outputs

    def _losses(self, X):
        x = float(np.mean(X))
        return x, [(self.w - x) ** 2 * (k + 1) for k in range(self.outputs)]

    def train(self, X, y):
        self.steps += 1
        x, losses = self._losses(X)
        self.w -= self.optimizer.lr * 2.0 * (self.w - x) * sum(k + 1 for k in range(self.outputs))
        if self.steps in self.nan_steps:
            self.w = float('nan')
        if self.trace is not None:
            self.trace.append(['train', self.name, round(x, 9)])
        return {'output_losses': losses, 'regularization_loss': 0}

    def test(self, X, y):
        x, losses = self._losses(X)
        if self.trace is not None:
            self.trace.append(['test', self.name, round(x, 9)])
        return {'output_losses': losses}

    def get_weights(self):
        return {self.name: {'w': [self.w]}}

    def set_weights(self, weights):
        mine = weights.get(self.name)
        if mine is None or math.isnan(mine['w'][0]):      # the reference skips NaN tensors
            return
        if self.trace is not None:
            self.trace.append(['set_weights', self.name, mine['w'][0]])
        self.w = mine['w'][0]

    def nan_weights(self):
        return math.isnan(self.w)


class ScriptedDataset:
    def __init__(self, names, n, seed):
        rng = np.random.default_rng(seed)
        self.samples = [{name: (rng.uniform(size=(1, 2, 2, 1)), rng.uniform(size=(1, 1)))
                         for name in names} for _ in range(n)]

    def __len__(self):
        return len(self.samples)

    def get(self, i):
        return self.samples[i]


SCENARIOS = {
    # name: (models [(name, w0, outputs, nan_steps)], n_train, n_val, epochs, lr, lr_step)
    'plain': ([('mono', 0.9, 1, ()), ('line', -0.4, 2, ())], 5, 2, 4, 0.05, 0.9),
    'nan_once': ([('mono', 0.9, 1, (7,)), ('line', -0.4, 2, ())], 4, 2, 4, 0.05, 0.9),
    # NaN in every step of 11 consecutive epoch attempts: 9 roll-backs to the last weights, then the
    # "too many attempts" branch (start weights, counter reset), one more, then healthy again
    'nan_storm': ([('char', 0.3, 1, tuple(range(4, 4 + 3 * 11)))], 3, 1, 3, 0.1, 0.8),
    'val_nan_best': ([('a', 0.2, 1, ()), ('b', 0.7, 1, (2, 3))], 2, 2, 3, 0.02, 0.95),
}


def build(scenario):
    specs, n_train, n_val, epochs, lr, lr_step = SCENARIOS[scenario]
    trace = []
    opt = ScriptedOptimizer(lr)
    models = {name: ScriptedModel(name, w0, opt, outputs, nan_steps, trace)
              for name, w0, outputs, nan_steps in specs}
    names = list(models)
    return models, opt, ScriptedDataset(names, n_train, 1), ScriptedDataset(names, n_val, 2), epochs, lr_step, trace


def _clean(x):
    """JSON-able, NaN-safe."""
    if isinstance(x, dict):
        return {k: _clean(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_clean(v) for v in x]
    if isinstance(x, float) and math.isnan(x):
        return 'nan'
    if isinstance(x, float) and math.isinf(x):
        return 'inf' if x > 0 else '-inf'
    return x


def run(scenario, trainer_factory):
    """`trainer_factory(models, opt, train_ds, val_ds, lr_step, save_fn)` → object with .train(n)."""
    models, opt, train_ds, val_ds, epochs, lr_step, trace = build(scenario)

    def save(names):
        trace.append(['save', list(names), opt.lr])

    random.seed(1234)
    trainer = trainer_factory(models, opt, train_ds, val_ds, lr_step, save)
    best, best_epoch = trainer.train(epochs)
    return _clean({'trace': trace, 'best': best, 'best_epoch': best_epoch, 'lr': opt.lr,
                   'weights': {n: m.w for n, m in models.items()}})


def reference_factory(ref_trainer_module):
    """Wraps the scripted models the way `my_model/model.py` wraps the real ones: a model system
    that fills `context['losses']`, and a context maker around `dataset.get`."""
    class System:
        def __init__(self, models):
            self.models = models

        def train(self, context):
            context['losses'] = {n: m.train(*context['data'][n]) for n, m in self.models.items()}

        def test(self, context):
            context['losses'] = {n: m.test(*context['data'][n]) for n, m in self.models.items()}

    class Tracker:
        def reset(self):
            pass

        def message(self, *a):
            pass

    def factory(models, opt, train_ds, val_ds, lr_step, save):
        return ref_trainer_module.Trainer(
            System(models), lambda get, args: {'data': get(*args)}, models, train_ds, val_ds,
            progress_tracker=Tracker(), optimizer=opt, learning_rate_step=lr_step, save_weights_func=save)
    return factory


def ours_factory(models, opt, train_ds, val_ds, lr_step, save):
    from univer_ocr_b200.trainer import Trainer
    return Trainer(models, train_ds, val_ds, optimizer=opt, learning_rate_step=lr_step,
                   save_weights_func=save, batch_size=1, log=lambda *a, **k: None)
