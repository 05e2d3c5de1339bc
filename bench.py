#!/usr/bin/env python
"""Headline benchmark: my_model inference images/s (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One STEP = one pass of the four sub-networks' forward over one batch of synthetic inputs per
GPU (page-sharded, no collective, weak scaling):
    Monochrome -> Paragraph on 64 page tiles (64, 496, 736, 1)
    Line  on 64 line tiles  (64, 128, 256, 1)
    Char  on 64 char lines  (64, 32, 256, 1)  (-> 16384 windows x 162 classes)
`value` = page tiles / s over all GPUs with inputs resident in HBM (CUDA events, max over
ranks); `e2e` = same through the public API with pinned HOST inputs: H2D of the step's inputs
and D2H of its outputs inside the timed region.  Prints ONE JSON line (rank 0).

`--impl reference` times the reference's CPU algorithm (oracle port: per-output-pixel NumPy
loops, float64 -- what the reference's `_forward_cpu` does) on all host cores, one sample per
core; see cpu_baseline.sample in its line.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAGE_HW, LINE_HW, CHAR_HW = (496, 736), (128, 256), (32, 256)
METRIC = 'my_model inference images/sec (page tiles through Monochrome->Paragraph + Line + Char forward)'
UNIT = 'images/s'


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        p = json.load(open(path))
        return {'hbm_gbs': float(p['hbm_gbs']), 'bf16_tflops': float(p['bf16_tflops']),
                'bf16_tflops_sustained': float(p.get('bf16_tflops_sustained', p['bf16_tflops'])),
                'source': 'measured (MEASURED_PEAKS.json)'}
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0,
                'source': 'fallback (B200_PROFILING.md)'}


# ------------------------------------------------------------------------------ synthetic data

def synth_tiles(rng, n, hw):
    """U[0,1) paper with sparse dark ink: value = 1 - Bernoulli(0.1) * U[0.5, 1)."""
    h, w = hw
    ink = (rng.random((n, h, w, 1), dtype=np.float32) < 0.1)
    x = 1.0 - ink * rng.uniform(0.5, 1.0, size=(n, h, w, 1)).astype(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)


# ------------------------------------------------------------------------------ clocks sampler

class ClockSampler:
    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu_index), f'--query-gpu={self.QUERY}',
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(max(mx)) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------ distributed glue

class Dist:
    def __init__(self, want_gpus):
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        self.torch = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local_rank))
            self.torch, self.dist = torch, dist

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max(self, value):
        if self.world == 1:
            return value
        t = self.torch.tensor([value], dtype=self.torch.float64, device='cuda')
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, value):
        if self.world == 1:
            return value
        t = self.torch.tensor([value], dtype=self.torch.float64, device='cuda')
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


# ------------------------------------------------------------------------------ B200 arm

def run_b200(args):
    dist = Dist(args.gpus)
    os.environ.setdefault('UOCR_DEVICE', str(dist.local_rank))
    import univer_ocr_b200.nn as nn
    from univer_ocr_b200 import my_model, roofline
    from univer_ocr_b200._lib import launch_count, lib
    from univer_ocr_b200.nn.progress_tracker import BaseProgressTracker, CudaEventTracker

    nn.CP.use_gpu()
    nn.CP.set_math_mode(args.math)
    B = args.batch
    rng = np.random.default_rng(1234 + dist.rank)
    np.random.seed(1234 + dist.rank)                    # layer initialisers draw from np.random
    models = {
        'monochrome': my_model.make_monochrome((B, *PAGE_HW, 1)),
        'paragraph': my_model.make_paragraph((B, *PAGE_HW, 1)),
        'line': my_model.make_line((B, *LINE_HW, 1)),
        'char': my_model.make_char((B, *CHAR_HW, 1)),
    }
    # centred Char FC weights: the reference's all-positive init saturates the softmax
    for key, p in models['char'].params().items():
        if 'dense' in key:
            w = p.value.get()
            p.value = (w - w.mean()) * 0.2

    n_sets = 2                                           # rotate inputs: 2 x 103 MB > L2, plus
    host_sets = []                                       # ~3.3 GB of intermediates per step
    for _ in range(n_sets):
        hs = {}
        for key, hw in (('page', PAGE_HW), ('line', LINE_HW), ('char', CHAR_HW)):
            buf = nn.CP.pinned_empty((B, *hw, 1), np.float32)
            buf[...] = synth_tiles(rng, B, hw)
            hs[key] = buf
        host_sets.append(hs)
    dev_sets = [{k: nn.CP.copy(np.asarray(v)) for k, v in hs.items()} for hs in host_sets]
    nn.CP.synchronize()

    # the Monochrome -> Paragraph chain, Line and Char do not feed each other: three forked streams, joined on the
    # compute stream (UOCR_BENCH_SERIAL=1: one stream, e.g. for per-layer timing)
    from univer_ocr_b200.pipeline import ConcurrentBranches
    fork = None if os.environ.get('UOCR_BENCH_SERIAL') == '1' else ConcurrentBranches(3)

    def step_serial(inp):
        mono = models['monochrome'].predict(inp['page'])[0]
        para = models['paragraph'].predict(mono)[0]
        line = models['line'].predict(inp['line'])[0]
        char = models['char'].predict(inp['char'])[0]
        return para, line, char

    def step(inp):
        if fork is None:
            return step_serial(inp)
        return tuple(fork.run(lambda: models['paragraph'].predict(models['monochrome'].predict(inp['page'])[0])[0],
                              lambda: models['line'].predict(inp['line'])[0],
                              lambda: models['char'].predict(inp['char'])[0]))

    stream = nn.CP.stream()

    def event():
        e = ctypes.c_void_p()
        lib.uocr_event_create(ctypes.byref(e))
        return e.value

    # ---------------- device-resident throughput (`value`) ----------------
    # the step over a resident input set is a fixed launch sequence: captured once per set into a CUDA graph
    # (pipeline.CapturedStep; UOCR_BENCH_GRAPH=0 issues the 14 launches one by one instead)
    use_graph = os.environ.get('UOCR_BENCH_GRAPH', '1') != '0'
    eager_step = step
    graph_note = None
    if use_graph:
        from univer_ocr_b200.pipeline import CapturedStep
        captured = {id(inp): CapturedStep(lambda inp=inp: eager_step(inp)) for inp in dev_sets}
        try:                                             # capture now; a failure falls back to eager launches, loudly
            for graph in captured.values():
                graph()
            nn.CP.synchronize()
        except Exception as exc:                         # noqa: BLE001
            graph_note = f'graph capture failed, launching kernel by kernel: {exc}'
            print(graph_note, file=sys.stderr, flush=True)
            captured, use_graph = {}, False

        def step(inp):
            graph = captured.get(id(inp))
            return graph() if graph is not None else eager_step(inp)

    for i in range(max(args.warmup, n_sets)):
        step(dev_sets[i % n_sets])
    nn.CP.synchronize()
    sampler = ClockSampler(dist.local_rank)
    sampler.start()
    dist.barrier()
    e0, e1 = event(), event()
    launches0 = launch_count()
    lib.uocr_event_record(e0, stream)
    for i in range(args.steps):
        step(dev_sets[i % n_sets])
    lib.uocr_event_record(e1, stream)
    lib.uocr_event_sync(e1)
    nn.CP.synchronize()
    launches = launch_count() - launches0
    dist.barrier()
    ms = ctypes.c_float(0)
    lib.uocr_event_elapsed_ms(e0, e1, ctypes.byref(ms))
    ms_per_step = dist.max(ms.value / args.steps)
    value = B * dist.world / (ms_per_step / 1e3)

    # ---------------- end to end through the public API with host buffers (`e2e`) ----------------
    # (a) synchronous call: H2D -> forward -> D2H -> sync, one batch at a time
    outs_host = None

    def e2e_step(hs):
        nonlocal outs_host
        inp = {k: nn.CP.copy(v) for k, v in hs.items()}          # async H2D from pinned memory
        outs = step(inp)
        if outs_host is None:
            outs_host = [nn.CP.pinned_empty(o.shape, np.float32) for o in outs]
        for o, h in zip(outs, outs_host):
            lib.uocr_memcpy_d2h(h.ctypes.data, o.ptr, o.nbytes, stream)
        nn.CP.synchronize()                                       # results are on the host
        return outs

    for i in range(3):
        e2e_step(host_sets[i % n_sets])
    dist.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(host_sets[i % n_sets])
    sync_s = dist.max((time.perf_counter() - t0) / args.steps)
    # (b) the pipelined public API (univer_ocr_b200.pipeline.InferencePipeline): H2D of batch i+1,
    # forward of batch i and D2H of batch i-1 overlap on three streams; every batch is still
    # uploaded from and downloaded to host memory inside the timed region, fill and drain included
    from univer_ocr_b200.pipeline import InferencePipeline
    pipe = InferencePipeline(eager_step, depth=3, graph=use_graph)
    for i in range(6):
        pipe.submit(host_sets[i % n_sets], i)
    for _ in pipe.drain():
        pass
    dist.barrier()
    t0 = time.perf_counter()
    delivered = 0
    for i in range(args.steps):
        if pipe.submit(host_sets[i % n_sets], i) is not None:
            delivered += 1
    for _ in pipe.drain():
        delivered += 1
    e2e_s = dist.max((time.perf_counter() - t0) / args.steps)
    assert delivered == args.steps
    dist.barrier()
    clocks = sampler.stop()
    h2d = sum(v.nbytes for v in host_sets[0].values())
    d2h = sum(h.nbytes for h in outs_host)
    e2e_value = B * dist.world / e2e_s

    # ---------------- per-layer device time -> dominant kernel -> roofline ----------------
    peaks = load_peaks()
    in_shapes = {'monochrome': (B, *PAGE_HW, 1), 'paragraph': (B, *PAGE_HW, 1),
                 'line': (B, *LINE_HW, 1), 'char': (B, *CHAR_HW, 1)}
    for i in range(2):                                           # the serial order allocates from the compute stream's
        step_serial(dev_sets[i % n_sets])                        # pool: fill it before timing layers
    nn.CP.synchronize()
    tracker = CudaEventTracker()
    work = {}
    for mname, model in models.items():
        work.update(roofline.plan_work(model, in_shapes[mname], training=False))
        for layer in model.layers.values():
            layer.progress_tracker = tracker
    prof_steps = max(2, min(args.steps, 5))
    for i in range(prof_steps):
        step_serial(dev_sets[i % n_sets])                      # per-layer event pairs: one stream
    per_layer = tracker.summary_ms()
    for model in models.values():
        for layer in model.layers.values():
            layer.progress_tracker = BaseProgressTracker()
    total_ms = sum(v[0] for v in per_layer.values()) / prof_steps
    breakdown = sorted(((name, v[0] / v[1]) for (name, ev), v in per_layer.items() if ev == 'forward'),
                       key=lambda t: -t[1])
    top_name, top_ms = breakdown[0]
    wk = work[top_name]
    if wk['bound'] == 'tensor':
        tf32_peak = peaks['bf16_tflops'] / 2.0
        achieved = wk['flops'] / (top_ms / 1e3) / 1e12
        roof = {'bound': 'tensor', 'achieved': achieved, 'peak': tf32_peak, 'unit': 'TFLOP/s',
                'frac': achieved / tf32_peak, 'traffic': None,
                'peak_note': f'TF32 dense taken as 1/2 of bf16 burst {peaks["bf16_tflops"]} TF/s, {peaks["source"]}'}
    else:
        achieved = wk['bytes'] / (top_ms / 1e3) / 1e9
        roof = {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                'frac': achieved / peaks['hbm_gbs'], 'traffic': None, 'peak_note': peaks['source']}
    try:                                            # DRAM traffic of that kernel from the committed ncu capture
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get('+'.join(wk.get('fused', [top_name])))
        if tr and tr['batch'] == B:
            roof['traffic'] = tr['bytes']
            roof['traffic_source'] = tr['profile']
            if tr.get('note'):
                roof['note'] = tr['note']
    except (OSError, ValueError):
        pass
    roof.update({'kernel': '+'.join(wk.get('fused', [top_name])), 'kernel_ms': top_ms, 'share_of_step': top_ms / total_ms,
                 'algorithmic_bytes': wk['bytes'], 'algorithmic_flops': wk['flops']})
    layers_out = []
    for name, lms in breakdown[:12]:
        w_ = work[name]
        layers_out.append({'layer': '+'.join(w_.get('fused', [name])), 'ms': round(lms, 4), 'bound': w_['bound'],
                           'GBps': round(w_['bytes'] / (lms / 1e3) / 1e9, 1),
                           'TFLOPs': round(w_['flops'] / (lms / 1e3) / 1e12, 2)})

    train = None if args.no_train else measure_train(args, dist, nn, my_model, models, dev_sets, rng, B, event)

    result = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': dist.world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32' if args.math == 'fp32' else 'tf32/f32',
        'data': 'synthetic',
        'config': {'workload': 'BASELINE configs[1]: my_model inference, batch 64 synthetic page tiles '
                               '(64,496,736,1) Monochrome->Paragraph + Line (64,128,256,1) + Char '
                               '(64,32,256,1), per GPU; pages sharded across GPUs, no collective',
                   'batch_per_gpu': B, 'math_mode': args.math, 'weights': 'random init (kaiming_uniform, seeded)',
                   'streams': 'Monochrome->Paragraph, Line and Char forward on three forked CUDA streams joined per step' if fork is not None else 'one stream',
                   'launch': 'one CUDA graph replay per step (captured per resident input set)' if use_graph else (graph_note or 'kernel by kernel'),
                   'l2_policy': 'inputs rotate over 2 resident sets (206 MB) and each step streams '
                                '~3.3 GB of intermediates: working set >> 126 MB L2'},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': e2e_s * 1e3,
                'api': 'univer_ocr_b200.pipeline.InferencePipeline(depth=3%s): pinned host in -> H2D -> forward -> D2H -> pinned host out' % (', graph=True' if use_graph else ''),
                'timing': 'host wall clock over K submitted batches incl. pipeline fill and drain, max over ranks',
                'sync_value': B * dist.world / sync_s, 'sync_ms_per_step': sync_s * 1e3},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': roof,
        'layers': layers_out,
    }
    if train is not None:
        result['train'] = train
    if dist.rank == 0 and dist.world == 1 and not args.no_cpu_baseline:
        result['cpu_baseline'] = cpu_baseline(budget_s=args.cpu_budget, cores=1)
    if dist.rank == 0:
        print(json.dumps(result), flush=True)
    dist.close()


def measure_train(args, dist, nn, my_model, models, dev_sets, rng, B, event):
    """BASELINE configs[2]: one data-parallel training step of the four sub-networks (forward +
    loss + backward + NCCL gradient allreduce + fused L2/Adam), batch B per GPU (weak scaling)."""
    import ctypes as ct
    from univer_ocr_b200._lib import launch_count, lib
    from univer_ocr_b200.parallel import DataParallel
    opt = nn.optimizers.Adam(lr=0.0015)
    dps = {name: DataParallel(model, optimizer=opt) for name, model in models.items()}
    inp = dev_sets[0]
    targets = {
        'monochrome': nn.CP.copy((rng.random((B, *PAGE_HW, 1), dtype=np.float32) < 0.2).astype(np.float32)),
        'paragraph': nn.CP.copy((rng.random((B, *PAGE_HW, 1), dtype=np.float32) < 0.2).astype(np.float32)),
        'line': nn.CP.copy((rng.random((B, *LINE_HW, 2), dtype=np.float32) < 0.2).astype(np.float32)),
    }
    onehot = np.zeros((B * CHAR_HW[1], my_model.N_CHARS), dtype=np.float32)
    onehot[np.arange(onehot.shape[0]), rng.integers(0, my_model.N_CHARS, size=onehot.shape[0])] = 1
    targets['char'] = nn.CP.copy(onehot)
    feeds = {'monochrome': inp['page'], 'paragraph': inp['page'], 'line': inp['line'], 'char': inp['char']}
    stream = nn.CP.stream()

    # the four sub-networks train independently (own parameters, own gradient allreduce): forked streams, joined per step
    from univer_ocr_b200.pipeline import ConcurrentBranches
    fork = None if os.environ.get('UOCR_BENCH_SERIAL') == '1' else ConcurrentBranches(len(dps))

    def step():
        names = list(dps)
        if fork is None:
            return {name: dps[name].train(feeds[name], targets[name]) for name in names}
        outs = fork.run(*[(lambda name=name: dps[name].train(feeds[name], targets[name])) for name in names])
        return dict(zip(names, outs))

    for _ in range(3):
        losses = step()
    nn.CP.synchronize()
    # the training step is a fixed launch sequence over fixed buffers as well (parameters, gradients and Adam state are
    # updated in place) and replays correctly from a CUDA graph (tests/test_gpu_parity.py, 2.85 -> 2.76 ms at one GPU).
    # Opt-in only (UOCR_BENCH_TRAIN_GRAPH=1): with world > 1 the captured sequence contains torch's NCCL allreduce, and
    # one of two 2-rank experiments with a captured allreduce deadlocked, so every N is measured kernel by kernel.
    launch_mode = 'kernel by kernel'
    if os.environ.get('UOCR_BENCH_TRAIN_GRAPH', '0') == '1':
        from univer_ocr_b200.pipeline import CapturedStep

        def bump():
            nn.CP.weights_generation += 1

        eager_train_step = step
        graph = CapturedStep(eager_train_step, warmup=0, track_weights=False, after_replay=bump)
        try:
            graph()
            nn.CP.synchronize()
            step = graph
            launch_mode = 'one CUDA graph replay per step'
        except Exception as exc:                         # noqa: BLE001
            print(f'train-step graph capture failed, launching kernel by kernel: {exc}', file=sys.stderr, flush=True)
            step = eager_train_step
    dist.barrier()
    e0, e1 = event(), event()
    launches0 = launch_count()
    steps = max(2, min(args.steps, 5))
    lib.uocr_event_record(e0, stream)
    for _ in range(steps):
        losses = step()
    lib.uocr_event_record(e1, stream)
    lib.uocr_event_sync(e1)
    nn.CP.synchronize()
    launches = launch_count() - launches0
    dist.barrier()
    ms = ct.c_float(0)
    lib.uocr_event_elapsed_ms(e0, e1, ct.byref(ms))
    ms_per_step = dist.max(ms.value / steps)
    n_params = sum(dp.flat.total for dp in dps.values())
    return {'metric': 'my_model train-step images/sec (fwd + loss + bwd + grad allreduce + L2 + Adam of all four sub-networks)',
            'value': B * dist.world / (ms_per_step / 1e3), 'unit': UNIT, 'ms_per_step': ms_per_step,
            'steps': steps, 'batch_per_gpu': B, 'global_batch': B * dist.world,
            'allreduce_bytes_per_step': 4 * n_params if dist.world > 1 else 0, 'gpu_launches': int(launches),
            'launch': launch_mode,
            'losses': {k: float(v['output_losses'][0]) for k, v in losses.items()}}


# ------------------------------------------------------------------------------ CPU arms
# (the only place outside tests/ and smoke() that executes oracle/)

def _cpu_sample(frac, seed=1234):
    """Forward of the four sub-networks over a strip of ONE image each with the loop-form oracle
    port; returns seconds.  `frac` = fraction of each tile's rows (columns for Char) processed."""
    from oracle import np_models
    rng = np.random.default_rng(seed)
    rows_p = max(16, int(round(PAGE_HW[0] * frac / 16)) * 16)
    rows_l = max(4, int(round(LINE_HW[0] * frac / 4)) * 4)
    cols_c = max(8, int(round(CHAR_HW[1] * frac)))
    t0 = time.perf_counter()
    x = synth_tiles(rng, 1, (rows_p, PAGE_HW[1])).astype(np.float64)
    for name in ('monochrome', 'paragraph'):
        spec = np_models.net_spec(name)
        x = np_models.forward(spec, np_models.init_weights(spec, rng), x, loop=True)
    spec = np_models.net_spec('line')
    np_models.forward(spec, np_models.init_weights(spec, rng),
                      synth_tiles(rng, 1, (rows_l, LINE_HW[1])).astype(np.float64), loop=True)
    spec = np_models.net_spec('char')
    np_models.forward(spec, np_models.init_weights(spec, rng),
                      synth_tiles(rng, 1, (CHAR_HW[0], cols_c)).astype(np.float64), loop=True)
    dt = time.perf_counter() - t0
    done = (rows_p / PAGE_HW[0] + rows_l / LINE_HW[0] + cols_c / CHAR_HW[1]) / 3.0
    return dt, done, (rows_p, rows_l, cols_c)


def _cpu_worker(frac):
    os.environ['OMP_NUM_THREADS'] = '1'
    return _cpu_sample(frac)


def cpu_baseline(budget_s=20.0, cores=1):
    """Oracle port ("port": NumPy restatement with the reference's per-pixel loops) on a bounded
    sample: a strip of one image per sub-network, sized for ~budget_s of single-core work."""
    probe_t, probe_done, _ = _cpu_sample(0.04)
    frac = float(min(1.0, max(0.04, 0.04 * budget_s / max(probe_t, 1e-3))))
    dt, done, dims = _cpu_sample(frac)
    return {'value': done / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'loop-form NumPy float64 port of the reference CPU path, single process: '
                      f'{dims[0]}/496 rows of one page tile through Monochrome->Paragraph, '
                      f'{dims[1]}/128 rows of one line tile, {dims[2]}/256 columns of one char line; '
                      f'{dt:.1f} s; images/s = mean tile fraction / time'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    total = args.steps + args.warmup
    per_step_budget = max(2.0, min(20.0, 150.0 / max(total, 1)))
    probe_t, _, _ = _cpu_sample(0.04)
    frac = float(min(1.0, max(0.04, 0.04 * per_step_budget / max(probe_t, 1e-3))))
    times, done, dims = [], None, None
    with mp.get_context('fork').Pool(cores) as pool:
        for i in range(total):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [frac] * cores)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt)
            done, dims = res[0][1], res[0][2]
    sec = float(np.mean(times))
    value = cores * done / sec
    sample = (f'{cores} processes x (loop-form NumPy float64 port: {dims[0]}/496 page-tile rows through '
              f'Monochrome->Paragraph + {dims[1]}/128 line-tile rows + {dims[2]}/256 char-line columns) '
              f'per step; images/s = cores x mean tile fraction / step time')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'BASELINE configs[1] on host cores (reference CPU algorithm, oracle port)',
                   'batch_per_gpu': args.batch},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='tiles per GPU per step')
    ap.add_argument('--math', default=os.environ.get('UOCR_MATH', 'tf32'), choices=['fp32', 'tf32'],
                    help='tf32: tcgen05 TF32 kernels for the dense contractions (default); fp32: FFMA check mode')
    ap.add_argument('--cpu-budget', type=float, default=15.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train', action='store_true', help='skip the training-step measurement')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
