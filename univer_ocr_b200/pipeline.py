"""Host <-> device pipelining for page-sharded inference.

The reference's PREDICT pipeline moves every batch host -> device -> host synchronously around
each sub-model (`my_model/model.py:688-717`: `to_gpu`, model, `from_gpu`).  On a B200 the four
forward passes of a 64-tile batch take ~1.6 ms while the PCIe copies of the same batch (104 MB
in, 121 MB out) take longer than that, so the copies are overlapped with compute:

    copy-in stream :  H2D(i+1)
    compute stream :            forward(i)
    copy-out stream:                        D2H(i-1)

Each in-flight batch owns a *slot* (device input buffers + pinned host output buffers); CUDA
events order the three streams per slot, the host only ever blocks on the D2H event of the slot
it is about to reuse.  Results are delivered in submission order.
"""
import ctypes

import numpy as np

from ._lib import lib
from .nn.gpu import CP, DeviceArray, stream as compute_stream


def _new_stream():
    s = ctypes.c_void_p()
    lib.uocr_stream_create(ctypes.byref(s))
    return s.value


def _new_event():
    e = ctypes.c_void_p()
    lib.uocr_event_create(ctypes.byref(e))
    return e.value


class _Slot:
    def __init__(self):
        self.dev_in = None          # {name: DeviceArray}
        self.host_out = None        # [pinned ndarray]
        self.dev_out = None         # keeps the step's outputs alive until their D2H finished
        self.ev_h2d, self.ev_comp, self.ev_d2h = _new_event(), _new_event(), _new_event()
        self.busy = False
        self.tag = None
        self.graph = None           # CapturedStep of this slot's forward (InferencePipeline(graph=True))


class InferencePipeline:
    """`fn(dict of DeviceArray) -> sequence of DeviceArray`, fed with dicts of (preferably pinned)
    host arrays of fixed shapes.

        pipe = InferencePipeline(step_fn, depth=3)
        for tag, outs in pipe.run(batches):      # outs: list of pinned host arrays (valid until
            consume(outs)                         # `depth` further batches have been submitted)

    Inputs and outputs keep their dtypes: a uint8 host array is uploaded as bytes into a uint8 device buffer (the
    reference's inputs are 8-bit PNG planes, `train_data_generator.py:24-37`; `fn` widens them with
    `glue.pixels_to_unit`), and whatever `fn` returns -- float32 maps, `glue.thresholded` uint8 masks,
    `glue.row_max_hits` uint8 tables -- is downloaded as is.  So a step can ship what the next host stage consumes
    (masks: `interpreter/interpreter.py:437-447`; hit tables: `:596-602`) instead of float32 everywhere."""

    def __init__(self, fn, depth=3, graph=False):
        """`graph=True`: every slot replays its forward as one CUDA graph (`CapturedStep` over the slot's fixed
        input buffers) instead of issuing it kernel by kernel."""
        self.fn = fn
        self.graph = bool(graph)
        self.depth = max(2, int(depth))
        CP.use_gpu()
        self.s_in, self.s_out = _new_stream(), _new_stream()
        self.slots = [_Slot() for _ in range(self.depth)]
        self.next = 0

    @staticmethod
    def _dtype_of(host_array):
        """Device dtype of an input: uint8 stays uint8 (image bytes), everything else is float32 storage."""
        return np.uint8 if np.asarray(host_array).dtype == np.uint8 else np.float32

    def _retire(self, slot):
        """Blocks until the slot's results are on the host and returns (tag, outputs)."""
        lib.uocr_event_sync(slot.ev_d2h)
        slot.busy = False
        slot.dev_out = None
        return slot.tag, slot.host_out

    def submit(self, host_inputs, tag=None):
        """Queues one batch; returns the (tag, outputs) of the batch that previously owned the
        slot, or None while the pipeline is filling."""
        slot = self.slots[self.next % self.depth]
        self.next += 1
        done = self._retire(slot) if slot.busy else None
        if slot.dev_in is None:
            # owned by the copy-in stream's pool for the pipeline's lifetime: a block recycled from the COMPUTE stream's
            # free list could still be read by kernels queued there when the first H2D (on s_in) overwrites it
            with CP.on_stream(self.s_in):
                slot.dev_in = {k: DeviceArray(v.shape, self._dtype_of(v)) for k, v in host_inputs.items()}
        comp = compute_stream()
        # copy-in: the previous forward that read these device buffers must have finished
        lib.uocr_stream_wait_event(self.s_in, slot.ev_comp)
        for k, v in host_inputs.items():
            src = np.ascontiguousarray(v, dtype=slot.dev_in[k].dtype)
            assert src.shape == slot.dev_in[k].shape, f'{k}: {src.shape} != {slot.dev_in[k].shape}'
            lib.uocr_memcpy_h2d(slot.dev_in[k].ptr, src.ctypes.data, src.nbytes, self.s_in)
        lib.uocr_event_record(slot.ev_h2d, self.s_in)
        # compute
        lib.uocr_stream_wait_event(comp, slot.ev_h2d)
        if self.graph:
            if slot.graph is None:
                slot.graph = CapturedStep(lambda s=slot: self.fn(s.dev_in))
            outs = list(slot.graph())
        else:
            outs = list(self.fn(slot.dev_in))
        lib.uocr_event_record(slot.ev_comp, comp)
        # copy-out
        if slot.host_out is None:
            slot.host_out = [CP.pinned_empty(o.shape, o.dtype) for o in outs]
        lib.uocr_stream_wait_event(self.s_out, slot.ev_comp)
        for o, h in zip(outs, slot.host_out):
            lib.uocr_memcpy_d2h(h.ctypes.data, o.ptr, o.nbytes, self.s_out)
        lib.uocr_event_record(slot.ev_d2h, self.s_out)
        slot.dev_out, slot.busy, slot.tag = outs, True, tag
        return done

    def drain(self):
        """Yields the results still in flight, in submission order."""
        for i in range(self.next - self.depth, self.next):
            if i < 0:
                continue
            slot = self.slots[i % self.depth]
            if slot.busy:
                yield self._retire(slot)

    def run(self, batches):
        for i, batch in enumerate(batches):
            done = self.submit(batch, tag=i)
            if done is not None:
                yield done
        yield from self.drain()


class ConcurrentBranches:
    """Runs independent pieces of work (e.g. the forward passes of sub-networks that do not feed each other) on
    separate CUDA streams, forked from and joined back to the compute stream:

        fork = ConcurrentBranches(3)
        para, line, char = fork.run(lambda: paragraph(monochrome(x)), lambda: line_net(l), lambda: char_net(c))

    Branch 0 stays on the compute stream; the others get their own stream (and their own allocator pool, so their
    intermediates are recycled in stream order).  Small-kernel chains (Line, Char: 15-50 us launches that fill only
    part of the GPU) then overlap with each other and with the big kernels' tails.  Results are valid on the
    compute stream when `run` returns."""

    def __init__(self, n):
        CP.use_gpu()
        self.side = [_new_stream() for _ in range(max(0, n - 1))]
        self.ev_fork = _new_event()
        self.ev_join = [_new_event() for _ in self.side]

    def run(self, *branches):
        assert len(branches) <= len(self.side) + 1
        main = compute_stream()
        lib.uocr_event_record(self.ev_fork, main)
        outs = [None] * len(branches)
        for i, fn in enumerate(branches[1:]):
            s = self.side[i]
            lib.uocr_stream_wait_event(s, self.ev_fork)
            with CP.on_stream(s):
                outs[i + 1] = fn()
            lib.uocr_event_record(self.ev_join[i], s)
        outs[0] = branches[0]()
        for i in range(len(branches) - 1):
            lib.uocr_stream_wait_event(main, self.ev_join[i])
        return outs


class CapturedStep:
    """A fixed launch sequence (one inference step over fixed device buffers) captured once into a CUDA graph and
    replayed with a single launch:

        step = CapturedStep(lambda: forward(inputs))     # `inputs`: DeviceArrays that are refilled in place
        outs = step()                                    # same output DeviceArrays every call

    The four sub-networks' forward is 14 kernels of 17-140 us; issued one by one from Python (ctypes call, fused-plan
    walk, stream fork / join events) that is ~0.2 ms of host time per step, which the GPU hides only while the host
    has a core to itself: with eight ranks on one host the step time stopped following the kernels
    (0.66 ms at 8 GPUs vs 0.50 ms at 1).  Replaying a graph costs one launch.

    `fn` runs `warmup` times eagerly first (fills allocator pools and the K-major weight caches), then once under
    stream capture -- forked streams (`ConcurrentBranches`) join the capture through their events.  Every block the
    sequence allocates and frees stays reserved for the graph (libuocr's allocator holds it, see uocr_graph_begin),
    so its memory footprint is the SUM of its intermediates.  `fn` must not synchronise, read values back or upload
    host data.  A parameter change (`CP.weights_generation`) triggers a re-capture on the next call.

    A training step can be captured too (`track_weights=False`: it updates the flat parameter buffers in place, so the
    graph stays valid): `after_replay` then has to do what the eager step does on the host besides launching kernels --
    for `DataParallel.train` that is bumping `CP.weights_generation` so that inference-side weight caches notice.
    Loss scalars returned by `fn` are device-resident (`LazyScalar`); read them after the replay of interest, each
    object caches its first read.  Keep collectives out of a captured step: single-process training replays correctly
    (tests/test_gpu_parity.py), but with torch's NCCL allreduce inside the captured sequence one of two 2-rank
    experiments deadlocked, so `bench.py` measures multi-rank training kernel by kernel."""

    def __init__(self, fn, warmup=2, track_weights=True, after_replay=None):
        self.fn, self.warmup = fn, warmup
        self.track_weights, self.after_replay = track_weights, after_replay
        self._exec, self._generation, self.out = None, None, None

    def _capture(self):
        for _ in range(self.warmup):
            self.fn()
        st = compute_stream()
        lib.uocr_graph_begin(st)
        exec_ = ctypes.c_void_p()
        try:
            out = self.fn()
        finally:
            lib.uocr_graph_end(st, ctypes.byref(exec_))
        self._exec, self.out, self._generation = exec_.value, out, CP.weights_generation

    def __call__(self):
        if self._exec is None or (self.track_weights and self._generation != CP.weights_generation):
            self.close()
            self._capture()
        lib.uocr_graph_launch(self._exec, compute_stream())
        if self.after_replay is not None:
            self.after_replay()
        return self.out

    def close(self):
        if self._exec is not None:
            CP.synchronize()                              # no replay may still be running on the held blocks
            lib.uocr_graph_destroy(self._exec)
            self._exec, self.out = None, None

    def __del__(self):
        try:
            self.close()
        except Exception:                                 # interpreter shutdown
            pass
