"""Per-layer timing hooks (reference: nn/progress_tracker.py).

Every layer / model `forward` and `backward` still reports to `progress_tracker.start_tracking /
stop_tracking(name, event)` like the reference's `@track_method` (:100-109), so a UI handler
written for the reference keeps working.  Difference: kernels are asynchronous here, so wall
clock around a call measures launch time unless `ProgressTracker(sync=True)` is used, which
synchronises the stream at every stop (what the reference's cuda.synchronize() after each
kernel did implicitly) -- opt-in, because a per-layer sync serialises the pipeline.
"""
from datetime import datetime
from functools import wraps


class Event:
    def __init__(self, name):
        self.name = name
        self.reset()

    def reset(self):
        self.done, self.started, self.stopped, self.time, self.counter = False, None, None, None, 0

    def start(self):
        self.done, self.started = False, datetime.now()

    def stop(self):
        self.stopped = datetime.now()
        delta = self.stopped - self.started
        self.time = delta if self.time is None else self.time + delta
        self.done = True
        self.counter += 1

    def to_dict(self):
        return {k: getattr(self, k) for k in ('name', 'done', 'started', 'stopped', 'time', 'counter')}


class BaseProgressTracker:
    """No-op tracker (the default on every layer)."""

    def __init__(self, *args, **kwargs):
        pass

    def register_layer(self, name):
        pass

    def get_summary(self):
        return {}

    def start_tracking(self, name, event):
        pass

    def stop_tracking(self, name, event):
        pass

    def message(self, message, data=None):
        pass

    def reset(self):
        pass


class ProgressTracker(BaseProgressTracker):
    def __init__(self, handler=print, sync=False):
        self.layers = {}
        self.handler = handler
        self.sync = sync

    def register_layer(self, name):
        self.layers[name] = {}

    def get_summary(self):
        return {name: [ev.to_dict() for ev in events.values()] for name, events in self.layers.items()}

    def start_tracking(self, name, event):
        events = self.layers.setdefault(name, {})
        if event not in events:
            events[event] = Event(event)
        events[event].start()
        self.handler(event, self.get_summary())

    def stop_tracking(self, name, event):
        if self.sync:
            from .gpu import CP
            CP.synchronize()
        self.layers[name][event].stop()
        self.handler(event, self.get_summary())

    def message(self, message, data=None):
        self.handler(message, data)

    def reset(self):
        self.handler('reset')
        for events in self.layers.values():
            for ev in events.values():
                ev.reset()


def track_method(event):
    def decorator(func):
        @wraps(func)
        def wrapper(self, *args, **kwargs):
            tracker = self.progress_tracker
            tracker.start_tracking(self.name, event)
            try:
                return func(self, *args, **kwargs)
            finally:
                tracker.stop_tracking(self.name, event)
        return wrapper
    return decorator


def track_function(name, event, progress_tracker):
    if progress_tracker is None:
        return lambda func: func
    progress_tracker.register_layer(name)

    def decorator(func):
        @wraps(func)
        def wrapper(*args, **kwargs):
            progress_tracker.start_tracking(name, event)
            try:
                return func(*args, **kwargs)
            finally:
                progress_tracker.stop_tracking(name, event)
        return wrapper
    return decorator


class CudaEventTracker(BaseProgressTracker):
    """Device-side per-layer timing through the same hook: a CUDA event pair is recorded on the
    compute stream around every tracked call; `summary_ms()` synchronises once and returns
    {(layer, event): [total_ms, calls]}.  Leaf layers only (a Model's own forward/backward
    contains its layers' time)."""

    def __init__(self, leaf_names=None):
        self._open = {}
        self._pairs = []
        self.leaf_names = set(leaf_names) if leaf_names is not None else None

    @staticmethod
    def _event():
        import ctypes
        from .._lib import lib
        e = ctypes.c_void_p()
        lib.uocr_event_create(ctypes.byref(e))
        return e.value

    def start_tracking(self, name, event):
        if self.leaf_names is not None and name not in self.leaf_names:
            return
        from .._lib import lib
        from .gpu import stream
        e = self._event()
        lib.uocr_event_record(e, stream())
        self._open[name, event] = e

    def stop_tracking(self, name, event):
        start = self._open.pop((name, event), None)
        if start is None:
            return
        from .._lib import lib
        from .gpu import stream
        e = self._event()
        lib.uocr_event_record(e, stream())
        self._pairs.append((name, event, start, e))

    def summary_ms(self):
        import ctypes
        from .._lib import lib
        from .gpu import CP
        CP.synchronize()
        out = {}
        for name, event, e0, e1 in self._pairs:
            ms = ctypes.c_float(0)
            lib.uocr_event_elapsed_ms(e0, e1, ctypes.byref(ms))
            rec = out.setdefault((name, event), [0.0, 0])
            rec[0] += ms.value
            rec[1] += 1
            lib.uocr_event_destroy(e0)
            lib.uocr_event_destroy(e1)
        self._pairs = []
        return out
