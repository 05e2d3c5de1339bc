"""Process-group plumbing of the data-parallel step: one process per GPU, NCCL through libuocr.

The reference has no distributed backend; north_star asks for a batch-sharded train
step with "an NCCL gradient allreduce over NVLink overlapped with backward" and no PyTorch in the
path.  `Communicator` is the host side of `uocr_nccl_*` / `uocr_allreduce_*` (include/uocr.h):

    comm = comm.init_from_env()          # RANK / WORLD_SIZE / LOCAL_RANK as torchrun exports them
    comm.allreduce_sum(device_array)     # in place, asynchronous on comm.stream (or a given stream)
    comm.allreduce_host([..], 'max')     # a few float64 values through the same communicator
    comm.barrier()

Rendezvous: rank 0 creates the NCCL unique id and publishes it through a file that only this launch
can see (name = MASTER_PORT + torchrun run id + the launcher's pid, which all ranks of a node share as
their parent); the other ranks poll for it.  Single node only -- the scope of this path (one NVSwitch box).
`UOCR_RDZV_FILE` overrides the path (e.g. a shared file system for an external launcher).

Host-side logic (`Trainer`, bucket scheduling) only needs `rank`, `world`, `allreduce_host` and
`broadcast_ints`; tests drive it with a gloo-backed stand-in on CPU.
"""
import ctypes
import importlib.util
import os
import tempfile
import time

import numpy as np

from ._lib import lib

_OPS = {'sum': 0, 'max': 1, 'min': 2}
_current = None


def _bundled_nccl_path():
    """The NCCL wheel next to PyTorch (nvidia-nccl-cu12), WITHOUT importing torch: if torch is imported later in
    the same process its libtorch_cuda resolves the SONAME libnccl.so.2 to whatever is mapped already, so libuocr
    maps the copy torch expects."""
    try:
        spec = importlib.util.find_spec('nvidia.nccl')
    except (ImportError, ValueError):
        spec = None
    if spec is None or not spec.submodule_search_locations:
        return None
    for root in spec.submodule_search_locations:
        path = os.path.join(root, 'lib', 'libnccl.so.2')
        if os.path.exists(path):
            return path
    return None


def rendezvous_file(env=os.environ):
    path = env.get('UOCR_RDZV_FILE')
    if path:
        return path
    tag = f"{env.get('MASTER_PORT', '0')}_{env.get('TORCHELASTIC_RUN_ID', 'none')}_{os.getppid()}"
    return os.path.join(tempfile.gettempdir(), f'uocr_nccl_{tag}.id')


class SingleProcess:
    """world == 1: every collective is the identity."""
    rank, world, stream = 0, 1, None

    def allreduce_sum(self, arr, start=0, count=None, stream=None):
        pass

    def broadcast(self, arr, root=0, stream=None):
        pass

    def allreduce_host(self, values, op='sum'):
        return [float(v) for v in values]

    def broadcast_ints(self, values, root=0):
        return [int(v) for v in values]

    def barrier(self):
        pass

    def close(self):
        pass


class Communicator(SingleProcess):
    """The process's NCCL communicator (one per process, bound to the process's device)."""

    def __init__(self, rank, world, unique_id):
        from .nn.gpu import CP, DeviceArray
        assert len(unique_id) == 128
        CP.use_gpu()                                      # selects the device (LOCAL_RANK / UOCR_DEVICE)
        path = os.environ.get('UOCR_NCCL_LIB') or _bundled_nccl_path()
        lib.uocr_nccl_load(path.encode() if path else None)
        lib.uocr_nccl_init(int(rank), int(world), bytes(unique_id))
        self.rank, self.world = int(rank), int(world)
        s = ctypes.c_void_p()
        lib.uocr_stream_create(ctypes.byref(s))
        self.stream = s.value                             # side stream for collectives that overlap compute
        self._scratch = DeviceArray((256,), np.float64)   # host-value reductions
        self._CP = CP

    @staticmethod
    def version():
        v = ctypes.c_int(0)
        lib.uocr_nccl_version(ctypes.byref(v))
        return v.value

    def allreduce_sum(self, arr, start=0, count=None, stream=None):
        """In-place sum of `arr` (float32 DeviceArray), or of its flat range [start, start + count)."""
        count = arr.size - start if count is None else count
        lib.uocr_allreduce_sum_f32(arr.ptr + 4 * start, count, self.stream if stream is None else stream)

    def broadcast(self, arr, root=0, stream=None):
        lib.uocr_broadcast_f32(arr.ptr, arr.size, int(root), self.stream if stream is None else stream)

    def allreduce_host(self, values, op='sum'):
        """A few Python floats reduced over the ranks (float64); synchronises the compute stream."""
        values = np.asarray(list(values), dtype=np.float64)
        out = np.empty_like(values)
        st = self._CP.stream()
        for lo in range(0, values.size, self._scratch.size):
            part = np.ascontiguousarray(values[lo:lo + self._scratch.size])
            lib.uocr_memcpy_h2d(self._scratch.ptr, part.ctypes.data, part.nbytes, st)
            lib.uocr_allreduce_f64(self._scratch.ptr, part.size, _OPS[op], st)
            got = np.empty_like(part)
            lib.uocr_memcpy_d2h(got.ctypes.data, self._scratch.ptr, part.nbytes, st)
            lib.uocr_stream_sync(st)
            out[lo:lo + part.size] = got
        return out.tolist()

    def broadcast_ints(self, values, root=0):
        """Rank `root`'s list of (|v| < 2^53) integers on every rank -- e.g. an epoch's shuffled sample order."""
        mine = [float(v) for v in values] if self.rank == root else [0.0] * len(values)
        return [int(round(v)) for v in self.allreduce_host(mine, 'sum')]

    def barrier(self):
        self.allreduce_host([1.0], 'sum')

    def close(self):
        global _current
        lib.uocr_device_sync()
        lib.uocr_nccl_finalize()
        if _current is self:
            _current = None


def init(rank, world, rdzv_file=None, timeout_s=180.0):
    """Creates the process-wide communicator (world 1: the no-op `SingleProcess`)."""
    global _current
    if _current is not None:
        return _current
    if world <= 1:
        _current = SingleProcess()
        return _current
    path = rdzv_file or rendezvous_file()
    if rank == 0:
        uid = ctypes.create_string_buffer(128)
        nccl = os.environ.get('UOCR_NCCL_LIB') or _bundled_nccl_path()
        lib.uocr_nccl_load(nccl.encode() if nccl else None)
        lib.uocr_nccl_unique_id(uid)
        tmp = f'{path}.{os.getpid()}.tmp'
        with open(tmp, 'wb') as f:
            f.write(uid.raw)
        os.replace(tmp, path)                             # atomic: readers see all 128 bytes or no file
        unique_id = uid.raw
    else:
        deadline = time.monotonic() + timeout_s
        while True:
            try:
                with open(path, 'rb') as f:
                    unique_id = f.read()
                if len(unique_id) == 128:
                    break
            except OSError:
                pass
            if time.monotonic() > deadline:
                raise TimeoutError(f'rank {rank}: no NCCL id at {path} after {timeout_s:.0f} s')
            time.sleep(0.02)
    _current = Communicator(rank, world, unique_id)
    _current.barrier()                                    # every rank has read the id
    if rank == 0:
        try:
            os.unlink(path)
        except OSError:
            pass
    return _current


def init_from_env(env=os.environ):
    return init(int(env.get('RANK', '0')), int(env.get('WORLD_SIZE', '1')))


def current():
    """The communicator `init` created, or the single-process stand-in."""
    return _current if _current is not None else SingleProcess()


def use(comm):
    """Installs `comm` (anything with the SingleProcess interface) as the process-wide communicator."""
    global _current
    _current = comm
    return comm
