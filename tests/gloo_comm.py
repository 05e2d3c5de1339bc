"""CPU stand-in for `univer_ocr_b200.comm.Communicator` (tests only): the same host interface -- rank, world,
allreduce_host, broadcast_ints, barrier -- over torch.distributed's gloo backend, so that the multi-rank HOST logic
of the product (Trainer sharding / shuffle agreement / loss reduction, bucket scheduling) runs with world_size 2 in
the authoring container, where there is no GPU for NCCL."""
import numpy as np


class GlooComm:
    stream = None

    def __init__(self):
        import torch
        import torch.distributed as dist
        assert dist.is_initialized()
        self._torch, self._dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.reductions = 0

    def allreduce_host(self, values, op='sum'):
        t = self._torch.tensor(list(values), dtype=self._torch.float64)
        ops = {'sum': self._dist.ReduceOp.SUM, 'max': self._dist.ReduceOp.MAX, 'min': self._dist.ReduceOp.MIN}
        self._dist.all_reduce(t, op=ops[op])
        self.reductions += 1
        return t.tolist()

    def broadcast_ints(self, values, root=0):
        mine = [float(v) for v in values] if self.rank == root else [0.0] * len(values)
        return [int(round(v)) for v in self.allreduce_host(mine, 'sum')]

    def barrier(self):
        self.allreduce_host([1.0])

    def allreduce_sum_numpy(self, array):
        """In-place sum of a float32 / float64 NumPy array over the ranks (stands in for uocr_allreduce_sum_f32)."""
        t = self._torch.from_numpy(array)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        return array
