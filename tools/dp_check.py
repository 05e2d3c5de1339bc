"""Multi-GPU equivalence check (run under torchrun with >= 2 ranks): two DataParallel steps on batch shards ==
two single-process steps on the whole batch.  The same comparison runs inside bench.py at WORLD_SIZE > 1
(`train.dp_max_rel_err`); this is the stand-alone form.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import univer_ocr_b200.nn as nn  # noqa: E402
from univer_ocr_b200 import comm as comm_, my_model  # noqa: E402

nn.CP.use_gpu()
comm = comm_.init_from_env()
err = bench.dp_self_check(comm, nn, my_model)
if comm.rank == 0:
    print(f'DP CHECK {"PASSED" if err <= 2e-4 else "FAILED"}: max rel err {err:.3e} over {comm.world} ranks '
          f'(NCCL {comm_.Communicator.version() if comm.world > 1 else "-"})')
comm.close()
sys.exit(0 if err <= 2e-4 else 1)
