#!/usr/bin/env python
"""MAOOAM-36 Lyapunov spectrum of an ensemble sharded over the GPUs of one node (BASELINE.json config 5).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/lyap_multi_gpu.py [members_per_gpu] [window]

Rank g integrates its block of members with the Benettin kernels (qgs_b200.toolbox.lyapunov.LyapunovsEstimator ==
the reference's API, lyapunov.py:232-358); the only exchange is the final all-reduce of 2 n_vec + 1 sums over NCCL.
Prints one JSON line on rank 0.
"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    import torch.distributed as dist
    from qgs_b200 import _lib
    from qgs_b200.ensemble import sharded_lyapunov_spectrum
    from qgs_b200.functions.tendencies import tendencies_from_tensor
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    window = float(sys.argv[2]) if len(sys.argv) > 2 else 20.
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.init(local)
    z = np.load(os.path.join(REPO, "tests", "golden", "tensor_maooam36.npz"))
    f, Df = tendencies_from_tensor(36, z["coo"], z["val"], z["jcoo"], z["jval"])
    ic = np.random.default_rng(21217).random((per_gpu * world, 36)) * 0.01
    np.random.seed(1000 + rank)                      # start bases: numpy's generator, per rank
    for attempt in range(2):                          # first pass warms the pool and the kernels
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mean, sem, res = sharded_lyapunov_spectrum(f, Df, ic, 0., window / 2, window, 0.1, 0.1, write_steps=10)
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - t0
    steps = int(round(window / 0.1))
    if rank == 0:
        print(json.dumps({"what": "MAOOAM-36 Lyapunov spectrum, sharded ensemble", "n_gpus": world,
                          "members": per_gpu * world, "n_vec": 36, "steps": steps, "wall_s": wall,
                          "member_steps_per_s_end_to_end": per_gpu * world * steps / wall,
                          "leading_exponents": [float(v) for v in mean[:4]],
                          "standard_errors": [float(v) for v in sem[:4]], "sum_of_exponents": float(mean.sum())}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
