"""World-size-2 test (gloo, CPU) of the multi-GPU path's host logic: contiguous sharding of a batch across
ranks, per-rank fitting, and the final host gather.  No collective is on the data path (fits are independent);
the gather below is what a multi-process caller does with the per-rank results.  The per-rank "fit" is the
host build of the kernel core (oracle/host_emu) because there is no GPU here; on the GPU box the same script
shape is what bench.py runs under torchrun."""

import os
import socket
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, REPO)
    import torch
    import torch.distributed as dist

    from dphtools_b200 import synth
    from dphtools_b200.lm import shard_bounds
    from oracle import host_emu

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synth.WORKLOADS["c3_gauss2d_7"](n, seed=21)
    lo, hi = shard_bounds(n, world)[rank]
    out = host_emu.fit_batch(0, "mle", w["data"][lo:hi], w["p0"][lo:hi])
    # final gather of the (ragged) shards on rank 0
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((lo, hi, out["popt"], out["info"]), gathered, dst=0)
    # whole-job throughput bookkeeping as in bench.py: max over ranks of the elapsed time
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        popt = np.empty((n, 5), np.float32)
        info = np.empty(n, np.int32)
        for a, b, p, i in gathered:
            popt[a:b], info[a:b] = p, i
        q.put((popt, info, float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 64])
def test_two_rank_sharding_equals_single_process(n):
    import torch.multiprocessing as mp

    from dphtools_b200 import synth
    from dphtools_b200.lm import shard_bounds
    from oracle import host_emu

    assert shard_bounds(10, 3) == [(0, 3), (3, 6), (6, 10)] and shard_bounds(2, 4) == [(0, 0), (0, 1), (1, 1), (1, 2)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    popt, info, tmax = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = synth.WORKLOADS["c3_gauss2d_7"](n, seed=21)
    ref = host_emu.fit_batch(0, "mle", w["data"], w["p0"])
    np.testing.assert_array_equal(popt, ref["popt"])
    np.testing.assert_array_equal(info, ref["info"])
    assert tmax == 2.0
