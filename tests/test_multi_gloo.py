"""N > 1 host path on CPU: world_size 2 over gloo.  Each rank owns its carriers (no data-path
collective); only timings and the tiny per-carrier results cross ranks.  The per-carrier engine
here is the CPU oracle standing in for the GPU kernels (the host logic is what is under test)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gr_amps_b200 import multi, synth
    from tests import oracle_lib as O
    plan = multi.carrier_plan(world, 3)                 # 3 carriers on 2 ranks: rank 0 gets 0 and 2
    mine = plan[rank]
    results, samples = [], 0
    for c in mine:
        x, hs, _ = synth.config2_period(n_total=55 * 38400, snr_db=20.0, seed=c.seed, center=c.center_freq, min10=c.min10)
        _, d = O.rx_chain_f32(x, center=c.center_freq)
        b = O.rx_detect(d)
        r = O.recc_decode(b[0][2])
        results.append((c.index, r.min.decode(), len(b)))
        samples += len(x)
    dist.barrier()
    total, tmax = multi.whole_job_throughput(samples, 1.0 + rank)       # pretend rank r took 1+r seconds
    allres = multi.gather_results(results)
    if rank == 0:
        q.put((total, tmax, allres))
    dist.destroy_process_group()


def test_two_ranks_independent_carriers():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    total, tmax, allres = q.get()
    assert total == 3 * 55 * 38400 and tmax == 2.0
    flat = sorted(x for per_rank in allres for x in per_rank)
    assert flat == [(0, "2125551230", 1), (1, "2125551231", 1), (2, "2125551232", 1)]


def test_carrier_plan():
    from gr_amps_b200 import multi
    plan = multi.carrier_plan(8)
    assert [len(p) for p in plan] == [1] * 8
    assert [p[0].center_freq for p in plan] == [-160e3 + 30e3 * g for g in range(8)]
    plan = multi.carrier_plan(4, 10)
    assert [[c.index for c in p] for p in plan] == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]
    assert multi.max_over_ranks(3.5) == 3.5 and multi.whole_job_throughput(10, 2.0) == (10.0, 2.0)
