"""N>1 path on CPU: two gloo ranks shard a batch of reads, decode their shards independently (the
CPU oracle stands in for the GPU decode here -- tests may use it as the checker/decoder), and rank
0 gathers the lists in input order.  No data-path collective is involved."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from nanopore_dna_storage_b200 import synth
    from nanopore_dna_storage_b200.sharding import decode_sharded, plan
    from oracle import binding as ob
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synth.Workload("shard", 6, 1, 40, 2, 20, sub=0.01, dele=0.01, ins=0.005, seed=31)
    reads = synth.make_reads(w, 7)
    posts = [r.post for r in reads]

    def decode_fn(ps):
        return [ob.oracle_decode(p, 6, 1, 40, 2, 20, num_thr=1)[1].tolist() for p in ps]

    out = decode_sharded(posts, decode_fn, rank, world)
    mine = plan([len(p) for p in posts], world)[rank]
    dist.barrier()
    if rank == 0:
        want = decode_fn(posts)
        q.put((out == want, len(mine)))
    else:
        q.put((out is None, len(mine)))
    dist.destroy_process_group()


def _gpu_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from nanopore_dna_storage_b200 import synth
    from nanopore_dna_storage_b200.decoder import ListDecoder
    from nanopore_dna_storage_b200.sharding import decode_sharded
    from oracle import binding as ob
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synth.Workload("shard", 8, 5, 40, 4, 20, sub=0.01, dele=0.01, ins=0.005, seed=32)
    reads = synth.make_reads(w, 11)
    posts = [r.post for r in reads]
    dev = rank % torch.cuda.device_count()   # one GPU per rank where the box has them
    with ListDecoder(8, 5, 40, 4, 20, device=dev) as dec:
        out = decode_sharded(posts, lambda ps: [b.tolist() for b, _ in dec.decode(ps)], rank, world)
    dist.barrier()
    if rank == 0:
        want = [ob.oracle_decode(p, 8, 5, 40, 4, 20)[1].tolist() for p in posts]
        q.put((out == want, dev))
    else:
        q.put((out is None, dev))
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.gpu
def test_sharded_decode_with_the_cuda_decoder():
    """decode_sharded with the real ListDecoder: one rank per GPU (both ranks on GPU 0 when the box has
    one), lists gathered on rank 0 in input order and compared with the oracle."""
    import torch
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_gpu_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for ok, _ in res)
    if torch.cuda.device_count() >= 2:
        assert {d for _, d in res} == {0, 1}


def test_two_rank_sharded_decode_gathers_in_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for ok, _ in res)
    assert sum(n for _, n in res) == 7 and all(n >= 3 for _, n in res)
