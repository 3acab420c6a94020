"""N > 1 path on CPU: two gloo ranks pull chunks from the shared store-backed work queue
(the multi-GPU scheduler of bench.py / sushie_b200.batch) and together cover every locus
exactly once, with no tensor exchanged on the data path."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_chunks, out_q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from sushie_b200 import batch

    store = dist.TCPStore("127.0.0.1", port, world, rank == 0)
    dist.init_process_group("gloo", store=store, rank=rank, world_size=world)
    q = batch.StoreQueue(store, n_chunks)
    mine = []
    while (c := q.next()) is not None:
        mine.append(c)
    dist.barrier()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)  # results are gathered, inputs never move
    if rank == 0:
        out_q.put(gathered)
    dist.destroy_process_group()


def test_store_queue_two_ranks():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    port, n_chunks = _free_port(), 37
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_chunks, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = out_q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = sorted(c for part in gathered for c in part)
    assert flat == list(range(n_chunks))
    assert all(len(part) > 0 for part in gathered) or n_chunks < 2


def _dist_worker(rank, world, port, n_items, out_q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from sushie_b200 import batch

    store = dist.TCPStore("127.0.0.1", port, world, rank == 0)
    dist.init_process_group("gloo", store=store, rank=rank, world_size=world)
    costs = [float((7 * i) % 13) for i in range(n_items)]
    records, trace = batch.run_distributed(costs, lambda ids: [10 * i for i in ids], store=store, rank=rank, world=world,
                                           workers_per_rank=2, key="test/items", min_items=2, max_items=9)
    gathered = [None] * world
    dist.all_gather_object(gathered, (records, trace))
    if rank == 0:
        out_q.put(gathered)
    dist.destroy_process_group()


def test_run_distributed_two_ranks_cover_every_item_once():
    """`batch.run_distributed` (the config-5 arm of bench.py): two gloo ranks, two worker threads each, one global
    cost-sorted chunk list behind the store counter -- every item is processed exactly once, by some rank."""
    import torch.multiprocessing as mp

    from sushie_b200 import batch

    n_items = 101
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, n_items, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = out_q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    items = sorted(i for recs, _ in gathered for i, _ in recs)
    assert items == list(range(n_items))
    assert all(r == 10 * i for recs, _ in gathered for i, r in recs)
    chunks = sorted(c for _, tr in gathered for _, c, _ in tr)
    plan = batch.plan_items([float((7 * i) % 13) for i in range(n_items)], 2, 2, 9)
    assert chunks == list(range(len(plan)))
    # longest-processing-time order, chunk sizes shrink towards the end of the queue
    sizes = [len(c) for c in plan]
    assert sizes == sorted(sizes, reverse=True) and sizes[0] == 9 and sizes[-1] >= 1
    first_costs = [float((7 * i) % 13) for i in plan[0]]
    assert min(first_costs) >= max(float((7 * i) % 13) for i in plan[-1])
