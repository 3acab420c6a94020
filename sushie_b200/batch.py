"""Host work queue that shards independent fits over the GPUs of one box.

Every locus (and every cross-validation fold / meta / mega fit of a locus) is an
independent IBSS problem -- the reference parallelises only by running one process per gene
(``misc/run_sushie.sh``) -- so there is no data-path collective: problems are cut into
chunks (cost-sorted, longest first), one worker per GPU pulls chunks from a shared queue
until it is empty, and results are put back in input order.  A locus is never split across
GPUs.  Under ``torchrun`` (one process per GPU: ``run_distributed``, used by ``bench.py``'s config-5 arm and
``scripts/bench_cv.py``) every rank plans the same global chunk list and the queue head is an atomic counter in
the ``torch.distributed`` store (``StoreQueue``) instead of a thread lock.
"""
from __future__ import annotations

import queue
import threading
from typing import Callable, Dict, List, Optional, Sequence, Tuple

from . import _capi, config


def estimate_bytes(prep) -> int:
    """Device bytes one prepared problem needs (storage + state + double-buffered outputs)."""
    pr = prep.problem
    es = 8 if config.precision == 64 else 4
    rows = int(sum(pr.n))
    p_pad = (pr.p + 31) // 32 * 32
    mats = rows * p_pad * es
    state = 8 * (rows * (pr.L + 3) + pr.k * p_pad * 24)
    out = 2 * 8 * pr.L * pr.p * (2 + pr.k + pr.k * pr.k)
    staging = int(max(pr.n)) * pr.p * 8  # pr.n is an int32 array: keep the sum in Python integers
    return int(mats + state + out + staging)


def plan_chunks(preps: Sequence, chunk_bytes: Optional[int] = None, chunk_loci: Optional[int] = None,
                n_workers: int = 1, n_devices: Optional[int] = None) -> List[List[int]]:
    """Group problem indices into chunks.

    Problems that cannot share a launch configuration (different ``max_iter`` /
    ``min_tol``, or hard-call vs dense genotype storage) never share a chunk.  Within a group, indices are sorted by descending
    cost (longest-processing-time first) and cut where the memory budget or the locus
    cap is reached.  With several workers and no explicit cap (``chunk_loci`` = 0) a group is cut into at least one
    chunk per DEVICE (``n_devices``; default: one device per worker), and into up to ``config.chunks_per_worker`` per
    worker while the chunks keep at least ``config.min_chunk_loci`` loci (two per SM).  The workers of one GPU take
    turns in the IBSS kernel, so cutting a small group into one chunk per worker would only run its loci one after
    the other instead of side by side on the SMs: chunks are for pipelining the host half of big batches, and a
    group below ``2 * min_chunk_loci`` loci per device stays whole."""
    chunk_bytes = config.chunk_bytes if chunk_bytes is None else chunk_bytes
    chunk_loci = config.chunk_loci if chunk_loci is None else chunk_loci
    groups: Dict[Tuple, List[int]] = {}
    for i, p in enumerate(preps):
        # loci eligible for one-byte hard-call storage get chunks of their own (storage is per batch)
        groups.setdefault((p.max_iter, p.min_tol, getattr(p, "hard_calls", False)), []).append(i)
    chunks: List[List[int]] = []
    for ids in groups.values():
        ids = sorted(ids, key=lambda i: (-preps[i].problem.cost, i))
        # target chunk sizes: an explicit cap, or -- several workers, no cap -- at least one chunk per worker,
        # more (for balance) only while the chunks stay large enough; sizes differ by at most one locus
        caps: List[int] = []
        if chunk_loci:
            caps = [chunk_loci]
        elif n_workers > 1:
            n_dev = n_workers if n_devices is None else max(1, n_devices)
            n_cut = max(n_dev, min(config.chunks_per_worker * n_workers, len(ids) // max(1, config.min_chunk_loci)))
            n_cut = min(n_cut, len(ids))
            caps = [len(ids) // n_cut + (1 if c < len(ids) % n_cut else 0) for c in range(n_cut)]
        cur: List[int] = []
        cur_bytes = 0
        n_done = 0  # chunks closed in this group (indexes the target sizes)
        for i in ids:
            b = estimate_bytes(preps[i])
            cap = caps[min(n_done, len(caps) - 1)] if caps else 0
            if cur and (cur_bytes + b > chunk_bytes or (cap and len(cur) >= cap)):
                chunks.append(cur)
                cur, cur_bytes = [], 0
                n_done += 1
            cur.append(i)
            cur_bytes += b
        if cur:
            chunks.append(cur)
    # biggest chunks first so the tail of the queue is made of small work items
    chunks.sort(key=lambda c: -sum(preps[i].problem.cost for i in c))
    return chunks


class ChunkQueue:
    """Thread-safe queue head over a fixed chunk list (the in-process work queue)."""

    def __init__(self, n_chunks: int):
        self._q: "queue.SimpleQueue[int]" = queue.SimpleQueue()
        for c in range(n_chunks):
            self._q.put(c)

    def next(self) -> Optional[int]:
        try:
            return self._q.get_nowait()
        except queue.Empty:
            return None


class StoreQueue:
    """Queue head shared by the ranks of a ``torch.distributed`` job: an atomic counter
    in the rendezvous store (no tensors, no collective on the data path)."""

    def __init__(self, store, n_chunks: int, key: str = "sushie_b200/chunk_head"):
        self._store, self._n, self._key = store, n_chunks, key

    def next(self) -> Optional[int]:
        c = int(self._store.add(self._key, 1)) - 1
        return c if c < self._n else None


class _Finisher:
    """One background thread per GPU worker for the host-only half of the epilogue (output tables).

    A worker hands over the deferred results of a chunk and goes straight back to the queue; while it sits
    in the next chunk's engine call (the GIL is released inside the C ABI) this thread assembles the tables
    of the previous one, so table assembly costs wall time only where it exceeds the device time."""

    def __init__(self, results: list):
        self._results = results
        self._q: "queue.SimpleQueue" = queue.SimpleQueue()
        self._error: Optional[BaseException] = None
        self._thread = threading.Thread(target=self._work, daemon=True)
        self._thread.start()

    def _work(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            ids, deferred = item
            if self._error is not None:
                continue
            try:
                for i, finish in zip(ids, deferred):
                    self._results[i] = finish()
            except BaseException as exc:  # surfaced by close()
                self._error = exc

    def submit(self, ids: List[int], deferred: list) -> None:
        self._q.put((ids, deferred))

    def close(self) -> None:
        self._q.put(None)
        self._thread.join()
        if self._error is not None:
            raise self._error


def drain(q, chunks: List[List[int]], preps: Sequence, runner: Callable, device: int, results: list, log: list,
          defer: bool = False) -> None:
    """Worker loop: pull chunks until the queue is empty.  With ``defer`` the runner returns the host-only
    half of every result as a callable, which a finisher thread runs while this loop fetches the next chunk."""
    finisher = _Finisher(results) if defer else None
    try:
        while True:
            c = q.next()
            if c is None:
                break
            ids = chunks[c]
            if defer:
                finisher.submit(ids, runner([preps[i] for i in ids], device, defer=True))
            else:
                for i, r in zip(ids, runner([preps[i] for i in ids], device)):
                    results[i] = r
            log.append((device, c, len(ids)))
    finally:
        if finisher is not None:
            finisher.close()


def run_sharded(preps: Sequence, runner: Callable, devices: Optional[Sequence[int]] = None) -> list:
    """Run every prepared problem; one worker thread (one ``sb_ctx``) per device, each with a finisher
    thread for the output tables."""
    if len(preps) == 0:
        return []
    if devices is None:
        devices = config.devices or [0]
    devices = list(devices)
    # several worker threads per GPU: the host half of one chunk overlaps the device half of the next
    workers = [d for d in devices for _ in range(max(1, config.workers_per_device))]
    chunks = plan_chunks(preps, n_workers=len(workers), n_devices=len(devices))
    q = ChunkQueue(len(chunks))
    results: list = [None] * len(preps)
    trace: list = []
    defer = len(chunks) > 1  # a single chunk has nothing to overlap with
    if len(workers) == 1 or len(chunks) == 1:
        drain(q, chunks, preps, runner, devices[0], results, trace, defer)
        return results
    errors: list = []

    def work(dev):
        try:
            drain(q, chunks, preps, runner, dev, results, trace, defer)
        except BaseException as exc:  # surfaced after join
            errors.append(exc)
        finally:
            _capi.release_thread_engines()  # this thread ends here: do not leak its sb_ctx

    threads = [threading.Thread(target=work, args=(d,), daemon=True) for d in workers[: len(chunks)]]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


def plan_items(costs: Sequence[float], n_workers: int, min_items: int = 8, max_items: int = 64) -> List[List[int]]:
    """Cut work items (genes) into chunks for dynamic pulling: items sorted by descending cost, chunk sizes that
    shrink as the queue drains (guided self-scheduling: about 1 / (2 n_workers) of what is left, between
    ``min_items`` and ``max_items``), so the GPUs finish together however uneven the fits are.  Every rank computes
    the same list from the same costs."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    chunks: List[List[int]] = []
    pos = 0
    while pos < len(order):
        left = len(order) - pos
        size = max(min_items, min(max_items, -(-left // (2 * max(1, n_workers)))))
        chunks.append(order[pos: pos + size])
        pos += size
    return chunks


def run_distributed(costs: Sequence[float], process_chunk: Callable[[List[int]], list], *, store=None, rank: int = 0,
                    world: int = 1, workers_per_rank: int = 2, key: str = "sushie_b200/items", min_items: int = 8,
                    max_items: int = 64):
    """One process per GPU (``torchrun``): every rank plans the same global chunk list and pulls chunk indices from
    an atomic counter in the ``torch.distributed`` store (``StoreQueue``) -- a host work queue, no collective on the
    data path.  ``process_chunk(item_ids)`` does everything for those items on this rank's GPU (loading, fits,
    epilogue) and returns one small record per item.  ``workers_per_rank`` threads pull concurrently so that the host
    half of one chunk (set selection, tables, scores) overlaps the device half of the next.

    Returns ``(records, trace)``: this rank's ``(item, record)`` pairs and the chunks it pulled."""
    chunks = plan_items(costs, world, min_items, max_items)
    q = StoreQueue(store, len(chunks), key) if store is not None else ChunkQueue(len(chunks))
    records: list = []
    trace: list = []
    errors: list = []
    lock = threading.Lock()
    # Look-ahead of ONE chunk per rank: a worker may pull from the global queue only when no chunk of this rank is
    # still waiting for the GPU (the gate is handed on when a chunk enters the IBSS kernel).  Every worker pulling
    # as soon as it is free left up to `workers_per_rank` - 1 chunks parked behind one GPU at the end of a run while
    # other ranks had run dry (8 GPUs: the slowest rank did 1.22x the mean number of genes).
    gate = threading.Semaphore(1)

    def work():
        try:
            while True:
                gate.acquire()
                handed_on = [False]

                def hand_on():
                    if not handed_on[0]:
                        handed_on[0] = True
                        gate.release()

                _capi.run_hooks.on_start = hand_on
                try:
                    c = q.next()
                    if c is None:
                        break
                    out = process_chunk(chunks[c])
                finally:
                    hand_on()  # (also when the chunk never reached the GPU)
                    _capi.run_hooks.on_start = None
                with lock:
                    records.extend(zip(chunks[c], out))
                    trace.append((rank, c, len(chunks[c])))
        except BaseException as exc:  # surfaced after join
            errors.append(exc)
        finally:
            _capi.release_thread_engines()

    threads = [threading.Thread(target=work, daemon=True) for _ in range(max(1, workers_per_rank))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return records, trace


def device_count() -> int:
    return _capi.device_count()
