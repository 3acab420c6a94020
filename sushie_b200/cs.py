"""Credible sets, purity and the output tables (reference ``make_cs``, ``infer.py:835-1008``).

On the inference path the selection (sort by alpha, running sum, cut at ``threshold``) is computed on the device
(``sb_batch_select``) and arrives here as ``selection_from_device``; the stand-alone ``make_cs`` entry point, which
starts from host arrays, selects in numpy (``select_sets``).  The purity Gram matrices are computed on the GPU from
the device-resident genotypes (or LD) through ``sb_batch_purity_many``.  The ``max_select`` sub-sample reproduces
``jax.random.choice(PRNGKey(seed), idx, (max_select,), replace=False)`` (threefry2x32, same un-split key for every
effect, ``infer.py:894,940-943``).  The two pandas tables are assembled by ``frames`` -- lazily, on first access of
``SushieResult.cs`` / ``.alphas``.
"""
from __future__ import annotations

import functools
from typing import Callable, List, Tuple

import numpy as np
import pandas as pd

from . import log
from .utils import make_pip

try:  # pandas >= 2.2: a frame from ready-made (dtype-homogeneous, transposed) blocks without sanitising every column
    from pandas.api.internals import create_dataframe_from_blocks as _FRAME_FROM_BLOCKS
except ImportError:  # pragma: no cover - older pandas: the dict constructor
    _FRAME_FROM_BLOCKS = None

# ---- threefry2x32 counter PRNG (jax.random's default implementation) -------------------

_U32 = np.uint32
_ROTATIONS = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rol(v: np.ndarray, s: int) -> np.ndarray:
    return (v << _U32(s)) | (v >> _U32(32 - s))


def _threefry_pairs(k0: int, k1: int, c0: np.ndarray, c1: np.ndarray):
    with np.errstate(over="ignore"):
        ks = [_U32(k0), _U32(k1), _U32(k0) ^ _U32(k1) ^ _U32(0x1BD11BDA)]
        x0 = c0.astype(_U32) + ks[0]
        x1 = c1.astype(_U32) + ks[1]
        for i in range(5):
            for s in _ROTATIONS[i & 1]:
                x0 = x0 + x1
                x1 = _rol(x1, s) ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + _U32(i + 1)
    return x0, x1


def _threefry_stream(key: Tuple[int, int], n: int) -> np.ndarray:
    """``n`` 32-bit words for counters 0..n-1 in JAX's layout (first half / second half)."""
    cnt = np.arange(n + (n & 1), dtype=_U32)
    cnt[n:] = 0
    h = cnt.size // 2
    a, b = _threefry_pairs(key[0], key[1], cnt[:h], cnt[h:])
    return np.concatenate([a, b])[:n]


@functools.lru_cache(maxsize=1024)
def _split2(key: Tuple[int, int]):
    w = _threefry_stream(key, 4)
    return (int(w[0]), int(w[1])), (int(w[2]), int(w[3]))


def prng_key(seed: int) -> Tuple[int, int]:
    """``jax.random.PRNGKey(seed)`` as a pair of 32-bit words."""
    return ((int(seed) >> 32) & 0xFFFFFFFF, int(seed) & 0xFFFFFFFF)


def split_key(key: Tuple[int, int]):
    """``jax.random.split(key, 2)`` -> (first, second)."""
    return _split2(key)


@functools.lru_cache(maxsize=4096)
def _shuffle_order(key: Tuple[int, int], n: int) -> np.ndarray:
    """Positions picked by JAX's sort-based shuffle of ``n`` items under ``key``: rounds of stable sorts by
    fresh 32-bit words.  It does not depend on the values, and ``make_cs`` re-uses one key for every effect
    of every locus (``infer.py:894,940-943``), so the orders are cached per (key, n) -- a genome-wide batch
    asks for the same few thousand sizes over and over (int32 entries: at most ~50 MB for 4096 sizes)."""
    order = np.arange(n, dtype=np.int32)
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))
    for _ in range(rounds):
        key, sub = _split2(key)
        order = order[np.argsort(_threefry_stream(sub, n), kind="stable")]
    order.setflags(write=False)
    return order


def shuffle_with_key(key: Tuple[int, int], values: np.ndarray) -> np.ndarray:
    """JAX's sort-based shuffle of ``values`` (``jax.random.permutation`` /
    ``choice(..., replace=False)``)."""
    values = np.asarray(values)
    return values[_shuffle_order((int(key[0]), int(key[1])), int(values.size))]


def jax_choice_no_replace(seed: int, values: np.ndarray, n_draws: int) -> np.ndarray:
    """First ``n_draws`` entries of JAX's sort-based shuffle of ``values``."""
    return shuffle_with_key(prng_key(seed), values)[:n_draws]


# ---- credible sets ----------------------------------------------------------------------


def select_sets(alpha: np.ndarray, threshold: float, max_select: int, seed: int):
    """Per effect: descending order, cumulative alpha, number of SNPs in the set and the
    (possibly sub-sampled) SNP indices used for purity."""
    n_l, p = alpha.shape
    # pandas ``sort_values(ascending=False)``: argsort of the reversed column, reversed -- all effects in one call
    rev = np.argsort(alpha[:, ::-1], axis=1, kind="quicksort")
    order_all = (p - 1 - rev)[:, ::-1]
    csum_all = np.cumsum(np.take_along_axis(alpha, order_all, axis=1), axis=1)
    n_rows = np.count_nonzero(csum_all < threshold, axis=1)
    orders, csums, sizes, purity_idx = [], [], [], []
    for ldx in range(n_l):
        n_row = int(n_rows[ldx])
        size = n_row if n_row == p else n_row + 1
        order = order_all[ldx]
        idx = order[:size].astype(np.int64)
        if idx.size > max_select:
            idx = jax_choice_no_replace(seed, idx, max_select)
        orders.append(order)
        csums.append(csum_all[ldx])
        sizes.append(size)
        purity_idx.append(idx)
    return orders, csums, sizes, purity_idx


def combine_purity(min_abs: np.ndarray, ns: np.ndarray, purity_method: str) -> np.ndarray:
    """min |corr| per (effect, ancestry) -> one purity per effect (``infer.py:952-965``)."""
    ns = np.asarray(ns, dtype=float).reshape(-1)
    if purity_method == "weighted":
        return min_abs @ (ns / ns.sum())
    if purity_method == "max":
        return min_abs.max(axis=1)
    if purity_method == "min":
        return min_abs.min(axis=1)
    raise ValueError(f"Invalid purity method {purity_method}. Choose from 'weighted', 'max', or 'min'.")


def build_tables(
    alpha: np.ndarray,
    log_bf: np.ndarray,
    ns: np.ndarray,
    min_abs_corr: Callable[[List[np.ndarray]], np.ndarray],
    threshold: float,
    purity: float,
    purity_method: str,
    max_select: int,
    seed: int,
):
    """Assemble ``cs``, ``alphas``, ``pip_all``, ``pip_cs`` exactly in the reference's schema.

    ``min_abs_corr(list_of_index_arrays) -> [n_sets, k]`` supplies the purity raw material
    (device kernel)."""
    if purity_method not in ("weighted", "max", "min"):
        raise ValueError(f"Invalid purity method {purity_method}. Choose from 'weighted', 'max', or 'min'.")
    alpha = np.asarray(alpha, dtype=float)
    selection = select_sets(alpha, threshold, max_select, seed)
    min_abs = min_abs_corr(selection[3])
    return finish_tables(alpha, log_bf, ns, selection, min_abs, purity, purity_method)


def selection_from_device(order: np.ndarray, csum: np.ndarray, sizes: np.ndarray, l_order, max_select: int, seed: int):
    """The tuple :func:`select_sets` returns, from the device's selection (``sb_batch_select``: SNP order by
    descending alpha, running sum, set size per effect in ENGINE effect order), put into the result's effect
    order ``l_order``; the ``max_select`` sub-sample of large sets is drawn here (cached threefry orders)."""
    l_order = np.asarray(l_order)
    if not np.array_equal(l_order, np.arange(l_order.size)):
        order, csum, sizes = order[l_order], csum[l_order], sizes[l_order]
    sizes = [int(v) for v in sizes]
    purity_idx = []
    for ldx, size in enumerate(sizes):
        idx = order[ldx, :size]
        if size > max_select:
            idx = jax_choice_no_replace(seed, idx, max_select)
        purity_idx.append(idx)
    return order, csum, sizes, purity_idx


def summarise_sets(alpha, ns, min_abs, purity: float, purity_method: str, selection=None):
    """Purity per effect, which sets are kept, and the two PIP vectors (``infer.py:952-985``) -- everything of
    ``make_cs`` except the two tables.  With ``selection`` the reference's log lines are emitted here."""
    n_l = alpha.shape[0]
    purities = combine_purity(np.asarray(min_abs, dtype=float).reshape(n_l, -1), ns, purity_method)
    kept = purities > purity
    pip_all = make_pip(alpha)
    pip_cs = make_pip(alpha[np.nonzero(kept)[0]])
    if selection is not None:
        orders, _, sizes, _ = selection
        members = [np.asarray(orders[ldx][: sizes[ldx]]) for ldx in range(n_l) if kept[ldx]]
        if members:
            snp_in_cs = np.concatenate(members)
            if len(snp_in_cs) != len(np.unique(snp_in_cs)):
                log.logger.warning(
                    "Same SNPs appear in different credible set, which is very unusual."
                    + " You may want to check this gene in details."
                )
        log.logger.info(
            f"{len(members)} out of {n_l} credible sets remain after pruning based on purity ({purity})."
            + " For detailed results, specify --alphas."
        )
    return purities, kept, pip_all, pip_cs


@functools.lru_cache(maxsize=64)
def _alphas_layout(n_l: int):
    """Column index of the ``alphas`` table for ``n_l`` effects and where its float64 / int64 columns sit (building
    the string index costs more than filling the frame, and it is the same for every locus)."""
    columns, floats, ints = ["SNPIndex"], [], ["SNPIndex"]
    for ldx in range(n_l):
        tag = f"l{ldx + 1}"
        columns += [f"alpha_{tag}", f"in_cs_{tag}", f"purity_{tag}", f"kept_{tag}", f"log_bf_{tag}"]
        floats += [f"alpha_{tag}", f"purity_{tag}", f"log_bf_{tag}"]
        ints += [f"in_cs_{tag}", f"kept_{tag}"]
    columns += ["pip_all", "pip_cs"]
    floats += ["pip_all", "pip_cs"]
    return _layout(columns, floats, ints)


def _layout(columns: List[str], floats: List[str], ints: List[str]):
    position = {name: i for i, name in enumerate(columns)}
    index = pd.Index(columns)
    return (index, tuple(columns), tuple(floats), np.asarray([position[c] for c in floats], dtype=np.intp),
            tuple(ints), np.asarray([position[c] for c in ints], dtype=np.intp))


_CS_LAYOUT = None


def _cs_layout():
    global _CS_LAYOUT
    if _CS_LAYOUT is None:
        _CS_LAYOUT = _layout(["CSIndex", "SNPIndex", "alpha", "c_alpha", "pip_all", "pip_cs"],
                             ["alpha", "c_alpha", "pip_all", "pip_cs"], ["CSIndex", "SNPIndex"])
    return _CS_LAYOUT


def _frame_from_blocks(layout, float_block: np.ndarray, int_block: np.ndarray) -> pd.DataFrame:
    """A frame of float64 and int64 columns in the layout's column order from its two ready-made blocks
    (``[columns of the dtype, rows]``; ``pandas.api.internals.create_dataframe_from_blocks``) -- the dict constructor
    spends ~1 ms per fit sanitising and consolidating 53 columns.  Older pandas: the dict constructor."""
    index, columns, floats, float_pos, ints, int_pos = layout
    n_rows = float_block.shape[1]
    if _FRAME_FROM_BLOCKS is None:
        merged = {name: float_block[r] for r, name in enumerate(floats)}
        merged.update({name: int_block[r] for r, name in enumerate(ints)})
        return pd.DataFrame({name: merged[name] for name in columns})
    return _FRAME_FROM_BLOCKS([(float_block, float_pos), (int_block, int_pos)], index=pd.RangeIndex(n_rows), columns=index)


def frames(alpha, log_bf, selection, purities, kept, pip_all, pip_cs):
    """The ``cs`` and ``alphas`` tables in the reference's schema (``infer.py:897-1008``) from the selected sets."""
    n_l, p = alpha.shape
    orders, csums, sizes, _ = selection
    layout = _alphas_layout(n_l)
    fb = np.empty((3 * n_l + 2, p))                    # alpha, purity, log_bf per effect; pip_all, pip_cs
    ib = np.zeros((2 * n_l + 1, p), dtype=np.int64)    # SNPIndex; in_cs, kept per effect
    ib[0] = np.arange(p)
    cs_parts = []
    for ldx in range(n_l):
        order, csum, size = orders[ldx], csums[ldx], sizes[ldx]
        fb[3 * ldx] = alpha[ldx]
        fb[3 * ldx + 1] = purities[ldx]
        fb[3 * ldx + 2] = log_bf[ldx]
        ib[2 * ldx + 1, order[:size]] = 1
        ib[2 * ldx + 2] = int(kept[ldx])
        if kept[ldx]:
            sel = np.asarray(order[:size])
            cs_parts.append((ldx + 1, sel, alpha[ldx][sel], csum[:size]))
    fb[3 * n_l] = pip_all
    fb[3 * n_l + 1] = pip_cs
    full_alphas = _frame_from_blocks(layout, fb, ib)

    # one frame for all kept sets, built in one go (a frame per effect, or adding columns afterwards, costs
    # more than the fit at batch throughput)
    n_rows = sum(len(part[1]) for part in cs_parts)
    cfb = np.empty((4, n_rows))                         # alpha, c_alpha, pip_all, pip_cs
    cib = np.empty((2, n_rows), dtype=np.int64)         # CSIndex, SNPIndex
    pos = 0
    for cs_index, sel, a_sel, c_sel in cs_parts:
        end = pos + len(sel)
        cib[0, pos:end] = cs_index
        cib[1, pos:end] = sel
        cfb[0, pos:end] = a_sel
        cfb[1, pos:end] = c_sel
        pos = end
    cfb[2] = pip_all[cib[1]]
    cfb[3] = pip_cs[cib[1]]
    cs = _frame_from_blocks(_cs_layout(), cfb, cib)
    return cs, full_alphas


def finish_tables(alpha, log_bf, ns, selection, min_abs, purity: float, purity_method: str):
    """Host-only second half of :func:`build_tables`: from the selected sets (``select_sets`` or the device's
    selection) and their min |corr| per ancestry to the output tables and PIPs."""
    alpha = np.asarray(alpha, dtype=float)
    log_bf = np.asarray(log_bf, dtype=float)
    if purity_method not in ("weighted", "max", "min"):
        raise ValueError(f"Invalid purity method {purity_method}. Choose from 'weighted', 'max', or 'min'.")
    purities, kept, pip_all, pip_cs = summarise_sets(alpha, ns, min_abs, purity, purity_method, selection)
    cs, full_alphas = frames(alpha, log_bf, selection, purities, kept, pip_all, pip_cs)
    return cs, full_alphas, pip_all, pip_cs
