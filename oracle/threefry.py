"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the JAX PRNG pieces SuShiE touches.

The reference draws the purity sub-sample with
``jax.random.choice(PRNGKey(seed), snp_idx, shape=(max_select,), replace=False)``
(reference ``sushie/infer.py:894`` and ``sushie/infer.py:940-943``) and shuffles
cross-validation rows the same way (``sushie/cli.py:329-338``).  The arithmetic lives
in the third-party dependency ``jax==0.4.33`` (``setup.cfg:35-40``), which is not
vendored under ``/root/reference`` and is not installable here, so this file restates
its *published* algorithm:

* ``threefry2x32`` (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
  SC'11; 20 rounds, rotation constants 13/15/26/6 and 17/29/16/24, key-schedule
  parity constant 0x1BD11BDA),
* ``PRNGKey(seed)`` = ``(seed >> 32, seed & 0xFFFFFFFF)``,
* ``split`` = threefry over ``iota(2*num)``, reshaped ``(num, 2)``,
* ``random_bits(32)`` = threefry over ``iota(size)`` (non-partitionable layout: the
  counter vector is split into a first and a second half that form the two lanes),
* ``permutation`` of a 1-D array = ``ceil(3*log(size)/log(2**32-1))`` rounds of
  "split the key, draw 32 random bits per element, stable sort by them",
* ``choice(replace=False)`` = first ``n`` entries of ``permutation``.

Pinned by: the Random123 known-answer vectors for threefry2x32 and the
``split(PRNGKey(0))`` values printed in the JAX documentation
(``tests/test_oracle.py``).  Parity with the real ``jax.random.choice`` is otherwise unpinned (JAX cannot be
imported in this environment); ``tests/refshim`` uses this module where the reference's source calls ``jax.random``.
"""
from __future__ import annotations

import math

import numpy as np

_ROT_A = (13, 15, 26, 6)
_ROT_B = (17, 29, 16, 24)
_M32 = np.uint64(0xFFFFFFFF)


def _rotl(x: np.ndarray, d: int) -> np.ndarray:
    x = x & _M32
    return ((x << np.uint64(d)) | (x >> np.uint64(32 - d))) & _M32


def threefry2x32(key, x0, x1):
    """Threefry-2x32, 20 rounds.  ``key`` is a pair of uint32; x0/x1 are uint32 arrays."""
    k0 = np.uint64(int(key[0]) & 0xFFFFFFFF)
    k1 = np.uint64(int(key[1]) & 0xFFFFFFFF)
    k2 = (k0 ^ k1 ^ np.uint64(0x1BD11BDA)) & _M32
    ks = (k0, k1, k2)
    a = (np.asarray(x0, dtype=np.uint64) + k0) & _M32
    b = (np.asarray(x1, dtype=np.uint64) + k1) & _M32
    for block in range(5):
        rots = _ROT_A if block % 2 == 0 else _ROT_B
        for r in rots:
            a = (a + b) & _M32
            b = _rotl(b, r)
            b = a ^ b
        a = (a + ks[(block + 1) % 3]) & _M32
        b = (b + ks[(block + 2) % 3] + np.uint64(block + 1)) & _M32
    return a.astype(np.uint32), b.astype(np.uint32)


def _threefry_counts(key, counts: np.ndarray) -> np.ndarray:
    """jax ``threefry_2x32(key, count)``: odd sizes are zero-padded, the flat counter
    vector is cut in two halves that become the two lanes, outputs are concatenated."""
    counts = np.asarray(counts, dtype=np.uint32).ravel()
    odd = counts.size % 2
    if odd:
        counts = np.concatenate([counts, np.zeros(1, dtype=np.uint32)])
    half = counts.size // 2
    o0, o1 = threefry2x32(key, counts[:half], counts[half:])
    out = np.concatenate([o0, o1])
    return out[:-1] if odd else out


def prng_key(seed: int):
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def split(key, num: int = 2) -> np.ndarray:
    return _threefry_counts(key, np.arange(2 * num, dtype=np.uint32)).reshape(num, 2)


def random_bits32(key, size: int) -> np.ndarray:
    return _threefry_counts(key, np.arange(size, dtype=np.uint32))


def permutation(key, values: np.ndarray) -> np.ndarray:
    x = np.asarray(values).copy()
    size = x.size
    rounds = int(np.ceil(3 * np.log(max(1, size)) / np.log(np.iinfo(np.uint32).max)))
    for _ in range(rounds):
        key, sub = split(key, 2)
        bits = random_bits32(sub, size)
        x = x[np.argsort(bits, kind="stable")]
    return x


def choice_without_replacement(key, values: np.ndarray, n_draws: int) -> np.ndarray:
    if n_draws > np.asarray(values).size:
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    return permutation(key, values)[:n_draws]


__all__ = [
    "threefry2x32",
    "prng_key",
    "split",
    "random_bits32",
    "permutation",
    "choice_without_replacement",
]
_ = math  # keep the import list explicit for readers diffing against the JAX source
