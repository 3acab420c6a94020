"""Run-time selectors that replace the reference's ``--platform`` / ``--jax-precision``
switches (``sushie/cli.py:2051-2055``): which GPU(s) to use and the device storage
precision of genotypes / LD (arithmetic is always fp64)."""
from __future__ import annotations

import os
from typing import List, Optional

# 64 -> fp64 storage (parity mode, the reference CLI default); 32 -> fp32 storage.
precision: int = int(os.environ.get("SUSHIE_B200_PRECISION", "64"))
# Device storage of individual-level genotypes: "auto" keeps hard calls (integer-valued, int8 range,
# not residualised on covariates) as one byte per entry and folds centring / scaling into the
# vectors (8x less HBM traffic than fp64, same fp64 arithmetic); "dense" always stores the
# standardised matrix at `precision`.
genotype_storage: str = os.environ.get("SUSHIE_B200_GENO", "auto")
# hard calls that are only 0 / 1 / 2 use the bit-plane traversal (table look-ups); 0 keeps them on the one-byte path
bit_planes: bool = bool(int(os.environ.get("SUSHIE_B200_BIT_PLANES", "1")))
# GPU ordinals used by the batch entry points; None = device 0 only.
devices: Optional[List[int]] = None
# thread blocks cooperating on one locus; 0 lets the engine choose.
team_size: int = int(os.environ.get("SUSHIE_B200_TEAM", "0"))
# keep one team size for a whole batch (disables the suspend / regroup phases)
single_phase: bool = bool(int(os.environ.get("SUSHIE_B200_SINGLE_PHASE", "0")))
# loci per device-resident chunk in the batch entry points (0 = size by memory budget).
chunk_loci: int = int(os.environ.get("SUSHIE_B200_CHUNK", "0"))
# automatic chunking with several GPU workers (chunk_loci = 0): at least one chunk per worker, up to
# `chunks_per_worker` per worker while a chunk keeps `min_chunk_loci` loci (two per SM)
chunks_per_worker: int = int(os.environ.get("SUSHIE_B200_CHUNKS_PER_WORKER", "3"))
min_chunk_loci: int = int(os.environ.get("SUSHIE_B200_MIN_CHUNK", "296"))
# worker threads (each with its own sb_ctx and resident batch) per GPU in the batch entry points.  They take turns in
# the IBSS kernel (`_capi.device_run_lock`); while one is in it, the others do the host half of their previous chunk
# (download, sub-sampling, results) and upload their next one, so a chunk is always ready when the kernel ends --
# three, because a worker's own epilogue -> prologue chain is serial and two workers end up in it together
workers_per_device: int = int(os.environ.get("SUSHIE_B200_WORKERS_PER_DEVICE", "3"))
# SMs the IBSS kernel leaves free (sb_opts.reserve_sms; default 0 = the whole device).  Measured on config-5 chunks:
# the other workers' small kernels (column statistics, set selection, purity) are ~8 % of a chunk's device time on
# the WHOLE GPU, so they crawl on a handful of reserved SMs; letting them queue behind the running IBSS kernel and
# fill its tail is faster.  Kept as a switch for latency-sensitive mixes.
reserve_sms: int = int(os.environ.get("SUSHIE_B200_RESERVE_SMS", "0"))
# print the wall time of the host-side phases of every chunk (upload, run, download, purity, tables) to stderr
trace_host: bool = bool(int(os.environ.get("SUSHIE_B200_TRACE_HOST", "0")))
# device memory budget per chunk, bytes
chunk_bytes: int = int(float(os.environ.get("SUSHIE_B200_CHUNK_GB", "24")) * (1 << 30))


def reserve_for(device: int) -> int:
    """``sb_opts.reserve_sms`` for a chunk about to run on ``device``."""
    return max(0, reserve_sms)


def storage_code() -> int:
    from . import _capi

    if precision == 64:
        return _capi.SB_F64
    if precision == 32:
        return _capi.SB_F32
    raise ValueError(f"sushie_b200.config.precision must be 32 or 64, got {precision}")
