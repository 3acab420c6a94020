"""sushie_b200 -- B200-native SuShiE: the IBSS variational-inference hot path of
mancusolab/sushie behind the reference's own Python API (``infer_sushie``,
``infer_sushie_ss``), running as hand-written sm_100a CUDA behind a C ABI."""
from . import config, log, utils  # noqa: F401
from . import infer, infer_ss  # noqa: F401
from . import cli, io  # noqa: F401  (same sub-modules as the reference package: cli, infer, io, log, utils)
from .infer import infer_sushie, infer_sushie_batch, make_cs  # noqa: F401
from .infer_ss import infer_sushie_ss, infer_sushie_ss_batch  # noqa: F401

__version__ = "0.1.0"
__all__ = [
    "config", "cli", "infer", "infer_ss", "io", "log", "utils",
    "infer_sushie", "infer_sushie_batch", "infer_sushie_ss", "infer_sushie_ss_batch", "make_cs",
]
