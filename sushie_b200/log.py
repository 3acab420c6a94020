"""Logger shared by the package (same name as the reference's: ``sushie/log.py:3``)."""
import logging

logger = logging.getLogger("sushie")
