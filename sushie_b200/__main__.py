"""``python -m sushie_b200 finemap ...`` == the reference's ``sushie finemap ...`` console script."""
import sys

from .cli import run_cli

sys.exit(run_cli())
