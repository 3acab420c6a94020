"""``python -m sushie_b200 finemap ...`` == the reference's ``sushie finemap ...`` console script."""
import sys

from .cli import run_cli

if __name__ == "__main__":  # worker processes of `finemap-batch` re-import this module
    sys.exit(run_cli())
