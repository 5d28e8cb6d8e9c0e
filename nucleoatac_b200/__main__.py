"""`python -m nucleoatac_b200 <occ|vprocess|nuc|merge|nfr|run> ...` (nucleoatac/cli.py)."""
import sys

from .cli import nucleoatac_main

if __name__ == "__main__":   # worker processes (fuzz.py) re-import the main module under another name: they must not run the CLI
    sys.exit(nucleoatac_main())
