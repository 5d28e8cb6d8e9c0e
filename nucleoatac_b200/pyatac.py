"""`python -m nucleoatac_b200.pyatac <vplot|ins|cov|bias|sizes> ...` -- the pyatac tools either side of the scoring path."""
import sys

from .pyatac_tools import pyatac_main

if __name__ == "__main__":
    sys.exit(pyatac_main())
