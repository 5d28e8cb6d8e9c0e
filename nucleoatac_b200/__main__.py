import sys

from .cli import nucleoatac_main

sys.exit(nucleoatac_main())
