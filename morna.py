#!/usr/bin/env python
"""Drop-in entry point: `python morna.py index ...` / `python morna.py search ...`
with the reference's flags (see morna_b200/cli.py)."""
import sys

from morna_b200.cli import main

if __name__ == "__main__":
    sys.exit(main())
