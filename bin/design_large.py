#!/usr/bin/env python3
"""design.py with the 'large' defaults of the reference's bin/design_large.py:19-21
(-m 5, -e 50, --filter-with-lsh-minhash 0.6)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import design  # noqa: E402

if __name__ == '__main__':
    design.main(design.init_and_parse_args('large'))
