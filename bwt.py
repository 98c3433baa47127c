#!/usr/bin/env python3
"""``python bwt.py ref.fa [flags]`` -- same CLI as the reference, GPU hot path.
Importing this file as ``bwt`` gives the reference's module surface."""
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import bwt_algorithm_b200  # noqa: E402,F401
from bwt_algorithm_b200.bwt import *  # noqa: E402,F401,F403
from bwt_algorithm_b200.bwt import (  # noqa: E402,F401
    _count_equal_range, _kasai_lcp_uint8, _natural_sort_key, _process_chromosome_worker, main)

if __name__ == "__main__":
    main()
