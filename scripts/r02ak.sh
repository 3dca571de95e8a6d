#!/bin/bash
# HEAD sanity after the source tidy-up: smoke + the 1-D parity subset
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 40 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "burgers_1d_step or pairs_are_independent or nyquist" 2>&1 | tail -2
