#!/bin/bash
# last confirm of the round: full GPU suite, then the driver's default bench command
OUT=gpurun_out; mkdir -p $OUT
timeout 110 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $OUT/r02aj_tests.txt
timeout 85 python bench.py --steps 5 --warmup 3 > $OUT/r02aj_default.json 2> $OUT/r02aj_default.err
tail -c 600 $OUT/r02aj_default.json | tr ',' '\n' | grep -E '"value"|frac' | head -8
