#!/bin/bash
o=gpurun_out
timeout 600 python -m pytest tests/test_gr_glue.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --windows 600 --overlap-windows 3000 --array-channels 8 --array-windows 6 --steps 2 --warmup 1 \
   --verify-windows 200 --verify-full-jiggle 50 --verify-overlap-windows 100 --ref-windows 64 > $o/r2_bench_small.json 2> $o/r2_bench_small.err
echo "small bench exit $?"; tail -5 $o/r2_bench_small.err; head -c 9000 $o/r2_bench_small.json
( time timeout 1500 python bench.py > $o/r2_bench_full.json 2> $o/r2_bench_full.err ) 2>&1 | tail -3
echo "full bench exit $?"; tail -5 $o/r2_bench_full.err
