#!/bin/bash
# One-GPU measurement set of a round: bench lines, reference arm, launch list, full ncu capture.
# Usage (on the GPU box, from the repo root): bash tools/final_profile.sh r1
tag=${1:-r1}
o=gpurun_out
timeout 400 python bench.py > $o/${tag}_bench_1gpu.json 2> $o/${tag}_bench_1gpu.err
timeout 400 python bench.py --impl reference > $o/${tag}_bench_1gpu_reference_arm.json 2> $o/${tag}_ref.err
timeout 300 python bench.py --overlap --no-cpu-baseline > $o/${tag}_bench_1gpu_overlap.json 2> $o/${tag}_ovl.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -s 12 -c 4 -f -o $o/${tag}_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_full.log 2>&1
ls -la $o/${tag}_*
