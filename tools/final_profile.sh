#!/bin/bash
# One-GPU measurement set of a round: bench lines, reference arm, launch list, full ncu captures.
# Usage (on the GPU box, from the repo root): bash tools/final_profile.sh r2
tag=${1:-r2}
o=gpurun_out
quick="--no-cpu-baseline --no-verify --skip overlap50,sliding9,array64,receiver"
timeout 900 python bench.py > $o/${tag}_bench_1gpu.json 2> $o/${tag}_bench_1gpu.err
timeout 600 python bench.py --impl reference > $o/${tag}_bench_1gpu_reference_arm.json 2> $o/${tag}_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 $quick > $o/${tag}_ncu_list.log 2>&1
# one launch of every heavy kernel at 10 000 windows (first step: the kernels of the first slice)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_spectrogram|k_coarse|k_fine_points|k_fine_lags|k_fine_finish" -c 10 -f -o $o/${tag}_full \
    python bench.py --steps 1 --warmup 0 $quick > $o/${tag}_ncu_full.log 2>&1
./tools/fp32_pipes > $o/${tag}_fp32_pipes.txt 2>&1
ls -la $o/${tag}_*
