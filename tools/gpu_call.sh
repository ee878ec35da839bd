#!/bin/bash
o=gpurun_out
nproc > $o/r2_box8.txt; free -g >> $o/r2_box8.txt; nvidia-smi topo -m >> $o/r2_box8.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/link_study.py > $o/r2_link_study.json 2> $o/r2_link_study.err
echo "link exit $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 > $o/r2_bench_8gpu.json 2> $o/r2_bench_8gpu.err
echo "bench exit $?"; tail -2 $o/r2_bench_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > $o/r2_bench_8gpu_reference_arm.json 2> $o/r2_ref8.err
echo "ref exit $?"; cat $o/r2_bench_8gpu_reference_arm.json | head -c 400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 3 --warmup 3 > $o/r2_bench_4gpu.json 2> $o/r2_bench_4gpu.err
echo "bench4 exit $?"
