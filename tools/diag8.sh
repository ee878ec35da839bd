#!/bin/bash
# 8-GPU host-fed diagnosis: pipeline traces under PCIe contention
run() { tag=$1; shift; env "$@" UWSPR_B200_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/diag8_$tag.json 2> gpurun_out/diag8_$tag.err; python -c "
import json; d=json.load(open('gpurun_out/diag8_$tag.json')); print('$tag', round(d['value']), d['e2e']['ms_per_step'], d['e2e']['pcie_h2d_gbs'])"; grep "nwin 10000 chunks" gpurun_out/diag8_$tag.err | tail -3; }
run default A=1
run plain UWSPR_B200_NO_EARLY_D2H=1 UWSPR_B200_NO_TAIL_SPLIT=1
run chunk2500 UWSPR_B200_HOST_CHUNK=2500
