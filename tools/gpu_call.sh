#!/bin/bash
o=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-verify --skip overlap50,receiver 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); o=d['array64']; print('headline', round(d['value']), {k: round(x,3) for k,x in d['stage_ms'].items()}, d['decoded']['correct'], 'e2e', round(d['e2e']['value'])); print('array64', round(o['value']), round(o['e2e']['value']), o['stage_ms_rank0'], o['decoded']['correct'])"
