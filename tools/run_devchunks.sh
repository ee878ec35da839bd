#!/bin/bash
# device-resident step time as a function of how many chunks alternate between the two streams
for n in 1 2 4 8 16; do
  UWSPR_B200_DEV_CHUNKS=$n python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dev_chunks $n', round(d['value']), round(d['ms_per_step'],2), {k: round(x,2) for k,x in d['stage_ms'].items()}, d['decoded']['correct'])"
done
