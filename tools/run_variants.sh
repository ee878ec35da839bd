#!/bin/bash
# runs bench.py (headline arms only) against every library build under build/variants/ (kernel tuning experiments)
for v in base $(ls build/variants 2>/dev/null); do
  if [ "$v" = base ]; then unset UWSPR_B200_LIB; else export UWSPR_B200_LIB=$PWD/build/variants/$v/gr-uwspr_b200/libuwspr_b200.so; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-verify --skip overlap50,sliding9,array64,receiver 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), {k: round(x,3) for k,x in d['stage_ms'].items()}, d['decoded']['correct'], 'e2e_ms', round(d['e2e']['ms_per_step'],2))"
done
