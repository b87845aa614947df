#!/bin/bash
# A/B of library variants on ONE box: profiles/ab_run.sh <rounds> <variant.so[:ENV=val]>...
L=multimodal-fusion-based-pre-routing-timing-prediction-_b200/libtm_b200.so
cp $L /tmp/orig.so
R=$1; shift
for r in $(seq $R); do
  for v in "$@"; do
    so=${v%%:*}; envs=""
    if [[ "$v" == *:* ]]; then envs=${v#*:}; fi
    cp ab/$so $L
    out=$(env $envs python bench.py --no-configs --no-cpu-baseline --config5 0 --no-shuffled --no-sustained 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['value'])")
    echo "$r $v $out"
  done
done
cp /tmp/orig.so $L
