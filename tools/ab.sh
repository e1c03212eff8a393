#!/bin/bash
# same-box A/B of library builds / bench flags: tools/ab.sh "<lib.so|-> [bench flags]" ...  (each spec runs twice, alternating)
set -u
mkdir -p gpurun_out
specs=("$@")
i=0
for rep in 1 2; do
  for spec in "${specs[@]}"; do
    read -r lib flags <<< "$spec"
    flagsenv=""; case "$flags" in *=*) flagsenv="$flags"; flags="";; esac   # "lib VAR=value" sets an environment variable instead
    i=$((i+1))
    echo "== $lib $flags (rep $rep)"
    if [ "$lib" = "-" ]; then unset KB_LIB; else export KB_LIB=$lib; fi
    env $flagsenv KB_TRACE_STALLS=5 python bench.py --no-nn --no-cpu --steps 200 --warmup 10 $flags 2>gpurun_out/ab_$i.err | grep '^{' > gpurun_out/ab_$i.json
    python - <<PY
import json
d = json.loads(open('gpurun_out/ab_$i.json').read()); c = d['config']
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'f32', round(d['e2e_f32']['value'],1), 'blocking', {k: round(v,1) for k, v in d['blocking_calls'].items() if isinstance(v, float)})
print('   clk', d['clocks'], 'lat', c['call_latency_ms'])
PY
    grep "kb stall" gpurun_out/ab_$i.err | head -5
  done
done
