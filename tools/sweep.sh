#!/bin/bash
# Tooling: quick option sweeps of bench.py (prints ms/step and phase split). usage: tools/sweep.sh "<bench args>" opt1 opt2 ...
ARGS="$1"; shift
for o in "$@"; do
  OPTS=""; for kv in $(echo $o | tr ',' ' '); do OPTS="$OPTS --opt $kv"; done
  python bench.py $ARGS --no-cpu-baseline --no-e2e $OPTS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); x=d['detail']
print('$o', 'ms/step %.2f' % d['ms_per_step'], 'frac %.3f' % x['step_fp64_frac'], {k: round(v,2) for k,v in x['phase_ms'].items()}, 'syrk TF %.1f var TF %.1f' % (d['roofline']['achieved'], x['kernels']['variance_gemm']['achieved_tflops']), 'kuf ms %.1f' % x['kernels']['kuf']['ms_total'])
"
done
