#!/bin/bash
# runs bench.py (index build only) against the default library and every gpurun_variants/*.so
for lib in default gpurun_variants/*.so; do
  if [ "$lib" = default ]; then unset BWTK_LIB; else export BWTK_LIB=$PWD/$lib; fi
  python bench.py --no-cpu --skip-extras --steps 10 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ks={k['kernel']:k['ms'] for k in d.get('kernels',[])}
print('$lib', d['ms_per_step'], d['e2e']['ms_per_step'], sorted(ks.items(), key=lambda kv:-kv[1])[:6])
"
done
