#!/bin/bash
# A/B timing of library variants on one GPU: tools/ab.sh "<libs>" "<variants>" [extra bench args]; default lib = "-"
libs="$1"; vars="$2"; shift 2
for L in $libs; do for v in $vars; do
  if [ "$L" = "-" ]; then unset RT_B200_LIB; else export RT_B200_LIB=$PWD/build/lib_$L.so; fi
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --variant $v "$@" 2> gpurun_out/ab_${L}_v$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('lib $L variant $v: ms %.4f  kern %.4f  e2e %.4f (kernel in e2e %.4f)  Mrays/s %.0f' % (d['ms_per_step'], d['frame_kernel_ms_max_rank'], d['e2e']['ms_per_step'], d['e2e'].get('frame_kernel_ms_rank0', 0), d['value']))
" | tee -a gpurun_out/ab.log
done; done
