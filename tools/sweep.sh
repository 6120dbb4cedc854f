#!/bin/bash
# usage: tools/sweep.sh "<variants>" "<leaf sizes>" [extra bench args]  -> one line per combination
for v in $1; do for lm in $2; do
python bench.py --profile --steps 5 --warmup 2 --variant $v --leaf-max $lm $3 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('variant',d['config']['variant'],'leaf',d['config']['leaf_max'],'Mrays/s %.0f'%d['value'],'ms %.3f'%d['ms_per_step'],'warm %.0f'%d['value_warm_l2'],'nodes/ray %.1f'%d['roofline']['nodes_per_ray'],'tris/ray %.1f'%d['roofline']['tris_per_ray'],'e2e %.0f'%d['e2e']['value'],'lines/ray %.2f blocks/ray %.2f'%(d['roofline']['node_lines_per_ray'],d['roofline']['tri_blocks_per_ray']),'rays %.2fM'%((d['config']['rays_primary']+d['config']['rays_shadow'])/1e6))
except Exception as e: print('FAILED', e)
"
done; done
