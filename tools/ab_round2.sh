#!/bin/bash
# A/B of library variants on the GPU box (run under gpurun): the whole GPU
# suite on the default build first, then bench.py per variant (twice,
# alternating; 200 steps each, results checked by bench.py itself).
# Usage: bash tools/ab_round2.sh <tag> <variant.so>...
tag=${1:-ab}; shift
out=gpurun_out; mkdir -p $out
( time python -m pytest tests -x -q -m gpu ) > $out/pytest_$tag.log 2>&1
tail -5 $out/pytest_$tag.log
for rep in 1 2; do
  for so in vkhel_b200/lib/libvkhel.so "$@"; do
    VKHEL_LIB_PATH=$so python bench.py --steps 200 --warmup 5 --no-cpu 2>/dev/null | python3 -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$so'.split('/')[-1], 'NTT/s %.0f'%d['value'], 'step %.4f ms'%d['ms_per_step'], 'fwd %.4f inv %.4f'%(d['roofline']['forward_ms'], d['roofline']['inverse_ms']), 'ok' if d['config']['round_trip_exact'] else 'MISMATCH', d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))
"
  done
done | tee $out/variants_$tag.txt
