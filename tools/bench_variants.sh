#!/bin/bash
# run bench.py once per library variant; prints value / fwd / inv per variant
for so in "$@"; do
  VKHEL_LIB_PATH=$so python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python3 -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$so'.split('/')[-1], 'NTT/s %.0f'%d['value'], 'step %.4f ms'%d['ms_per_step'], 'fwd %.4f inv %.4f'%(d['roofline']['forward_ms'], d['roofline']['inverse_ms']), 'ok' if d['config']['round_trip_exact'] else 'MISMATCH')
"
done
