#!/bin/bash
# tuning sweep of the candidate threshold c and the strip length S (bench.py reads NTL_CAND_C / NTL_STRIP_LEN)
for S in ${SWEEP_S:-128 256 512}; do for C in ${SWEEP_C:-5 6 7 8}; do
  NTL_CAND_C=$C NTL_STRIP_LEN=$S timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('S=$S c=$C value=%.1f e2e=%.1f ms=%.3f dense=%.3f select=%.3f gap=%.3f emit=%.3f pack=%.3f map=%.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], s['dense'], s['select'], s['gap'], s['emit'], s['pack'], s['lookup']+s['chain']+s['tally']+s['index']))"
done; done
