#!/bin/bash
# hash-kernel bimodality: fresh-process bench runs for both ring depths
mkdir -p gpurun_out
for rep in 1 2 3 4 5 6 7 8; do for st in 4 8; do
KZGB200_SHA_STAGES=$st timeout 300 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
o=json.loads(sys.stdin.read().strip().split('\n')[-1]); p=o['phases_ms']; print('stages=$st rep=$rep ms=%.2f hash=%.2f' % (o['ms_per_step'], p['challenge_sha256']))"
done; done
