timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for d in 1 0; do
echo "DEFER=$d"
KZGB200_DEFER_SUBGROUP=$d timeout 300 python bench.py --steps 8 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k:round(v,2) for k,v in d['phases_ms'].items()}, round(d['pipelined']['value']), round(d['pipelined']['e2e']), round(d['exact_transcript']['value']))"
done
