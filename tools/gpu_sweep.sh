cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python bench.py --steps 20 --warmup 3 --no-fits --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],2), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms'].items()})"
timeout 500 python tools/batch_phases.py 2>&1 | grep "tile= 64" | grep "B=32\|B=64"
