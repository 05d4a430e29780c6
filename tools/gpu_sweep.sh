cd "${GRAFT_REPO_ROOT:-.}"
run() {
env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-fits --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['value'],2), round(d['ms_per_step'],3), round(d['phases_ms']['potrf'],3))"
}
run CS_PRIO=2,0,1,5,5
run CS_PRIO=2,0,1,5,5 CS_PANEL=6
run CS_PRIO=2,0,1,5,5 CS_PANEL=5
run CS_PRIO=1,0,1,5,5
run CS_PRIO=3,0,1,5,5
run CS_PRIO=2,0,1,5,4
run CS_PRIO=2,0,0,5,5
run CS_PRIO=2,0,2,5,5
run CS_PRIO=2,0,1,5,5
