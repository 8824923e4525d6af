for c in 32 64 96 226; do echo "== WB_HOST_CHUNK=$c"; WB_HOST_CHUNK=$c python bench.py --steps 3 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().splitlines()[-1]); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), [round(x) for x in d['e2e']['ms_per_step_rank0']])"; done
