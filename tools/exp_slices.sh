for s in 1 4 8; do echo "== slices=$s"; python bench.py --slices $s --steps 2 --warmup 2 --no-e2e --no-cpu 2>gpurun_out/slices_$s.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().splitlines()[-1]); print('value', d['value'], 'ms', d['ms_per_step'])"; tail -2 gpurun_out/slices_$s.err; done
