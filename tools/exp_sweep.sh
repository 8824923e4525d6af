for d in 0 1 2; do echo "== WB_SWEEP_DEBUG=$d"; WB_SWEEP_DEBUG=$d python bench.py --utts 128 --steps 2 --warmup 1 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('sweep ms', d['kernels']['band_sweep_kernel']['ms_per_step'])"; done
