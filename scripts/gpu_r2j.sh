mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-frames 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'closest', d['config']['closest_hit_mrays_s'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'ray_batch', d['e2e']['ray_batch']['value'])"
