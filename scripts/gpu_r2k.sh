mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "drop_in or batched_frame or point_ao" 2>&1 | tail -4
python scripts/lsh_frame_times.py 2>&1 | tee gpurun_out/r02_lsh_frame.txt
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-frames 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'ray_batch', d['e2e']['ray_batch']['value'])"
