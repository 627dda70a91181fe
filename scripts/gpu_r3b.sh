mkdir -p gpurun_out
python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r3b_tests.txt
python bench.py --no-cpu-baseline > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err; tail -n 2 gpurun_out/r3b_bench.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r3b_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'closest',d['config']['closest_hit_mrays_s'],'e2e',d['e2e']['value']); print(d['config']['double_exact'])"
