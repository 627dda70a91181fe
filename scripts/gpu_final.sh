mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/final_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final_smoke.txt
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -n 1 gpurun_out/final_bench.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/final_ref.json 2>/dev/null
python -c "
import json
d=json.loads([l for l in open('gpurun_out/final_bench.json') if l.startswith('{')][-1]); r=json.loads([l for l in open('gpurun_out/final_ref.json') if l.startswith('{')][-1])
print('value',round(d['value'],1),'closest',round(d['config']['closest_hit_mrays_s'],1),'e2e',round(d['e2e']['value'],1),'ref',round(r['value'],2),'e2e ratio',round(d['e2e']['value']/r['value'],1), 'launches', d['gpu_launches'], 'limiter inst/ray', d['roofline']['limiter']['inst_per_ray'])"
