mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r3w_tests.txt
python bench.py --no-cpu-baseline > gpurun_out/r3w_bench1.json 2> gpurun_out/r3w_bench1.err; echo "N=1 rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r3w_bench1.json') if l.startswith('{')][-1]); f=d['config']['frames']; c5=f['configs[4]']; c1=f['configs[0]']
print('value', round(d['value'],1), 'C5 peer', round(c5['fused_peer_store']['wall_ms_to_rank0_host_framebuffer'],1), [round(x) for x in c5['fused_peer_store']['device_ms_per_rank']], 'nccl', round(c5['nccl_gather']['wall_ms_to_rank0_host_framebuffer'],1), c5['equals_one_rank_frame'], 'C1', round(c1['wall_ms_to_rank0_host_framebuffer'],2), c1['equals_reference_framebuffer'])
"
