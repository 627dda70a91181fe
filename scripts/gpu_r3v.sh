mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r3v_bench2.json 2> gpurun_out/r3v_bench2.err
echo "N=2 rc=$?"; tail -n 2 gpurun_out/r3v_bench2.err
python bench.py --no-cpu-baseline > gpurun_out/r3v_bench1.json 2> gpurun_out/r3v_bench1.err; echo "N=1 rc=$?"
python -c "
import json
for n in (1,2):
    d=json.loads([l for l in open(f'gpurun_out/r3v_bench{n}.json') if l.startswith('{')][-1]); f=d['config']['frames']; c5=f['configs[4]']; c1=f['configs[0]']
    print(n, 'C5 peer', round(c5['fused_peer_store']['wall_ms_to_rank0_host_framebuffer'],1), [round(x) for x in c5['fused_peer_store']['device_ms_per_rank']], 'nccl', round(c5['nccl_gather']['wall_ms_to_rank0_host_framebuffer'],1), c5['equals_one_rank_frame'], 'C1', round(c1['wall_ms_to_rank0_host_framebuffer'],2), c1['equals_reference_framebuffer'])
"
