mkdir -p gpurun_out
for e in "B200_POOL=0" "B200_POOL32=0" "B200_HYBRID=0" "B200_STREAMED=0" "B200_POOL_CLOSEST=0" "B200_POOL_CLOSEST64=0" "B200_ANYHIT_ORDER=0" "B200_ANYHIT_ORDER=2" "B200_ANYHIT_ORDER=3" "B200_HYBRID_OWN=1" "B200_POOL_TOPSMEM=1" "B200_BUILD=host" "B200_BUILD=device"; do
  echo "== $e: $(env $e python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k 'closest_hit or occlusion_matches or edge_case or axis_parallel or hybrid_occlusion or c1_frame_against or plane_sphere or whitted or dirtmap or point_gathers' 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/r3p_knobs.txt
