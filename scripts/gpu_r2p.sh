# round 2 profile pass: launch list of the bench + one `ncu --set full` capture per kernel (B200_STREAMED=0: ncu serialises the copy stream)
mkdir -p gpurun_out
export B200_STREAMED=0
NCU="ncu --set full --clock-control none --import-source on -f"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-frames > gpurun_out/r02_ncu_bench.log 2>&1
$NCU -k regex:occluded_pool32 -s 2 -c 1 -o gpurun_out/r02_occ_f32_c3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-frames > gpurun_out/r02_ncu1.log 2>&1
$NCU -k regex:closest_pool32 -s 2 -c 1 -o gpurun_out/r02_closest_f32_c3 python scripts/prof_r02.py c3 > gpurun_out/r02_ncu2.log 2>&1
$NCU -k regex:occluded_pool_kernel -s 2 -c 1 -o gpurun_out/r02_occ_f64_c3 python scripts/prof_r02.py c3 > gpurun_out/r02_ncu3.log 2>&1
$NCU -k regex:closest_pool_kernel -s 2 -c 1 -o gpurun_out/r02_closest_f64_c3 python scripts/prof_r02.py c3 > gpurun_out/r02_ncu4.log 2>&1
$NCU -k regex:occluded_pool32 -s 2 -c 1 -o gpurun_out/r02_occ_f32_c5 python scripts/prof_r02.py c5 > gpurun_out/r02_ncu5.log 2>&1
B200_POOL_TOPSMEM=1 $NCU -k regex:occluded_pool_kernel -s 2 -c 1 -o gpurun_out/r02_occ_f32_c3_topsmem python scripts/prof_r02.py c3 > gpurun_out/r02_ncu6.log 2>&1
B200_POOL32=0 $NCU -k regex:occluded_pool_kernel -s 2 -c 1 -o gpurun_out/r02_occ_f32_c3_generic python scripts/prof_r02.py c3 > gpurun_out/r02_ncu7.log 2>&1
tail -2 gpurun_out/r02_ncu*.log
ls -la gpurun_out/*.ncu-rep
