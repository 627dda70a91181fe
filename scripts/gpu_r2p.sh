# round 2 profile pass: launch list of the bench + one `ncu --set full` capture per kernel (B200_STREAMED=0: ncu serialises the copy stream).
# scripts/prof_r02.py launches every kernel twice, then once more inside cudaProfilerStart/Stop: --profile-from-start off picks that launch.
mkdir -p gpurun_out
export B200_STREAMED=0
NCU="ncu --set full --clock-control none --import-source on -f"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-frames > gpurun_out/r02_ncu_bench.log 2>&1
$NCU -k regex:occluded_pool32 -s 2 -c 1 -o gpurun_out/r02_occ_f32_c3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-frames > gpurun_out/r02_ncu1.log 2>&1
P="--profile-from-start off"
$NCU $P -k regex:closest_pool32 -c 1 -o gpurun_out/r02_closest_f32_c3 python scripts/prof_r02.py c3 > gpurun_out/r02_ncu2.log 2>&1
$NCU $P -k regex:occluded_hybrid -c 1 -o gpurun_out/r02_occ_hybrid_c3 python scripts/prof_r02.py c3 > gpurun_out/r02_ncu8.log 2>&1
$NCU $P -k regex:closest_hybrid -c 1 -o gpurun_out/r02_closest_hybrid_c3 python scripts/prof_r02.py c3 > gpurun_out/r02_ncu9.log 2>&1
$NCU $P -k regex:occluded_pool_kernel -c 1 -o gpurun_out/r02_occ_f64_c3 python scripts/prof_r02.py c3 > gpurun_out/r02_ncu3.log 2>&1
$NCU $P -k regex:closest_pool_kernel -c 1 -o gpurun_out/r02_closest_f64_c3 python scripts/prof_r02.py c3 > gpurun_out/r02_ncu4.log 2>&1
$NCU $P -k regex:occluded_pool32 -c 1 -o gpurun_out/r02_occ_f32_c5 python scripts/prof_r02.py c5 > gpurun_out/r02_ncu5.log 2>&1
B200_POOL_TOPSMEM=1 $NCU $P -k regex:occluded_pool_kernel -c 1 -o gpurun_out/r02_occ_f32_c3_topsmem python scripts/prof_r02.py c3 > gpurun_out/r02_ncu6.log 2>&1
B200_POOL32=0 $NCU $P -k regex:occluded_pool_kernel -c 1 -o gpurun_out/r02_occ_f32_c3_generic python scripts/prof_r02.py c3 > gpurun_out/r02_ncu7.log 2>&1
tail -2 gpurun_out/r02_ncu*.log
ls -la gpurun_out/*.ncu-rep
# summaries on the box (the reports together exceed what gpurun brings back); only the two newest kernels' reports travel
S=gpurun_out/summaries; mkdir -p $S
T4="4194304"
python scripts/ncu_md.py gpurun_out/r02_occ_f32_c3.ncu-rep $S/r02_occluded_f32_c3 "r02: fp32 occlusion (pool32.cuh), the timed kernel of bench.py: 1 M triangles, the 16 Mi-ray C3 batch" 16777216
python scripts/ncu_md.py gpurun_out/r02_closest_f32_c3.ncu-rep $S/r02_closest_f32_c3 "r02: fp32 closest hit (pool32.cuh), 1 M triangles, first 4 Mi AO rays of the C3 batch" $T4
python scripts/ncu_md.py gpurun_out/r02_occ_hybrid_c3.ncu-rep $S/r02_occluded_hybrid_c3 "r02: double-exact occlusion through the fp32 records (hybrid.cuh), 1 M triangles, first 4 Mi AO rays of the C3 batch as doubles" $T4
python scripts/ncu_md.py gpurun_out/r02_closest_hybrid_c3.ncu-rep $S/r02_closest_hybrid_c3 "r02: double-exact closest hit through the fp32 records (hybrid.cuh), 1 M triangles, first 4 Mi AO rays of the C3 batch as doubles" $T4
python scripts/ncu_md.py gpurun_out/r02_occ_f64_c3.ncu-rep $S/r02_occluded_f64_c3 "r02: double occlusion (pool.cuh, double records), 1 M triangles, first 4 Mi AO rays of the C3 batch" $T4
python scripts/ncu_md.py gpurun_out/r02_closest_f64_c3.ncu-rep $S/r02_closest_f64_c3 "r02: double closest hit (pool_closest.cuh), 1 M triangles, first 4 Mi AO rays of the C3 batch" $T4
python scripts/ncu_md.py gpurun_out/r02_occ_f32_c5.ncu-rep $S/r02_occluded_f32_c5 "r02: fp32 occlusion (pool32.cuh) on the 10 M-triangle soup of configs[4] -- 1.2 GB of records, beyond L2 -- 4 Mi AO rays" $T4
python scripts/ncu_md.py gpurun_out/r02_occ_f32_c3_topsmem.ncu-rep $S/r02_occluded_f32_c3_topsmem "r02: X1 experiment -- generic pooled fp32 occlusion with the top 256 nodes staged in shared memory by cp.async.bulk (B200_POOL_TOPSMEM=1), first 4 Mi rays of C3" $T4
python scripts/ncu_md.py gpurun_out/r02_occ_f32_c3_generic.ncu-rep $S/r02_occluded_f32_c3_generic "r02: generic pooled fp32 occlusion (pool.cuh, B200_POOL32=0), first 4 Mi rays of C3 -- the A/B arm of the two profiles beside it" $T4
rm -f gpurun_out/r02_closest_f32_c3.ncu-rep gpurun_out/r02_occ_f32_c3.ncu-rep gpurun_out/r02_occ_f64_c3.ncu-rep gpurun_out/r02_closest_f64_c3.ncu-rep gpurun_out/r02_occ_f32_c5.ncu-rep gpurun_out/r02_occ_f32_c3_topsmem.ncu-rep gpurun_out/r02_occ_f32_c3_generic.ncu-rep
du -sh gpurun_out
