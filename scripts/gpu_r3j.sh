mkdir -p gpurun_out
run() { python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-frames 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); x=d['config']['double_exact']; print('C3 value', round(d['value'],1), 'closest', round(d['config']['closest_hit_mrays_s'],1), 'e2e', round(d['e2e']['value'],1), 'hyb occ', round(x['hybrid_mrays_s'],1), 'hyb closest', round(x['closest_hit_hybrid_mrays_s'],1))"; }
for v in default o32x32 h64x16 h128x8 hc64x10 c64x14; do
  echo "== variant $v"
  if [ $v = default ]; then unset B200_LIB; else export B200_LIB=$PWD/lucille_b200/variants/lib_$v.so; fi
  run
done 2>&1 | tee gpurun_out/r3j_cta.txt
