mkdir -p gpurun_out
for v in default cr1 cr2 cr8 cr12; do
  echo "== variant $v"
  if [ $v = default ]; then unset B200_LIB; else export B200_LIB=$PWD/lucille_b200/variants/lib_$v.so; fi
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-frames 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C3 value', d['value'], 'closest', d['config']['closest_hit_mrays_s'])"
done 2>&1 | tee gpurun_out/r3d_refill.txt
