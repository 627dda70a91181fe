mkdir -p gpurun_out
for v in default c128x7 c128x6 c192x4 c128x8; do
  echo "== variant $v"
  if [ $v = default ]; then unset B200_LIB; else export B200_LIB=$PWD/lucille_b200/variants/lib_$v.so; fi
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-frames 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C3 value', d['value'], 'closest', d['config']['closest_hit_mrays_s'], 'e2e', d['e2e']['value'])"
  python scripts/c5_rate.py 2>&1 | grep -E "C5 soup closest|primary"
done 2>&1 | tee gpurun_out/r2u_closest.txt
unset B200_LIB
python scripts/transport_rates.py 2>&1 | grep -v "lucille\]" | tee gpurun_out/r2u_transports.txt
