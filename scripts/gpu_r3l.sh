mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r3l_tests.txt
run() { python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-frames 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); x=d['config']['double_exact']; print('C3 value', round(d['value'],1), 'closest', round(d['config']['closest_hit_mrays_s'],1), 'e2e', round(d['e2e']['value'],1), 'hyb occ', round(x['hybrid_mrays_s'],1), 'e2e64', round(x['e2e_points_f64_mrays_s'],1))"; }
for o in auto 0 1 2; do
  echo "== order $o"
  if [ $o = auto ]; then unset B200_ANYHIT_ORDER; else export B200_ANYHIT_ORDER=$o; fi
  run
  python scripts/c5_rate.py 2>&1 | grep -E "C5 soup anyhit"
done 2>&1 | tee gpurun_out/r3l_order.txt
unset B200_ANYHIT_ORDER
python scripts/transport_rates.py 2>&1 | grep -E "^AO|sun-sky" | tee -a gpurun_out/r3l_order.txt
