mkdir -p gpurun_out
for o in 1 2 3; do
  echo "== order $o"; export B200_ANYHIT_ORDER=$o
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-frames 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C3 value', round(d['value'],1))"
  python scripts/c5_rate.py 2>&1 | grep -E "C5 soup anyhit"
done 2>&1 | tee gpurun_out/r3m_order.txt
