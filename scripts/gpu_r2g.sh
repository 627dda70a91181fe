mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in default c3; do
  echo "== $v"
  if [ $v = default ]; then unset B200_LIB; else export B200_LIB=$PWD/lucille_b200/variants/lib_$v.so; fi
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'closest', d['config']['closest_hit_mrays_s'], 'e2e', d['e2e']['value'])"
done | tee gpurun_out/r2g_closest.txt
