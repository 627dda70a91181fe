# round 2, session 2: full GPU suite (new: shade callers), then the L2-prefetch experiment (X1) on the 10 M-triangle scene and on C3
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2q_tests.txt
for v in default pf1 pf2 pf1far far; do
  echo "== variant $v"
  if [ $v = default ]; then unset B200_LIB; else export B200_LIB=$PWD/lucille_b200/variants/lib_$v.so; fi
  python scripts/c5_rate.py 2>&1 | grep -E "C5 soup"
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-frames 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C3 value', d['value'], 'closest', d['config']['closest_hit_mrays_s'], 'e2e', d['e2e']['value'])"
done 2>&1 | tee gpurun_out/r2q_prefetch.txt
