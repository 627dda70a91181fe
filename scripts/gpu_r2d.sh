# round 2, call D: pool32 (static-smem, diet) vs generic pooled kernel, same library
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in 1 0 1 0; do
  echo "== B200_POOL32=$v"
  B200_POOL32=$v ORDERS=batch,random python scripts/exp_sort.py 2>&1 | grep order
done | tee gpurun_out/r2d_ab.txt
