# A/B runs of scripts/gpu_diag.py under different B200_* environment knobs
mkdir -p gpurun_out
for v in "$@"; do
  echo "=== $v"
  env $v python scripts/gpu_diag.py 65536 2>&1 | grep -E "anyhit|closest|primary|C1|soup"
done | tee gpurun_out/ab.txt
