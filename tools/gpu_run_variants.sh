# same-run A/B of match-kernel build variants (variants/var_*.so, built with CVG_NVCC_EXTRA / CVGRAFT_SO): ms per 64 pairs
mkdir -p gpurun_out
for r in 1 2; do
  echo "main"; CVG_LANES=1 timeout 120 python tools/gpu_prof_match.py 64 2>&1 | tail -1 | cut -c1-40
  for v in ${VARIANTS:-g4 g1 dyn saw i8}; do
    echo "$v"; CVGRAFT_SO=$PWD/variants/var_$v.so CVG_LANES=1 timeout 120 python tools/gpu_prof_match.py 64 2>&1 | tail -1 | cut -c1-40
  done
done
if [ -f variants/var_prof.so ]; then
  CVGRAFT_SO=$PWD/variants/var_prof.so CVG_TC_EXP=32 CVG_LANES=1 timeout 120 python tools/gpu_prof_match.py 64 2>&1 | tail -5 | cut -c1-400
  CVGRAFT_SO=$PWD/variants/var_prof.so CVG_TC_EXP=33 CVG_LANES=1 timeout 120 python tools/gpu_prof_match.py 64 2>&1 | tail -5 | cut -c1-400
fi
