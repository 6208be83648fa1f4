mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lanes_multi.py -m gpu -x -q -k "match or guard" 2>&1 | tail -2
for r in 1 2; do
echo base; CVGRAFT_SO=$PWD/variants/var_base.so CVG_LANES=1 timeout 120 python tools/gpu_prof_match.py 64 2>&1 | tail -1 | cut -c1-40
echo new; CVG_LANES=1 timeout 120 python tools/gpu_prof_match.py 64 2>&1 | tail -1 | cut -c1-40
done
echo "mma only base"; CVGRAFT_SO=$PWD/variants/var_base.so CVG_TC_EXP=1 CVG_LANES=1 timeout 120 python tools/gpu_prof_match.py 64 2>&1 | tail -1 | cut -c1-40
echo "mma only new"; CVG_TC_EXP=1 CVG_LANES=1 timeout 120 python tools/gpu_prof_match.py 64 2>&1 | tail -1 | cut -c1-40
