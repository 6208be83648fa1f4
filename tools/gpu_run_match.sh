mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lanes_multi.py -m gpu -x -q -k "match or guard" 2>&1 | tail -2
for r in 1 2 3; do
for v in ins4; do echo $v; CVGRAFT_SO=$PWD/variants/var_$v.so CVG_LANES=1 python tools/gpu_prof_match.py 64 2>&1 | tail -1 | cut -c1-40; done
echo ins8; CVG_LANES=1 python tools/gpu_prof_match.py 64 2>&1 | tail -1 | cut -c1-40
done
