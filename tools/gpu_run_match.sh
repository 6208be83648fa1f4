mkdir -p gpurun_out
for r in 1 2; do
for v in sl20 sl100; do echo $v; CVGRAFT_SO=$PWD/variants/var_$v.so CVG_LANES=1 python tools/gpu_prof_match.py 64 2>&1 | tail -1 | cut -c1-40; CVGRAFT_SO=$PWD/variants/var_$v.so CVG_TC_EXP=1 CVG_LANES=1 python tools/gpu_prof_match.py 64 2>&1 | tail -1 | cut -c1-40; done
echo base; CVG_LANES=1 python tools/gpu_prof_match.py 64 2>&1 | tail -1 | cut -c1-40; CVG_TC_EXP=1 CVG_LANES=1 python tools/gpu_prof_match.py 64 2>&1 | tail -1 | cut -c1-40
done
