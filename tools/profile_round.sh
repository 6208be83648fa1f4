#!/bin/bash
# Round evidence run (one gpurun call): bench lines, launch list of the same command, ncu --set full of the top kernels.
# usage: gpurun -- bash tools/profile_round.sh r02
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $out/${tag}_gpu_info.csv
lscpu | head -25 > $out/${tag}_cpu_info.txt
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_n1.json 2> $out/${tag}_bench_reference_n1.err
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > $out/${tag}_ncu_launches.log 2>&1
CVG_LANES=1 python tools/gpu_prof_match.py 64 > $out/${tag}_plain_prof.log 2>&1 &&
CVG_LANES=1 ncu --set full --clock-control none --import-source on -k regex:match_tc -s 1 -c 1 -o $out/${tag}_prof_match \
    python tools/gpu_prof_match.py 64 > $out/${tag}_ncu_match.log 2>&1
CVG_LANES=1 ncu --set full --clock-control none --import-source on -k regex:ransac_hyp_t -s 4 -c 1 -o $out/${tag}_prof_hyp \
    python tools/gpu_prof_match.py 64 > $out/${tag}_ncu_hyp.log 2>&1
CVG_LANES=1 ncu --set full --clock-control none --import-source on -k regex:ransac_finish -s 1 -c 1 -o $out/${tag}_prof_finish \
    python tools/gpu_prof_match.py 64 > $out/${tag}_ncu_finish.log 2>&1
CVG_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:match_tc --csv --log-file $out/${tag}_single_pair_match.csv \
    python tools/gpu_prof_match.py 1 > $out/${tag}_ncu_single.log 2>&1
cat $out/${tag}_bench_n1.json | cut -c1-400; cat $out/${tag}_bench_reference_n1.json | cut -c1-300
tail -n 2 $out/${tag}_ncu_launches.log $out/${tag}_ncu_match.log $out/${tag}_ncu_hyp.log $out/${tag}_ncu_finish.log
