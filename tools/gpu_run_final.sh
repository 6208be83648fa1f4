#!/bin/bash
# Closing run of a round (one gpurun call): GPU tests, smoke, the bench line, its launch list, one full ncu capture of the match kernel.
# usage: gpurun -- bash tools/gpu_run_final.sh r02g
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; cut -c1-600 $out/${tag}_bench_n1.json
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > $out/${tag}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > $out/${tag}_ncu_launches.log 2>&1
CVG_LANES=1 timeout 120 python tools/gpu_prof_match.py 64 > $out/${tag}_plain_prof.log 2>&1 &&
CVG_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_tc -s 1 -c 1 -o $out/${tag}_prof_match \
    python tools/gpu_prof_match.py 64 > $out/${tag}_ncu_match.log 2>&1
tail -n 1 $out/${tag}_plain_prof.log | cut -c1-60
tail -n 2 $out/${tag}_ncu_launches.log $out/${tag}_ncu_match.log
