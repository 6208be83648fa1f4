#!/bin/bash
# Closing check when the GPU budget is short: GPU tests, smoke, the full bench line.  usage: gpurun -- bash tools/gpu_run_final_short.sh r02h
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; cut -c1-300 $out/${tag}_bench_n1.json
