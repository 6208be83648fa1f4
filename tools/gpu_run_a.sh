set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
for d in 2 4; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras --depth $d > gpurun_out/r02a_bench_d$d.json 2>> gpurun_out/r02a_bench.err; done
CVG_LANES=4 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras --depth 4 > gpurun_out/r02a_bench_l4d4.json 2>> gpurun_out/r02a_bench.err
if [ -f data_cache/features_full.bin ]; then
  ./host/cvg_replay data_cache/features_full.bin gpurun_out/replay_out > gpurun_out/r02a_replay.log 2>&1
  ./host/cvg_replay data_cache/features_full.bin gpurun_out/replay_out_sync --sync >> gpurun_out/r02a_replay.log 2>&1
  diff -r gpurun_out/replay_out gpurun_out/replay_out_sync >> gpurun_out/r02a_replay.log 2>&1 && echo "replay outputs equal" >> gpurun_out/r02a_replay.log
  diff -r gpurun_out/replay_out tests/golden/replay_output >> gpurun_out/r02a_replay.log 2>&1 && echo "replay equals golden" >> gpurun_out/r02a_replay.log
  rm -rf gpurun_out/replay_out gpurun_out/replay_out_sync
  cat gpurun_out/r02a_replay.log
fi
