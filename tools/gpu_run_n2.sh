set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_lanes_multi.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02b_pytest_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_n2.log
tail -5 gpurun_out/r02b_pytest_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r02b_bench_n2.json
tail -5 gpurun_out/r02b_bench_n2.err
if [ -f data_cache/features_full.bin ]; then
  ./host/cvg_replay data_cache/features_full.bin gpurun_out/replay_out2 --gpus 2 > gpurun_out/r02b_replay_n2.log 2>&1
  diff -r gpurun_out/replay_out2 tests/golden/replay_output >> gpurun_out/r02b_replay_n2.log 2>&1 && echo "replay equals golden" >> gpurun_out/r02b_replay_n2.log
  rm -rf gpurun_out/replay_out2
  cat gpurun_out/r02b_replay_n2.log
fi
