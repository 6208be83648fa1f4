set -x
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8; lscpu | grep -E "Model name|Socket|NUMA node|^CPU\(s\)" ; nvidia-smi topo -m | head -14
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/h2d_bw_multi.py > gpurun_out/r02_h2d_bw_n$N.json 2> gpurun_out/r02_h2d_bw_n$N.err; echo "h2d rc=$?"; cat gpurun_out/r02_h2d_bw_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"
tail -c 1800 gpurun_out/r02_bench_n$N.json; tail -3 gpurun_out/r02_bench_n$N.err
if [ -f data_cache/features_full.bin ]; then
  ./host/cvg_replay data_cache/features_full.bin gpurun_out/replay_out8 --gpus $N > gpurun_out/r02_replay_n$N.log 2>&1
  diff -r gpurun_out/replay_out8 tests/golden/replay_output >> gpurun_out/r02_replay_n$N.log 2>&1 && echo "replay equals golden" >> gpurun_out/r02_replay_n$N.log
  rm -rf gpurun_out/replay_out8; cat gpurun_out/r02_replay_n$N.log
fi
timeout 600 python -m pytest tests/test_gpu_lanes_multi.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -3
