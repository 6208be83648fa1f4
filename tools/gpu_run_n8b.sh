set -x
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02f_bench_n$N.json 2> gpurun_out/r02f_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
j=json.load(open('gpurun_out/r02f_bench_n$N.json'))
print('value',round(j['value']), j['ms_per_step'], 'sync', round(j['value_single_context']['value']), 'serial', round(j['value_serial']['value']), 'e2e', round(j['e2e']['value']), 'u8', j['e2e_u8'] and round(j['e2e_u8']['value']), 'launches', j['gpu_launches'])
print('c4', j.get('c4')); print('c5', j.get('c5_match'))
PY
if [ -f data_cache/features_full.bin ]; then
  ./host/cvg_replay data_cache/features_full.bin gpurun_out/replay_out8 --gpus $N > gpurun_out/r02f_replay_n$N.log 2>&1
  diff -r gpurun_out/replay_out8 tests/golden/replay_output >> gpurun_out/r02f_replay_n$N.log 2>&1 && echo "replay equals golden" >> gpurun_out/r02f_replay_n$N.log
  rm -rf gpurun_out/replay_out8; cat gpurun_out/r02f_replay_n$N.log
fi
