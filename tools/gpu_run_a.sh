set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/r02e_bench.json'))
print('value',round(j['value']), j['ms_per_step'], 'sync', round(j['value_single_context']['value']), 'serial', round(j['value_serial']['value']), 'e2e', round(j['e2e']['value']), 'u8', j['e2e_u8'] and round(j['e2e_u8']['value']), 'real', j['real_dataset'] and (j['real_dataset'].get('seconds'), j['real_dataset'].get('seconds_sync_calls')), 'launches', j['gpu_launches'], 'match', j['roofline']['kernel_ms_per_launch'])
PY
