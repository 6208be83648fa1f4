mkdir -p gpurun_out
for cfg in "6 6" "8 8" "6 4"; do set -- $cfg
CVG_LANES=$1 timeout 300 python bench.py --steps 24 --warmup 4 --no-cpu --depth $2 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print('lanes $1 depth $2: value',round(j['value']), 'sync', round(j['value_single_context']['value']), 'e2e', round(j['e2e']['value']), 'u8', round(j['e2e_u8']['value']), 'real', j['real_dataset']['seconds'], j['real_dataset']['seconds_sync_calls'])"
done
./host/cvg_replay data_cache/features_full.bin gpurun_out/rp1 | tail -2; ./host/cvg_replay data_cache/features_full.bin gpurun_out/rp2 --depth 3 | tail -1; rm -rf gpurun_out/rp1 gpurun_out/rp2
