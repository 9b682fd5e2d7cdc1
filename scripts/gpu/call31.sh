# 8 GPUs: the driver's bench command with cuts placed for the stretch of the orbit about to be rendered
set -x
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 400 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_c31_bench_config4_n8.json 2> gpurun_out/r02_c31_bench8.err; echo "bench8 rc=$?"; tail -2 gpurun_out/r02_c31_bench8.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c31_bench_config4_n8.json').read().strip().splitlines()[-1])
print('n8 ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'miss', d['halo_miss'], d['strips']['strip_bounds'], d['strips'].get('strip_bounds_e2e'), d.get('invalid'))
print('   per rank', d['stage_ms_per_rank'])
for r in d.get('strip_refinement') or []: print('   ', r.get('window'), r['bounds'][1], max(r['kernel_ms_per_rank']), min(r['kernel_ms_per_rank']))
PY
