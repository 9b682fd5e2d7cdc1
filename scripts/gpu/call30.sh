# 4 GPUs: cuts placed for the stretch of the orbit that is about to be rendered (value window, then the end-to-end window)
set -x
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR4 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_c30_bench_config4_n4.json 2> gpurun_out/r02_c30_bench4.err; echo "bench4 rc=$?"; tail -3 gpurun_out/r02_c30_bench4.err | cut -c1-400
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c30_bench_config4_n4.json').read().strip().splitlines()[-1])
print('n4 ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'miss', d['halo_miss'], d['strips']['strip_bounds'], d['strips'].get('strip_bounds_e2e'), d.get('invalid'))
print('   per rank', d['stage_ms_per_rank'])
for r in d.get('strip_refinement') or []: print('   ', r.get('window'), r['bounds'], r['kernel_ms_per_rank'])
PY
