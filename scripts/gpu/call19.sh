# 2 GPUs: the split spatial pass across processes -- bit-exactness vs the full frame, then the bench
set -x
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR2 scripts/verify_multigpu.py config4 3 > gpurun_out/r02_c19_verify_config4_n2.json 2> gpurun_out/r02_c19_verify.err; echo "verify rc=$?"; tail -c 600 gpurun_out/r02_c19_verify_config4_n2.json; tail -3 gpurun_out/r02_c19_verify.err | cut -c1-300
timeout 600 $TR2 scripts/verify_multigpu.py config3 4 > gpurun_out/r02_c19_verify_config3_n2.json 2>> gpurun_out/r02_c19_verify.err; echo "verify rc=$?"; tail -c 600 gpurun_out/r02_c19_verify_config3_n2.json
timeout 600 $TR2 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_c19_bench_config4_n2.json 2> gpurun_out/r02_c19_bench2.err; echo "bench2 rc=$?"; tail -2 gpurun_out/r02_c19_bench2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c19_bench_config4_n2.json').read().strip().splitlines()[-1])
print('n2 ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'miss', d['halo_miss'], d['strips']['strip_bounds'], d['stage_ms_per_rank'])
print(d.get('strip_refinement'))
PY
