# round-2 GPU call 27 (8 GPUs): the final library on strips -- verify, the driver's bench command (config4 / config2), the k x passes sweep on a continuous orbit
set -x
python -c "import restir_b200 as rb; print('build', rb.api.build_id())"
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $TR8 scripts/verify_multigpu.py config4 3 > gpurun_out/r02_c27_verify_config4_n8.json 2> gpurun_out/r02_c27_verify.err; echo "verify rc=$?"; tail -c 400 gpurun_out/r02_c27_verify_config4_n8.json
timeout 600 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_c27_bench_config4_n8.json 2> gpurun_out/r02_c27_bench8.err; echo "bench8 rc=$?"; tail -2 gpurun_out/r02_c27_bench8.err | cut -c1-300
B=$(python -c "import json; d=json.loads(open('gpurun_out/r02_c27_bench_config4_n8.json').read().strip().splitlines()[-1]); print(','.join(str(b) for b in d['strips']['strip_bounds']))")
timeout 600 $TR8 scripts/config5_sweep.py config4 8 $B > gpurun_out/r02_c27_config5_sweep_n8.json 2> gpurun_out/r02_c27_config5.err; echo "sweep rc=$?"; tail -c 1200 gpurun_out/r02_c27_config5_sweep_n8.json
timeout 400 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 --workload config2 > gpurun_out/r02_c27_bench_config2_n8.json 2> gpurun_out/r02_c27_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for f in ("r02_c27_bench_config4_n8","r02_c27_bench_config2_n8"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'miss', d['halo_miss'], d['strips']['strip_bounds'], d.get('invalid'))
        print('   per rank', d['stage_ms_per_rank'])
        print('   refine', [(r['bounds'][1], max(r['kernel_ms_per_rank'])) for r in d.get('strip_refinement') or []])
    except Exception as e: print(f, 'ERR', e)
PY
