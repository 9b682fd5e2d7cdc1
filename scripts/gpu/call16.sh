# round-2 GPU call 16 (8 GPUs): strips over 8 GPUs after the k_shadow changes -- bit-exactness at 4K, scaling numbers, BASELINE config 5 sweep
set -x
python -c "import restir_b200 as rb; print('build', rb.api.build_id())"
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $TR8 scripts/verify_multigpu.py config4 3 > gpurun_out/r02_c16_verify_config4_n8.json 2> gpurun_out/r02_c16_verify.err; echo "verify rc=$?"; tail -c 700 gpurun_out/r02_c16_verify_config4_n8.json
timeout 600 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_c16_bench_config4_n8.json 2> gpurun_out/r02_c16_bench8.err; echo "bench8 rc=$?"; tail -3 gpurun_out/r02_c16_bench8.err | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --quick > gpurun_out/r02_c16_bench_config4_n1.json 2> gpurun_out/r02_c16_bench1.err; echo "bench1 rc=$?"
timeout 400 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 --workload config2 > gpurun_out/r02_c16_bench_config2_n8.json 2> gpurun_out/r02_c16_bench2.err; echo "bench2 rc=$?"
timeout 900 $TR8 scripts/config5_sweep.py config4 1024 > gpurun_out/r02_c16_config5_n8.json 2> gpurun_out/r02_c16_config5.err; echo "sweep rc=$?"; tail -c 1500 gpurun_out/r02_c16_config5_n8.json
python - <<'PY'
import json
for f in ("r02_c16_bench_config4_n1","r02_c16_bench_config4_n8","r02_c16_bench_config2_n8"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'miss', d['halo_miss'], d['strips']['strip_bounds'] if 'strips' in d else None, d.get('invalid'))
        print('   per rank', d['stage_ms_per_rank'])
        print('   refine', d.get('strip_refinement'))
    except Exception as e: print(f, 'ERR', e)
PY
