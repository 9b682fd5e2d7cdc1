# round-2 GPU call 7 (8 GPUs): strips over 8 / 4 GPUs with the peer data plane -- bit-exactness at 4K and the scaling numbers
set -x
nvidia-smi --query-gpu=index,name --format=csv | head -3
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522"
timeout 600 $TR8 scripts/verify_multigpu.py config4 3 > gpurun_out/r02_c7_verify_config4_n8.json 2> gpurun_out/r02_c7_verify.err; echo "verify rc=$?"; tail -c 700 gpurun_out/r02_c7_verify_config4_n8.json
timeout 600 $TR8 bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/r02_c7_bench_config4_n8.json 2> gpurun_out/r02_c7_bench8.err; echo "bench8 rc=$?"; tail -3 gpurun_out/r02_c7_bench8.err | cut -c1-300
timeout 600 $TR4 bench.py --gpus 4 --steps 40 --warmup 5 > gpurun_out/r02_c7_bench_config4_n4.json 2> gpurun_out/r02_c7_bench4.err; echo "bench4 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 40 --warmup 5 --quick > gpurun_out/r02_c7_bench_config4_n1.json 2> gpurun_out/r02_c7_bench1.err; echo "bench1 rc=$?"
timeout 400 $TR8 bench.py --gpus 8 --steps 60 --warmup 5 --workload config2 > gpurun_out/r02_c7_bench_config2_n8.json 2> gpurun_out/r02_c7_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for f in ("r02_c7_bench_config4_n1","r02_c7_bench_config4_n4","r02_c7_bench_config4_n8","r02_c7_bench_config2_n8"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'miss', d['halo_miss'], d['strips']['strip_bounds'], d.get('invalid'))
        print('   per rank', d['stage_ms_per_rank'])
    except Exception as e: print(f, 'ERR', e)
PY
