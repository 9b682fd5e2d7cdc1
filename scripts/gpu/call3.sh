# round-2 GPU call 3 (2 GPUs): the library's peer data plane -- in-process test, cross-process verification, bench A/B vs NCCL
set -x
nvidia-smi topo -m | head -6
timeout 600 python -m pytest tests -m gpu -x -q -k "strip_group or strips" 2>&1 | tail -8 > gpurun_out/r02_c3_pytest.txt; cat gpurun_out/r02_c3_pytest.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/verify_multigpu.py config3 4 > gpurun_out/r02_c3_verify_config3_n2.json 2> gpurun_out/r02_c3_verify.err; echo "verify rc=$?"; tail -c 600 gpurun_out/r02_c3_verify_config3_n2.json; tail -5 gpurun_out/r02_c3_verify.err
timeout 900 $TR scripts/verify_multigpu.py config4 3 > gpurun_out/r02_c3_verify_config4_n2.json 2> gpurun_out/r02_c3_verify4.err; echo "verify4 rc=$?"; tail -c 600 gpurun_out/r02_c3_verify_config4_n2.json; tail -5 gpurun_out/r02_c3_verify4.err
for ex in peer nccl; do
  timeout 900 $TR bench.py --gpus 2 --steps 30 --warmup 5 --exchange $ex > gpurun_out/r02_c3_bench_config4_n2_$ex.json 2> gpurun_out/r02_c3_bench_$ex.err; echo "bench $ex rc=$?"; tail -3 gpurun_out/r02_c3_bench_$ex.err
  python -c "import json; d=json.loads(open('gpurun_out/r02_c3_bench_config4_n2_$ex.json').read().strip().splitlines()[-1]); print('$ex', d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms_per_rank'], d['halo_miss'], d['strips'])"
  timeout 600 $TR bench.py --gpus 2 --steps 60 --warmup 5 --workload config2 --exchange $ex > gpurun_out/r02_c3_bench_config2_n2_$ex.json 2> gpurun_out/r02_c3_bench2_$ex.err; echo "bench2 $ex rc=$?"
  python -c "import json; d=json.loads(open('gpurun_out/r02_c3_bench_config2_n2_$ex.json').read().strip().splitlines()[-1]); print('$ex', d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms_per_rank'], d['halo_miss'])"
done
