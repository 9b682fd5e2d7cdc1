# round-2 GPU call 2: packet traversal -- parity suite, A/B over the fused kernel's register cap, ncu capture
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_c2_pytest.txt; cat gpurun_out/r02_c2_pytest.txt
grep -q "passed" gpurun_out/r02_c2_pytest.txt || echo "TESTS DID NOT PASS"
for lib in librestir_b200.so librestir_b200_m4.so librestir_b200_m5.so librestir_b200_m8.so; do
  for w in config4_1080p config3; do
    echo "== $lib $w"
    RSTR_LIBNAME=$lib python bench.py --workload $w --steps 40 --warmup 8 --no-cpu-baseline --no-targets 2>gpurun_out/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stage_ms'], d['e2e']['ms_per_step'])" || tail -5 gpurun_out/err.txt
  done
done > gpurun_out/r02_c2_ab.txt 2>&1
cat gpurun_out/r02_c2_ab.txt
python bench.py --workload config2 --steps 60 --warmup 8 --no-cpu-baseline --no-targets > gpurun_out/r02_c2_bench_config2.json 2>gpurun_out/err.txt; tail -c 400 gpurun_out/r02_c2_bench_config2.json
python bench.py --workload config4_1080p --steps 60 --warmup 8 --no-cpu-baseline --no-targets > gpurun_out/r02_c2_bench_config4_1080p.json 2>gpurun_out/err.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gbuffer_restir_a -s 10 -c 1 -f -o gpurun_out/r02_c2_fused_config4_1080p python bench.py --workload config4_1080p --steps 3 --warmup 3 --no-cpu-baseline --no-targets > gpurun_out/ncu.log 2>&1; tail -3 gpurun_out/ncu.log
