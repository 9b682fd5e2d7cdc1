set -x
python -c "import restir_b200 as rb; print('build', rb.api.build_id())"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or fused or reference_order or traced_tree or north_star or device_built or strips_equal" 2>&1 | tail -15 > gpurun_out/r02_c10_pytest.txt; cat gpurun_out/r02_c10_pytest.txt
for w in config4_1080p config3 config4; do
  timeout 300 python bench.py --workload $w --steps 40 --warmup 8 --quick > gpurun_out/r02_c10_bench_$w.json 2> gpurun_out/r02_c10_bench_$w.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r02_c10_bench_$w.json').read().strip().splitlines()[-1]); print('$w', d['ms_per_step'], d['stage_ms'], d['e2e']['ms_per_step'], d.get('build_id'))"
done
export RSTR_LIBNAME=librestir_b200_stats.so
timeout 300 python scripts/gpu_shadow_stats.py config4_1080p 2>&1 | tail -5 | tee gpurun_out/r02_c10_shadow_stats.txt
timeout 300 python scripts/gpu_shadow_stats.py config4 2>&1 | tail -5 | tee -a gpurun_out/r02_c10_shadow_stats.txt
timeout 300 python scripts/gpu_shadow_stats.py config4 1252 1431 2>&1 | tail -5 | tee -a gpurun_out/r02_c10_shadow_stats.txt
timeout 300 python scripts/gpu_shadow_stats.py config3 2>&1 | tail -5 | tee -a gpurun_out/r02_c10_shadow_stats.txt
