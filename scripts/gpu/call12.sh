set -x
python -c "import restir_b200 as rb; print('build', rb.api.build_id())"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or fused or north_star or knob or edge" 2>&1 | tail -5 > gpurun_out/r02_c12_pytest.txt; cat gpurun_out/r02_c12_pytest.txt
for rep in 1 2; do
for lib in librestir_b200.so librestir_b200_nocoop.so; do
for w in config4_1080p config4; do
  RSTR_LIBNAME=$lib timeout 300 python bench.py --workload $w --steps 40 --warmup 8 --quick > gpurun_out/r02_c12_bench_${w}_$lib.json 2> gpurun_out/r02_c12_bench_$w.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r02_c12_bench_${w}_$lib.json').read().strip().splitlines()[-1]); print('$lib $w', d['ms_per_step'], d['stage_ms'], d['e2e']['ms_per_step'], d.get('build_id'))"
done
done
done
for lib in librestir_b200.so librestir_b200_nocoop.so; do
  RSTR_LIBNAME=$lib timeout 300 python scripts/gpu_shadow_stats.py config4 1252 1431 2>&1 | tail -1 | tee -a gpurun_out/r02_c12_strip.txt
  RSTR_LIBNAME=$lib timeout 300 python scripts/gpu_shadow_stats.py config4 0 664 2>&1 | tail -1 | tee -a gpurun_out/r02_c12_strip.txt
done
