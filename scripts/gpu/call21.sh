set -x
python -c "import restir_b200 as rb; print('build', rb.api.build_id())"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or fused or reference_order or traced_tree or north_star or device_built or strip or knob or edge or multi_pass" 2>&1 | tail -5 > gpurun_out/r02_c21_pytest.txt; cat gpurun_out/r02_c21_pytest.txt
run() { # lib workload tag
  RSTR_LIBNAME=$1 timeout 300 python bench.py --workload $2 --steps 40 --warmup 8 --quick > gpurun_out/r02_c21_bench_$3.json 2> gpurun_out/r02_c21_bench_$3.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r02_c21_bench_$3.json').read().strip().splitlines()[-1]); print('$3', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['stage_ms'].items()}, round(d['e2e']['ms_per_step'],4), d.get('build_id'))" | tee -a gpurun_out/r02_c21_ab.txt
}
for rep in 1 2; do
for v in "" _noq; do
  run librestir_b200$v.so config4_1080p 1080p$v
  run librestir_b200$v.so config3 config3$v
  run librestir_b200$v.so config4 4k$v
done
done
for lib in librestir_b200.so librestir_b200_noq.so; do RSTR_LIBNAME=$lib timeout 300 python scripts/gpu_shadow_stats.py config4 1252 1431 2>&1 | tail -1 | sed "s/^/$lib /" | tee -a gpurun_out/r02_c21_strip.txt; done
