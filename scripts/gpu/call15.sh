set -x
python -c "import restir_b200 as rb; print('build', rb.api.build_id())"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or fused or north_star or knob or edge or strip" 2>&1 | tail -5 > gpurun_out/r02_c15_pytest.txt; cat gpurun_out/r02_c15_pytest.txt
run() { # lib workload tag env
  env $4 RSTR_LIBNAME=$1 timeout 300 python bench.py --workload $2 --steps 40 --warmup 8 --quick > gpurun_out/r02_c15_bench_$3.json 2> gpurun_out/r02_c15_bench_$3.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r02_c15_bench_$3.json').read().strip().splitlines()[-1]); print('$3', d['ms_per_step'], d['stage_ms'], d['e2e']['ms_per_step'], d.get('build_id'))" | tee -a gpurun_out/r02_c15_ab.txt
}
for rep in 1 2; do
run librestir_b200.so config4_1080p main_1080p X=1
run librestir_b200.so config4_1080p main_1080p_drain RSTR_SHADOW_DRAIN=1
run librestir_b200_order.so config4_1080p order_1080p X=1
run librestir_b200.so config4 main_4k X=1
run librestir_b200_order.so config4 order_4k X=1
run librestir_b200.so config3 main_config3 X=1
run librestir_b200_order.so config3 order_config3 X=1
run librestir_b200.so config2 main_config2 X=1
done
for e in 0 1; do
  RSTR_SHADOW_DRAIN=$e timeout 300 python scripts/gpu_shadow_stats.py config4 1252 1431 2>&1 | tail -1 | sed "s/^/drain=$e /" | tee -a gpurun_out/r02_c15_strip.txt
done
