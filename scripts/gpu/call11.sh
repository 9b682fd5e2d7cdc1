set -x
python -c "import restir_b200 as rb; print('build', rb.api.build_id())"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or fused or reference_order or traced_tree or north_star or device_built or strips_equal or knob or multi_pass or edge or unbiased_mode_against" 2>&1 | tail -15 > gpurun_out/r02_c11_pytest.txt; cat gpurun_out/r02_c11_pytest.txt
for lib in librestir_b200.so librestir_b200_nocoop.so; do
for w in config4_1080p config3 config4; do
  RSTR_LIBNAME=$lib timeout 300 python bench.py --workload $w --steps 40 --warmup 8 --quick > gpurun_out/r02_c11_bench_${w}_$lib.json 2> gpurun_out/r02_c11_bench_$w.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r02_c11_bench_${w}_$lib.json').read().strip().splitlines()[-1]); print('$lib $w', d['ms_per_step'], d['stage_ms'], d['e2e']['ms_per_step'], d.get('build_id'))"
done
done
for t in gpu gpu-radix; do
  timeout 300 python bench.py --workload config4_1080p --steps 40 --warmup 8 --quick --traced-tree $t > gpurun_out/r02_c11_bench_1080p_tree_$t.json 2> gpurun_out/r02_c11_bench_tree_$t.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r02_c11_bench_1080p_tree_$t.json').read().strip().splitlines()[-1]); print('$t', d['ms_per_step'], d['stage_ms'], d.get('traced_tree'))"
done
tail -3 gpurun_out/r02_c11_bench_tree_gpu.err
export RSTR_LIBNAME=librestir_b200_stats.so
timeout 300 python scripts/gpu_shadow_stats.py config4_1080p 2>&1 | tail -5 | tee gpurun_out/r02_c11_shadow_stats.txt
timeout 300 python scripts/gpu_shadow_stats.py config4 1252 1431 2>&1 | tail -5 | tee -a gpurun_out/r02_c11_shadow_stats.txt
