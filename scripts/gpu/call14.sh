set -x
for lib in librestir_b200_ts_coop.so librestir_b200_ts_nocoop.so; do
  for a in "config4_1080p" "config4" "config4 1252 1431" "config3"; do
    RSTR_LIBNAME=$lib timeout 300 python scripts/gpu_shadow_stats.py $a 2>&1 | tail -4 | grep -E "kernel|stage" | sed "s/^/$lib $a /" | tee -a gpurun_out/r02_c14_tail.txt
  done
done
for rep in 1 2; do
for lib in librestir_b200.so librestir_b200_nocoop.so; do
for w in config4_1080p config4; do
  RSTR_LIBNAME=$lib timeout 300 python bench.py --workload $w --steps 40 --warmup 8 --quick > gpurun_out/r02_c14_bench_${w}_$lib.json 2> gpurun_out/r02_c14_bench_$w.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r02_c14_bench_${w}_$lib.json').read().strip().splitlines()[-1]); print('$lib $w', d['ms_per_step'], d['stage_ms'], d['e2e']['ms_per_step'], d.get('build_id'))" | tee -a gpurun_out/r02_c14_ab.txt
done
done
done
