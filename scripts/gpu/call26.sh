set -x
RSTR_LIBNAME=librestir_b200_ppf.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or north_star" 2>&1 | tail -3 | tee gpurun_out/r02_c26_pytest.txt
run() { # lib workload tag
  RSTR_LIBNAME=$1 timeout 300 python bench.py --workload $2 --steps 40 --warmup 8 --quick > gpurun_out/r02_c26_bench_$3.json 2> gpurun_out/r02_c26_bench_$3.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r02_c26_bench_$3.json').read().strip().splitlines()[-1]); print('$3', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['stage_ms'].items()}, round(d['e2e']['ms_per_step'],4), d.get('build_id'))" | tee -a gpurun_out/r02_c26_ab.txt
}
for rep in 1 2; do
for v in "" _ppf; do
  run librestir_b200$v.so config4_1080p 1080p$v
  run librestir_b200$v.so config3 config3$v
  run librestir_b200$v.so config4 4k$v
done
done
