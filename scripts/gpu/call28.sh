set -x
run() { # lib tag
  RSTR_LIBNAME=$1 timeout 300 python bench.py --workload config4_1080p --steps 40 --warmup 8 --quick --traced-tree gpu > gpurun_out/r02_c28_bench_$2.json 2> gpurun_out/r02_c28_bench_$2.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r02_c28_bench_$2.json').read().strip().splitlines()[-1]); print('$2', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['stage_ms'].items()}, d.get('traced_tree'), d.get('build_id'))" | tee -a gpurun_out/r02_c28_ab.txt
}
run librestir_b200.so ploc16
run librestir_b200_ploc8.so ploc8
run librestir_b200_ploc32.so ploc32
run librestir_b200_ploc64.so ploc64
