# final 1-GPU check of the committed state: whole GPU suite + the default bench line
set -x
python -c "import restir_b200 as rb; print('build', rb.api.build_id())"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_c32_pytest_gpu.txt; cat gpurun_out/r02_c32_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/r02_c32_bench_default.json 2> gpurun_out/r02_c32_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02_c32_bench_default.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['halo_miss'], d['roofline']['traffic'], d['targets']['config4_1080p']['ms_per_frame'], d['build_id'])"
python -c "import __graft_entry__ as g; g.smoke()"
