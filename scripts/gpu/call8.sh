# round-2 GPU call 8: band-pipelined staged phase A -- parity, A/B over the number of bands and k_primary's register cap
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02_c8_pytest.txt; tail -6 gpurun_out/r02_c8_pytest.txt
for b in 1 2 4 8; do for w in config4_1080p config3; do echo "== bands $b $w"; python bench.py --workload $w --bands $b --steps 40 --warmup 8 --quick 2>gpurun_out/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stage_ms'], d['e2e']['ms_per_step'])"; done; done > gpurun_out/r02_c8_ab.txt 2>&1
grep -v "^+" gpurun_out/r02_c8_ab.txt
for lib in librestir_b200_p9.so librestir_b200_p10.so; do echo "== $lib"; RSTR_LIBNAME=$lib python bench.py --workload config4_1080p --steps 40 --warmup 8 --quick 2>gpurun_out/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stage_ms'], d['e2e']['ms_per_step'])"; done > gpurun_out/r02_c8_ab2.txt 2>&1
grep -v "^+" gpurun_out/r02_c8_ab2.txt
python bench.py --workload config4 --steps 30 --warmup 5 --quick > gpurun_out/r02_c8_bench_config4.json 2>gpurun_out/err.txt; python -c "import json; d=json.loads(open('gpurun_out/r02_c8_bench_config4.json').read().strip().splitlines()[-1]); print('config4 4K', d['ms_per_step'], d['stage_ms'], d['e2e']['ms_per_step'])"
