# round-2 GPU call 34 (1 GPU, the last ~3 GPU-minutes of the round): ReSTIR GI as a wavefront -- parity (staged == one kernel, GPU vs GPU, and
# vs fixture / oracle), device timing of both forms, launch-bounds variants of k_gi_bounce, one ncu capture of the three k_gi_bounce launches
set -x
python -c "import restir_b200 as rb; print('build', rb.api.build_id())"
timeout -k 5 70 python -m pytest tests/test_restir_gi.py -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r02_c34_pytest_gi.txt; cat gpurun_out/r02_c34_pytest_gi.txt
timeout -k 5 45 python scripts/gi_bench.py --workloads config3 config4_1080p config2 --steps 10 --modes traced staged --out gpurun_out/r02_c34_gi_bench.jsonl 2> gpurun_out/r02_c34_gi_bench.err
for v in gib6 gib8; do
  RSTR_LIBNAME=librestir_b200_$v.so timeout -k 5 25 python scripts/gi_bench.py --workloads config3 config4_1080p --steps 10 --modes staged --out gpurun_out/r02_c34_gi_bench_$v.jsonl 2>> gpurun_out/r02_c34_gi_bench.err
done
mkdir -p /tmp/rep
timeout -k 5 50 ncu --set full --clock-control none --import-source on -k regex:k_gi_ -s 5 -c 5 -f -o /tmp/rep/gi_staged_config3 python scripts/gi_bench.py --workloads config3 --steps 1 --warmup 1 --modes staged --out gpurun_out/ncu_gi.jsonl > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log | cut -c1-200
timeout -k 5 25 python scripts/ncu_summary.py /tmp/rep/gi_staged_config3.ncu-rep k_gi_bounce > gpurun_out/r02_c34_prof_k_gi_staged_config3.summary.txt 2>&1
RSTR_GI_PIPELINE=staged timeout -k 5 40 python -m pytest tests/test_restir_gi.py -m gpu -q --tb=short 2>&1 | tail -15 > gpurun_out/r02_c34_pytest_gi_staged_default.txt; cat gpurun_out/r02_c34_pytest_gi_staged_default.txt
cp gpurun_out/gi_parity.jsonl gpurun_out/r02_c34_gi_parity.jsonl 2>/dev/null
cp /tmp/rep/gi_staged_config3.ncu-rep gpurun_out/r02_c34_k_gi_staged_config3.ncu-rep 2>/dev/null
du -sh gpurun_out
