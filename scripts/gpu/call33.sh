# round-2 GPU call 33 (1 GPU, the last ~4 GPU-minutes of the round): ReSTIR GI -- parity tests, device timing, one ncu capture
set -x
python -c "import restir_b200 as rb; print('build', rb.api.build_id())"
timeout -k 5 100 python -m pytest tests/test_restir_gi.py -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r02_c33_pytest_gi.txt; cat gpurun_out/r02_c33_pytest_gi.txt
timeout -k 5 50 python scripts/gi_bench.py --workloads config2 config3 --steps 10 --modes traced exact --out gpurun_out/r02_c33_gi_bench.jsonl 2> gpurun_out/r02_c33_gi_bench.err
mkdir -p /tmp/rep
timeout -k 5 45 ncu --set full --clock-control none --import-source on -k regex:k_restir_indirect -s 2 -c 1 -f -o /tmp/rep/gi_config3 python scripts/gi_bench.py --workloads config3 --steps 1 --warmup 1 --modes traced --out gpurun_out/ncu_gi.jsonl > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log | cut -c1-200
timeout -k 5 20 python scripts/ncu_summary.py /tmp/rep/gi_config3.ncu-rep k_restir_indirect > gpurun_out/r02_c33_prof_k_restir_indirect_config3.summary.txt 2>&1
timeout -k 5 35 python scripts/gi_bench.py --workloads config4_1080p --steps 5 --modes traced --out gpurun_out/r02_c33_gi_bench.jsonl 2>> gpurun_out/r02_c33_gi_bench.err
timeout -k 5 25 python -c "import __graft_entry__ as g; g.smoke()"
cp /tmp/rep/gi_config3.ncu-rep gpurun_out/r02_c33_k_restir_indirect_config3.ncu-rep 2>/dev/null
du -sh gpurun_out
