# round-2 GPU call 35 (1 GPU, the last ~80 GPU-seconds of the round): the ray-queue form of ReSTIR GI (persistent walkers with lane refill) --
# parity on the GPU (queued == staged == one kernel), device timing next to the staged form, one ncu capture of a frame's walker launches
set -x
timeout -k 5 30 python -m pytest tests/test_restir_gi.py -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r02_c35_pytest_gi.txt; cat gpurun_out/r02_c35_pytest_gi.txt
cp gpurun_out/gi_parity.jsonl gpurun_out/r02_c35_gi_parity.jsonl 2>/dev/null
timeout -k 5 30 python scripts/gi_bench.py --workloads config3 config4_1080p config2 --steps 10 --modes staged queued --out gpurun_out/r02_c35_gi_bench.jsonl 2> gpurun_out/r02_c35_gi_bench.err
mkdir -p /tmp/rep
timeout -k 5 28 ncu --set full --clock-control none --import-source on -k regex:k_gi_walk -s 5 -c 5 -f -o /tmp/rep/gi_queued_config3 python scripts/gi_bench.py --workloads config3 --steps 1 --warmup 1 --modes queued --out gpurun_out/ncu_gi.jsonl > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log | cut -c1-200
timeout -k 5 15 python scripts/ncu_summary.py /tmp/rep/gi_queued_config3.ncu-rep k_gi_walk_closest > gpurun_out/r02_c35_prof_k_gi_walk_config3.summary.txt 2>&1
