set -x
nvidia-smi --query-gpu=index,name --format=csv | head -3
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "device_built" 2>&1 | tail -15 > gpurun_out/r02_c9_pytest.txt; cat gpurun_out/r02_c9_pytest.txt
for t in host gpu; do
  timeout 300 python bench.py --workload config4_1080p --steps 40 --warmup 8 --quick --traced-tree $t > gpurun_out/r02_c9_bench_1080p_$t.json 2> gpurun_out/r02_c9_bench_$t.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r02_c9_bench_1080p_$t.json').read().strip().splitlines()[-1]); print('$t', d['ms_per_step'], d['stage_ms'], d.get('traced_tree'))"
done
timeout 300 python bench.py --workload config3 --steps 40 --warmup 8 --quick --traced-tree gpu > gpurun_out/r02_c9_bench_config3_gpu.json 2>> gpurun_out/r02_c9_bench_gpu.err
python -c "import sys,json; d=json.loads(open('gpurun_out/r02_c9_bench_config3_gpu.json').read().strip().splitlines()[-1]); print('config3 gpu', d['ms_per_step'], d['stage_ms'], d.get('traced_tree'))"
timeout 600 python scripts/gpu_strip_probe.py > gpurun_out/r02_c9_probe.txt 2>&1; tail -12 gpurun_out/r02_c9_probe.txt
timeout 900 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r02_c9_probe_launches.csv python scripts/gpu_strip_probe.py > gpurun_out/r02_c9_probe_ncu.txt 2>&1
tail -3 gpurun_out/r02_c9_probe_ncu.txt
