# round-2 GPU call 23 (1 GPU): the evidence run -- whole GPU suite, default bench + reference arm, ncu launch list and full captures of the frame's kernels
set -x
python -c "import restir_b200 as rb; print('build', rb.api.build_id())"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_c23_pytest_gpu.txt; cat gpurun_out/r02_c23_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r02_c23_bench_default.json 2> gpurun_out/r02_c23_bench_default.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_c23_bench_default.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_c23_bench_reference_arm.json 2> gpurun_out/r02_c23_ref.err; echo "ref rc=$?"; tail -c 400 gpurun_out/r02_c23_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_c23_launches_config4.csv python bench.py --steps 4 --warmup 3 --quick > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log | cut -c1-200
for kn in k_primary k_shadow k_candidates k_temporal k_restir_b; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s 5 -c 1 -f -o gpurun_out/r02_c23_${kn}_config4 python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log | cut -c1-200
done
for kn in k_primary k_shadow; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s 5 -c 1 -f -o gpurun_out/r02_c23_${kn}_config4_1080p python bench.py --workload config4_1080p --steps 3 --warmup 3 --quick > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep | tail -8
