# round-2 GPU call 25 (1 GPU): the evidence run again (call 23's gpurun_out exceeded the 64 MiB that travel back): summaries are made on the box, reports deleted
set -x
python -c "import restir_b200 as rb; print('build', rb.api.build_id())"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_c25_pytest_gpu.txt; cat gpurun_out/r02_c25_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r02_c25_bench_default.json 2> gpurun_out/r02_c25_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_c25_bench_reference_arm.json 2> gpurun_out/r02_c25_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_c25_launches_config4.csv python bench.py --steps 4 --warmup 3 --quick > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log | cut -c1-200
mkdir -p /tmp/rep
for kn in k_primary k_shadow k_candidates k_temporal k_restir_b; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s 5 -c 1 -f -o /tmp/rep/${kn}_config4 python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log | cut -c1-200
  python scripts/ncu_summary.py /tmp/rep/${kn}_config4.ncu-rep $kn > gpurun_out/r02_c25_prof_${kn}_config4.summary.txt 2>&1
done
for kn in k_primary k_shadow; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s 5 -c 1 -f -o /tmp/rep/${kn}_config4_1080p python bench.py --workload config4_1080p --steps 3 --warmup 3 --quick > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log | cut -c1-200
  python scripts/ncu_summary.py /tmp/rep/${kn}_config4_1080p.ncu-rep $kn > gpurun_out/r02_c25_prof_${kn}_config4_1080p.summary.txt 2>&1
done
cp /tmp/rep/k_primary_config4.ncu-rep gpurun_out/r02_c25_k_primary_config4.ncu-rep
du -sh gpurun_out
