set -x
python -c "import restir_b200 as rb; print('build', rb.api.build_id())"
timeout 900 python -m pytest tests/test_denoiser.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_c17_pytest_denoiser.txt; cat gpurun_out/r02_c17_pytest_denoiser.txt
timeout 1200 python -m pytest tests/test_ref_cuda.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_c17_pytest_refcuda.txt; cat gpurun_out/r02_c17_pytest_refcuda.txt
for w in config2 config3; do RSTR_LIBNAME=librestir_b200_fmad.so timeout 600 python scripts/ref_cuda_compare.py denoisers $w 6 > gpurun_out/r02_c17_denoisers_vs_ref_cuda_$w.json 2> gpurun_out/r02_c17_dn_$w.err; tail -c 1500 gpurun_out/r02_c17_denoisers_vs_ref_cuda_$w.json; tail -3 gpurun_out/r02_c17_dn_$w.err; done
timeout 600 python scripts/gpu_strip_probe.py config4 0,639,837,1053,1249,1429,1639,1878,2160 > gpurun_out/r02_c17_probe.txt 2>&1; tail -10 gpurun_out/r02_c17_probe.txt
timeout 900 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r02_c17_probe_launches.csv python scripts/gpu_strip_probe.py config4 0,639,837,1053,1249,1429,1639,1878,2160 > gpurun_out/r02_c17_probe_ncu.txt 2>&1
tail -3 gpurun_out/r02_c17_probe_ncu.txt
