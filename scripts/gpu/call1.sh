# round-2 GPU call 1: whole -m gpu suite, default bench + reference arm, motion bounds, sanitizers
set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
nproc
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_c1_pytest.txt; cat gpurun_out/r02_c1_pytest.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c1_bench_config4.json 2> gpurun_out/r02_c1_bench_config4.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_c1_bench_config4.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_c1_ref_config4.json 2> gpurun_out/r02_c1_ref.err; echo "ref rc=$?"
python - <<'PY' > gpurun_out/r02_c1_motion.txt 2>&1
import sys; sys.path.insert(0,'.')
import restir_b200 as rb, bench
rb.init(0)
for w in ("config4","config4_1080p","config3","config2"):
    desc, spec, res, reuse, radius = bench.WORKLOADS[w]
    sd = bench.make_scene(spec,res); sc = rb.Scene.from_arrays(sd); base = rb.Camera.from_scene(sd)
    print(w, "motion_rows", bench.measure_motion_rows(sc, base, res[0], res[1], rb), flush=True)
    sc.close()
PY
cat gpurun_out/r02_c1_motion.txt
bash scripts/sanitize.sh
