# round-2 GPU call 5: precise ties + deferred-leaf shadow kernel; register / threshold A/B; reference CUDA build comparison (incl. -fmad=true experiment)
set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r02_c5_pytest.txt; tail -15 gpurun_out/r02_c5_pytest.txt
for lib in librestir_b200.so librestir_b200_p6.so librestir_b200_p5.so librestir_b200_p4.so librestir_b200_l4.so librestir_b200_l16.so librestir_b200_r16.so; do
  echo "== $lib config4_1080p"
  RSTR_LIBNAME=$lib python bench.py --workload config4_1080p --pipeline staged --steps 40 --warmup 8 --no-cpu-baseline --no-targets 2>gpurun_out/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stage_ms'], d['e2e']['ms_per_step'])" || tail -5 gpurun_out/err.txt
done > gpurun_out/r02_c5_ab.txt 2>&1
grep -v "^+" gpurun_out/r02_c5_ab.txt
for w in config3 config2; do for pl in staged fused; do echo "== $w $pl"; python bench.py --workload $w --pipeline $pl --steps 40 --warmup 8 --no-cpu-baseline --no-targets 2>gpurun_out/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stage_ms'], d['e2e']['ms_per_step'])"; done; done > gpurun_out/r02_c5_ab2.txt 2>&1
grep -v "^+" gpurun_out/r02_c5_ab2.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_c5_launches_config4_1080p.csv python bench.py --workload config4_1080p --pipeline staged --steps 6 --warmup 3 --no-cpu-baseline --no-targets > gpurun_out/ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_c5_launches_config4_1080p.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; k=h.index('Kernel Name'); v=h.index('Metric Value')
agg={}
for r in rows[hdr+1:]:
    if len(r)>v:
        a=agg.setdefault(r[k][:40],[0,0.0]); a[0]+=1; a[1]+=float(r[v].replace(',',''))
for n,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%-42s launches %3d  avg %.1f us" % (n,c,t/c/1000.0))
PY
python - <<'PY' > gpurun_out/r02_c5_fallback.txt 2>&1
import sys; sys.path.insert(0,'.')
import restir_b200 as rb, bench
rb.init(0)
for w in ("config4_1080p","config3","config2"):
    desc, spec, res, reuse, radius = bench.WORKLOADS[w]
    sd = bench.make_scene(spec,res); sc = rb.Scene.from_arrays(sd); base = rb.Camera.from_scene(sd)
    fr = sc.frame(*res); fr.set_pipeline(True); prm = rb.default_params(reuse=reuse, radius=radius)
    for k in range(10):
        cam = base.orbit(k); fr.gbuffer_render(cam); fr.restir_direct(cam, prm, k, 0); fr.gbuffer_update(cam)
    fr.sync(); print(w, "pixels recomputed by the reference-order walk over 10 frames:", sc.fallback_rays(), flush=True)
    fr.close(); sc.close()
PY
cat gpurun_out/r02_c5_fallback.txt
for w in config2 config3; do
  timeout 600 python scripts/ref_cuda_compare.py $w 30 > gpurun_out/r02_c5_ref_cuda_$w.json 2>gpurun_out/err.txt || tail -5 gpurun_out/err.txt
  RSTR_LIBNAME=librestir_b200_fmad.so timeout 600 python scripts/ref_cuda_compare.py $w 30 > gpurun_out/r02_c5_ref_cuda_${w}_fmad_true.json 2>gpurun_out/err.txt || tail -5 gpurun_out/err.txt
done
timeout 900 python scripts/ref_cuda_compare.py config4_1080p 20 > gpurun_out/r02_c5_ref_cuda_config4_1080p.json 2>gpurun_out/err.txt || tail -5 gpurun_out/err.txt
cat gpurun_out/r02_c5_ref_cuda_*.json
