# round-2 GPU call 6: 256-bit node / light loads, in-process strip-group test with split calls, contracted-build test; ncu of the staged kernels
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02_c6_pytest.txt; tail -12 gpurun_out/r02_c6_pytest.txt
for w in config4_1080p config3 config2; do echo "== $w"; python bench.py --workload $w --steps 40 --warmup 8 --no-cpu-baseline --no-targets 2>gpurun_out/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stage_ms'], d['e2e']['ms_per_step'])"; done > gpurun_out/r02_c6_ab.txt 2>&1
grep -v "^+" gpurun_out/r02_c6_ab.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_c6_launches_config4_1080p.csv python bench.py --workload config4_1080p --steps 6 --warmup 3 --no-cpu-baseline --no-targets > gpurun_out/ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_c6_launches_config4_1080p.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; k=h.index('Kernel Name'); v=h.index('Metric Value')
agg={}
for r in rows[hdr+1:]:
    if len(r)>v:
        a=agg.setdefault(r[k][:40],[0,0.0]); a[0]+=1; a[1]+=float(r[v].replace(',',''))
for n,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%-42s launches %3d  avg %.1f us" % (n,c,t/c/1000.0))
PY
for kn in k_primary k_shadow k_candidates; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s 6 -c 1 -f -o gpurun_out/r02_c6_${kn}_config4_1080p python bench.py --workload config4_1080p --steps 3 --warmup 3 --no-cpu-baseline --no-targets > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log
done
