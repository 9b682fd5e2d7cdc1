"""A/B timing of kernel variants inside ONE process (same GPU, same clocks): env toggles are read per process, so the
variants run in subprocesses back to back.  Usage: gpu_ab.py <workload> VAR=a,b [VAR2=c,d]"""
import itertools, json, os, subprocess, sys
work = sys.argv[1]
axes = [(a.split("=")[0], a.split("=")[1].split(",")) for a in sys.argv[2:]]
for combo in itertools.product(*[v for _, v in axes]):
    env = dict(os.environ)
    for (k, _), v in zip(axes, combo):
        env[k] = v
    out = subprocess.run([sys.executable, "bench.py", "--steps", "30", "--warmup", "8", "--workload", work, "--no-cpu-baseline"], capture_output=True, text=True, env=env)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(dict(zip([k for k, _ in axes], combo)), "ms/frame %.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["stage_ms"].items()}, "e2e %.3f" % d["e2e"]["ms_per_step"], "clk", d["clocks"]["sm_mhz"], flush=True)
    except Exception as e:
        print(combo, "FAILED", e, out.stderr[-800:])
