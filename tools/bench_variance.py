import sys, os, json, subprocess
cfg = sys.argv[1]
for s in (1, 2, 3):
    vals = []
    for rep in range(5):
        env = dict(os.environ, IMK_STREAMS=str(s))
        out = subprocess.run([sys.executable, "bench.py", "--steps", "10", "--no-cpu-baseline", "--no-other-configs", "--config", cfg], capture_output=True, text=True, env=env).stdout
        d = json.loads(out.strip().splitlines()[-1])
        vals.append(round(d["value"]))
    print(cfg, "streams", s, vals, flush=True)
