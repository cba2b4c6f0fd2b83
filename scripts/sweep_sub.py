"""Training-step time against the sub-batch size (CNB_SUB_TILES; set before the library is loaded)."""
import os, sys, subprocess, json
for tiles in sys.argv[1:]:
    env = dict(os.environ, CNB_SUB_TILES=tiles)
    out = subprocess.run([sys.executable, "bench.py", "--steps", "10", "--warmup", "3", "--no-cpu-baseline", "--quick"], env=env,
                         capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(tiles, round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()}, flush=True)
    except Exception as e:
        print(tiles, "failed", out.stderr[-500:], flush=True)
