"""Training-step time with K3 of sub-batch i overlapped with K2 of sub-batch i + 1 (k3_overlap), against the
sub-batch size and the SM partition.  Arguments: sub_tiles:k3_sms pairs (k3_sms 0 = no overlap)."""
import os, sys, subprocess, json
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for spec in sys.argv[1:]:
    tiles, k3 = spec.split(":")
    env = dict(os.environ, CNB_SUB_TILES=tiles)
    if int(k3) > 0:
        env.update(CNB_K3_OVERLAP="1", CNB_K3_SMS=k3)
    out = subprocess.run([sys.executable, "bench.py", "--steps", "10", "--warmup", "3", "--no-cpu-baseline", "--quick"], env=env,
                         capture_output=True, text=True, cwd=root)
    try:
        d = json.loads([l for l in out.stdout.strip().splitlines() if l.startswith("{")][-1])
        print(spec, round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()}, flush=True)
    except Exception as e:
        print(spec, "failed", out.stderr[-500:], flush=True)
