"""A/B of library variants (build.py --variant=<name> -D...): quick train bench per variant, several rounds, in one GPU session."""
import os, sys, subprocess, json
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
variants = sys.argv[1:] or [""]
for rnd in range(2):
    for v in variants:
        env = dict(os.environ)
        env.pop("CNB_LIB", None)
        for part in v.split(","):             # "libsuffix" and / or "NAME=VALUE" environment switches, comma separated
            if "=" in part: k, x = part.split("=", 1); env[k] = x
            elif part: env["CNB_LIB"] = part
        out = subprocess.run([sys.executable, "bench.py", "--steps", os.environ.get("AB_STEPS", "10"), "--warmup", "3", "--no-cpu-baseline", "--quick"], env=env,
                             capture_output=True, text=True, cwd=root)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            print(f"{v or 'default':10s} step {d['ms_per_step']:.3f} ms", {k: round(x, 3) for k, x in d["kernel_ms_per_step"].items()},
                  f"fwd {d['fwd_rays_per_s']/1e6:.2f}M latent {d['latent_fit_rays_per_s']/1e6:.2f}M", flush=True)
        except Exception:
            print(v, "failed", out.stderr[-400:], flush=True)
