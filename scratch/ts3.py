import os, sys, subprocess
for env in ({}, {"EMME_DENSE_NBO": "128"}, {"EMME_DENSE_NBO": "256"}, {"EMME_DENSE_NBO": "64"}):
    r = subprocess.run([sys.executable, "profiles/time_small.py", "child"], env={**os.environ, **env}, capture_output=True, text=True)
    print(env, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:], flush=True)
