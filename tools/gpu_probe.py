"""Run each GPU test node in its own process (a trapped kernel poisons the CUDA context) with a timeout and
collect a compact report in gpurun_out/probe_<tag>.txt.   Usage: python tools/gpu_probe.py <tag> <pytest args...>"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
tag = sys.argv[1]
args = sys.argv[2:]
r = subprocess.run([sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu"] + args, capture_output=True,
                   text=True, cwd=ROOT)
nodes = [l.strip() for l in r.stdout.splitlines() if "::" in l]
out = open(os.path.join(ROOT, "gpurun_out", f"probe_{tag}.txt"), "w")
npass = 0
for n in nodes:
    t0 = time.time()
    try:
        p = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", n], capture_output=True, text=True,
                           cwd=ROOT, timeout=180)
        ok = p.returncode == 0
        tail = "" if ok else "\n".join((p.stdout + p.stderr).splitlines()[-40:])
    except subprocess.TimeoutExpired:
        ok, tail = False, "TIMEOUT"
    npass += ok
    line = f"{'PASS' if ok else 'FAIL'} {time.time() - t0:6.1f}s {n}"
    print(line, flush=True)
    out.write(line + "\n")
    if tail:
        out.write(tail + "\n")
        print(tail[-3000:], flush=True)
out.write(f"{npass}/{len(nodes)} passed\n")
print(f"{npass}/{len(nodes)} passed")
