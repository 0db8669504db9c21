"""Run GPU tests one test-function per process (a CUDA fault in one kernel cannot poison the rest), each under a
timeout, and write a summary to gpurun_out/<name>.log.  Bring-up helper; the driver runs plain `pytest -m gpu`.

    python tools/gpu_run_isolated.py tests/test_kernels_gpu.py [-k expr] [--timeout 240]
"""
import argparse
import os
import subprocess
import sys
import time

ap = argparse.ArgumentParser()
ap.add_argument("paths", nargs="+")
ap.add_argument("-k", default=None)
ap.add_argument("--timeout", type=int, default=240)
ap.add_argument("--name", default="isolated")
args = ap.parse_args()

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
cmd = [sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu", *args.paths]
if args.k:
    cmd += ["-k", args.k]
out = subprocess.run(cmd, cwd=root, capture_output=True, text=True).stdout
funcs = []
for line in out.splitlines():
    if "::" in line:
        f = line.split("[")[0].strip()
        if f not in funcs:
            funcs.append(f)
log = open(os.path.join(root, "gpurun_out", args.name + ".log"), "w")
summary = []
for f in funcs:
    t0 = time.time()
    try:
        p = subprocess.run([sys.executable, "-m", "pytest", f, "-q", "-m", "gpu", "-x", "--no-header", "-rA"], cwd=root,
                           capture_output=True, text=True, timeout=args.timeout)
        rc, txt = p.returncode, p.stdout[-6000:] + p.stderr[-2000:]
    except subprocess.TimeoutExpired as e:
        rc, txt = -9, "TIMEOUT\n" + ((e.stdout or b"").decode()[-3000:] if isinstance(e.stdout, bytes) else str(e.stdout)[-3000:])
    dt = time.time() - t0
    status = "PASS" if rc == 0 else ("TIMEOUT" if rc == -9 else "FAIL")
    summary.append(f"{status:8s} {dt:6.1f}s {f}")
    log.write(f"===== {status} {f} ({dt:.1f}s)\n{txt}\n")
    log.flush()
    print(summary[-1], flush=True)
log.write("\n".join(summary) + "\n")
log.close()
print("\n".join(summary))
sys.exit(0 if all(s.startswith("PASS") for s in summary) else 1)
