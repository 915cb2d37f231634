"""Run GPU tests so that one hung kernel cannot eat the whole gpurun budget: every pytest process gets a wall-clock
limit; when it is exceeded the test that was running is recorded as HUNG and the remaining tests are re-run in a fresh
process. Usage: python tools/gpu_pytest.py [--per-proc 150] [--log gpurun_out/x.log] <pytest args / node ids>"""
import argparse
import os
import re
import subprocess
import sys
import time

ap = argparse.ArgumentParser()
ap.add_argument("--per-test", type=float, default=150.0, help="seconds without a new START line before a kill")
ap.add_argument("--log", default="gpurun_out/gpu_pytest.log")
ap.add_argument("rest", nargs=argparse.REMAINDER)
args = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.dirname(os.path.join(root, args.log)) or ".", exist_ok=True)
env = dict(os.environ, MPL_TEST_TRACE="1", PYTHONUNBUFFERED="1")
col = subprocess.run([sys.executable, "-m", "pytest", "--collect-only", "-q", "-p", "no:cacheprovider"] + args.rest,
                     cwd=root, capture_output=True, text=True, env=env)
ids = [l.strip() for l in col.stdout.splitlines() if "::" in l]
results = {}
log = open(os.path.join(root, args.log), "w")
remaining = list(ids)
while remaining:
    p = subprocess.Popen([sys.executable, "-u", "-m", "pytest", "-q", "-rfE", "--no-header", "-p", "no:cacheprovider",
                          "--tb=short"] + remaining, cwd=root, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         env=env)
    os.set_blocking(p.stdout.fileno(), False)
    current, last, buf, first = None, time.time(), "", True
    hung = False
    while True:
        try:
            raw = os.read(p.stdout.fileno(), 1 << 16)
        except BlockingIOError:
            raw = b""
        chunk = raw.decode("utf-8", "replace")
        if chunk:
            buf += chunk
            log.write(chunk); log.flush()
            for m in re.finditer(r"START (\S+)", chunk):
                if current is not None:
                    results.setdefault(current, "done")
                current, last, first = m.group(1), time.time(), False
        if p.poll() is not None:
            try:
                raw = os.read(p.stdout.fileno(), 1 << 20)
                tail = raw.decode("utf-8", "replace")
                buf += tail
                log.write(tail); log.flush()
            except (BlockingIOError, OSError):
                pass
            break
        limit = 240.0 if first else args.per_test  # the first import of torch on a fresh box is slow
        if time.time() - last > limit:
            p.kill(); p.wait()
            hung = True
            break
        time.sleep(0.2)
    if current is not None and not hung:
        results.setdefault(current, "done")
    for m in re.finditer(r"^RESULT (FAILED) (\S+)", buf, re.M):
        results[m.group(2)] = m.group(1)
    for m in re.finditer(r"^(FAILED|ERROR) (\S+)", buf, re.M):
        results[m.group(2)] = m.group(1)
    if hung and current is not None:
        results[current] = "HUNG"
        log.write(f"\n*** HUNG: {current}\n"); log.flush()
        remaining = remaining[remaining.index(current) + 1:] if current in remaining else []
    else:
        remaining = []
bad = {k: v for k, v in results.items() if v != "done"}
summary = f"\n=== gpu_pytest: {len(ids)} collected, {sum(v == 'done' for v in results.values())} passed, {len(bad)} bad\n"
for k, v in bad.items():
    summary += f"{v} {k}\n"
log.write(summary); log.close()
print(summary)
sys.exit(1 if bad else 0)
