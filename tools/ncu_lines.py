"""Join an ncu SASS source page with nvdisasm line info: stall samples / executed instructions per CUDA
source line of the step kernel.  Usage: python tools/ncu_lines.py <report.ncu-rep> [top N]
(needs build/kernels/step_kernel.cu.o from the same build as the profiled run)"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "build/kernels/step_kernel.cu.o")], cwd=tmp,
               check=True, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
addr2line = {}
cur = None
func = None
for ln in dis.splitlines():
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        func = m.group(1)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
    if m and func and "b2k_step_kernel" in func:
        addr2line[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
base = None
per = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]])
    if base is None:
        base = a
    key = addr2line.get(a - base, ("?", 0))

    def f(name):
        try:
            return float(r[ix[name]] or 0)
        except ValueError:
            return 0.0
    for name in ("# Samples", "Instructions Executed", "stall_long_sb", "stall_no_inst", "stall_wait", "stall_short_sb",
                 "stall_branch_resolving", "L2 Theoretical Sectors Local"):
        per[key][name] += f(name)
        tot[name] += f(name)
print(f"total samples {tot['# Samples']:.0f}  warp-instructions {tot['Instructions Executed']:.0f}")
byfile = collections.defaultdict(collections.Counter)
for (fn, ln), c in per.items():
    byfile[fn].update(c)
print("\nper file:")
for fn, c in sorted(byfile.items(), key=lambda kv: -kv[1]["# Samples"]):
    print(f"  {fn:26s} samples {100 * c['# Samples'] / tot['# Samples']:5.1f}%  inst {100 * c['Instructions Executed'] / tot['Instructions Executed']:5.1f}%"
          f"  long_sb {c['stall_long_sb']:.0f} no_inst {c['stall_no_inst']:.0f} wait {c['stall_wait']:.0f} short_sb {c['stall_short_sb']:.0f}")
print(f"\ntop {top} lines by stall samples:")
srccache = {}
for (fn, ln), c in sorted(per.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    if fn not in srccache:
        p = glob.glob(os.path.join(ROOT, "mujoco_ros_pkgs_b200/csrc/*", fn))
        srccache[fn] = open(p[0]).read().splitlines() if p else []
    text = srccache[fn][ln - 1].strip()[:90] if 0 < ln <= len(srccache[fn]) else ""
    print(f"{100 * c['# Samples'] / tot['# Samples']:5.1f}% inst {c['Instructions Executed']:8.0f} lsb {c['stall_long_sb']:5.0f} noi {c['stall_no_inst']:5.0f} "
          f"wait {c['stall_wait']:5.0f} ssb {c['stall_short_sb']:5.0f} loc {c['L2 Theoretical Sectors Local']:7.0f} | {fn}:{ln} {text}")

# ---- per-function aggregation (function = nearest preceding "__device__ ... name(" line in the file)
import bisect
funcs = {}
for fn in set(k[0] for k in per):
    p = glob.glob(os.path.join(ROOT, "mujoco_ros_pkgs_b200/csrc/*", fn))
    if not p:
        continue
    starts = []
    for i, l in enumerate(open(p[0]).read().splitlines(), 1):
        mm = re.match(r"\s*(?:B2K_DI|__device__|__global__)[^;(]*?\b(\w+)\s*\(", l)
        if mm:
            starts.append((i, mm.group(1)))
    funcs[fn] = starts
agg = collections.defaultdict(collections.Counter)
for (fn, ln), c in per.items():
    name = "?"
    if fn in funcs and funcs[fn]:
        idx = bisect.bisect_right([s[0] for s in funcs[fn]], ln) - 1
        if idx >= 0:
            name = funcs[fn][idx][1]
    agg[(fn, name)].update(c)
print("\nper function (inlined code is attributed to the function it was written in):")
nenv_steps = 4096.0
for (fn, name), c in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:45]:
    print(f"  {100 * c['# Samples'] / tot['# Samples']:5.1f}% samples  {c['Instructions Executed'] / nenv_steps:7.0f} inst/env  "
          f"lsb {c['stall_long_sb']:5.0f} noi {c['stall_no_inst']:5.0f} wait {c['stall_wait']:5.0f} ssb {c['stall_short_sb']:5.0f}  {fn}:{name}")
