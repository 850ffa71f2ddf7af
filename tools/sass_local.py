"""Static count of local-memory instructions (LDL / STL) per CUDA source line of the step kernel object.
Usage: python tools/sass_local.py [object] [top N] [file filter]"""
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.join(ROOT, "build/kernels/step_kernel.cu.o")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
filt = sys.argv[3] if len(sys.argv) > 3 else ""
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, check=True, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cur = None
cnt = collections.Counter()
total = 0
for ln in dis.splitlines():
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.search(r"\b(LDL|STL)\b", ln):
        total += 1
        if filt in cur[0]:
            cnt[cur] += 1
print("total LDL/STL", total)
for (fn, l), v in cnt.most_common(top):
    p = glob.glob(os.path.join(ROOT, "mujoco_ros_pkgs_b200/csrc/*", fn))
    text = open(p[0]).read().splitlines()[l - 1].strip()[:100] if p else ""
    print(f"{v:5d} {fn}:{l} {text}")
