"""Static code-size attribution of the step kernel: SASS instructions per source file / line range."""
import collections, glob, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "build/kernels/step_kernel.cu.o")], cwd=tmp, check=True, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", glob.glob(os.path.join(tmp, "*.cubin"))[0]], capture_output=True, text=True).stdout
cur = None
perfile = collections.Counter(); perline = collections.Counter(); sect = None; persect = collections.Counter()
for ln in dis.splitlines():
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*\.section\s+(\.text\.\S+?),", ln)
    if m: sect = m.group(1); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln) and cur:
        perfile[cur[0]] += 1; perline[cur] += 1; persect[sect] += 1
print("sections:", dict(persect))
tot = sum(perfile.values()); print("total", tot)
for f, c in perfile.most_common(): print(f"  {f:28s} {c:7d} {100*c/tot:5.1f}%")
# bucket lines per file in ranges of 25 lines
b = collections.Counter()
for (f, l), c in perline.items(): b[(f, l // 25 * 25)] += c
print("top 25-line buckets:")
for (f, l), c in b.most_common(int(sys.argv[1]) if len(sys.argv) > 1 else 40): print(f"  {f}:{l}-{l+24}  {c}")
