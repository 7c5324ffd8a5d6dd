"""Attribute the SASS instructions of one kernel to source lines (needs -lineinfo).  usage: sass_lines.py <substr> [top]"""
import collections, os, re, subprocess, sys, tempfile
sub = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pvtrace_b200/csrc/libpvtrace_b200.so")
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], stdout=subprocess.PIPE, text=True).stdout
fn = None; cur = None; cnt = collections.Counter(); total = collections.Counter()
for line in txt.splitlines():
    m = re.match(r'\s*\.text\.(\S+):', line)
    if m: fn = m.group(1); cur = None; continue
    if fn is None: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', line):
        total[fn] += 1
        if sub in fn: cnt[cur] += 1
for f, v in total.most_common(12): print(v, f[:110])
byfile = collections.Counter()
for k, v in cnt.items(): byfile[k[0] if k else None] += v
print(byfile.most_common())
src = {}
for k, v in cnt.most_common(top):
    if k and k[0] not in src:
        for root in ("pvtrace_b200/csrc",):
            pth = os.path.join(os.path.dirname(lib), k[0])
            src[k[0]] = open(pth).read().splitlines() if os.path.exists(pth) else []
    text = src.get(k[0], [])[k[1] - 1].strip()[:90] if k and src.get(k[0]) and k[1] <= len(src[k[0]]) else ""
    print(v, k, text)
