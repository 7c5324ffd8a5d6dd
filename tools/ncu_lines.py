"""Join an ncu source-page CSV (SASS rows) with nvdisasm line info: per source line samples / instructions / lane use.
usage: ncu_lines.py report.ncu-rep kernel_substr [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "pvtrace_b200/csrc/libpvtrace_b200.so")
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, stdout=subprocess.DEVNULL)
# (one cubin per translation unit of the library: take the one that holds the kernel)
txt = ""
for cub in sorted(f for f in os.listdir(d) if f.endswith(".cubin")):
    out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], stdout=subprocess.PIPE, text=True).stdout
    if sub in out:
        txt = out
        break
fn = None; cur = None; line_of = {}
for line in txt.splitlines():
    m = re.match(r'\s*\.text\.(\S+):', line)
    if m: fn = m.group(1); cur = None; continue
    if fn is None or sub not in fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+\S', line)
    if m: line_of[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ia, isamp, iinst, ithr = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
tot = [0, 0, 0]
for r in rows[2:]:
    if len(r) <= ithr or not r[ia].startswith("0x"): continue
    addr = int(r[ia], 16)
    if base is None: base = addr
    key = line_of.get(addr - base)
    a = agg[key]
    s, i, t = int(r[isamp] or 0), int(r[iinst] or 0), int(r[ithr] or 0)
    a[0] += s; a[1] += i; a[2] += t
    for ci, h in stall_cols:
        v = int(r[ci] or 0)
        if v: a[3][h[6:]] += v
    tot[0] += s; tot[1] += i; tot[2] += t
print(f"total samples {tot[0]}  warp-inst {tot[1]:.3e}  lanes/inst {tot[2] / max(tot[1], 1):.2f}")
src = {}
def text(k):
    if not k: return ""
    if k[0] not in src:
        p = os.path.join(root, "pvtrace_b200/csrc", k[0])
        src[k[0]] = open(p).read().splitlines() if os.path.exists(p) else []
    return src[k[0]][k[1] - 1].strip()[:70] if k[1] <= len(src[k[0]]) else ""
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ",".join(f"{n}:{v}" for n, v in a[3].most_common(3))
    print(f"{100 * a[0] / tot[0]:5.1f}% samp  {100 * a[1] / tot[1]:5.1f}% inst  lanes {a[2] / max(a[1], 1):5.1f}  {str(k):32s} {st:45s} {text(k)}")

# ---- per function: thread-instructions (the work) and warp-instructions (the issue slots) -----------------
import bisect
funcs = {}
for fname, lines in src.items():
    starts = []
    for no, text_ in enumerate(lines, 1):
        m = re.match(r'\s*(?:template <[^>]*>\s*)?(?:__device__|__global__|__host__).*?\b(\w+)\s*\(', text_)
        if m and not text_.strip().startswith("//"):
            starts.append((no, m.group(1)))
    funcs[fname] = starts
def func_of(k):
    if not k: return "?"
    st = funcs.get(k[0])
    if st is None:
        p = os.path.join(root, "pvtrace_b200/csrc", k[0])
        if os.path.exists(p):
            src[k[0]] = open(p).read().splitlines()
            starts = []
            for no, text_ in enumerate(src[k[0]], 1):
                m = re.match(r'\s*(?:template <[^>]*>\s*)?(?:__device__|__global__|__host__).*?\b(\w+)\s*\(', text_)
                if m and not text_.strip().startswith("//"): starts.append((no, m.group(1)))
            funcs[k[0]] = st = starts
        else:
            return k[0]
    i = bisect.bisect_right([s[0] for s in st], k[1]) - 1
    return f"{k[0]}:{st[i][1]}" if i >= 0 else k[0]
by = collections.defaultdict(lambda: [0, 0, 0])
for k, a in agg.items():
    f = func_of(k)
    by[f][0] += a[0]; by[f][1] += a[1]; by[f][2] += a[2]
print("\nper function:  samples%  warp-inst%  thread-inst%  lanes")
for f, a in sorted(by.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{100 * a[0] / tot[0]:6.1f} {100 * a[1] / tot[1]:6.1f} {100 * a[2] / tot[2]:6.1f} {a[2] / max(a[1], 1):6.1f}  {f}")
