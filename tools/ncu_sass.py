"""Dump the SASS of one kernel from an ncu report with per-instruction executed counts, source line and stall samples.
usage: ncu_sass.py report.ncu-rep kernel_substr > out.txt"""
import csv, io, os, re, subprocess, sys, tempfile
rep, sub = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("PVT_LIB", os.path.join(root, "pvtrace_b200/csrc/libpvtrace_b200.so"))
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
ia, isrc, isamp, iinst, ithr = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
base = None
for r in rows[2:]:
    if len(r) <= ithr or not r[ia].startswith("0x"): continue
    addr = int(r[ia], 16)
    if base is None: base = addr
    k = line_of.get(addr - base)
    i, t = int(r[iinst] or 0), int(r[ithr] or 0)
    print(f"{addr - base:6x} {i / 1e6:8.2f}M {t / max(i, 1):5.1f} {int(r[isamp] or 0):6d}  {(k[0][4:-4] + ':' + str(k[1])) if k else '?':14s} {r[isrc]}")
