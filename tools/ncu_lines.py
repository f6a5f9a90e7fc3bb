"""Per-source-line executed-instruction counts: joins `ncu --page source` (SASS, per-instruction counters) with
nvdisasm line info of the cubin.  usage: ncu_lines.py rep.ncu-rep lib.so kernel_substring [topN]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, lib, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# address -> (file,line) for the chosen kernel
addr2line = {}; cur = None; inside = False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        inside = kname in ln
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(\S+)", ln)
    if m: addr2line[int(m.group(1), 16)] = cur
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
h = src[1]; ix = {n: i for i, n in enumerate(h)}
rows = [r for r in src[2:] if len(r) >= len(h)]
base = int(rows[0][ix["Address"]], 16)
per = collections.Counter(); stall = collections.Counter(); static = collections.Counter(); tot = 0
for r in rows:
    a = int(r[ix["Address"]], 16) - base
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
    key = addr2line.get(a)
    per[key] += n; stall[key] += s; static[key] += 1; tot += n
stot = sum(stall.values())
srcs = {}
def text(key):
    if not key: return ""
    f, l = key
    if f not in srcs:
        for root in ("car_trailer_mpc_b200/csrc",):
            pth = os.path.join(root, f)
            srcs[f] = open(pth).read().splitlines() if os.path.exists(pth) else []
    return srcs[f][l - 1].strip()[:90] if l - 1 < len(srcs[f]) else ""
print(f"total executed {tot}, stall samples {stot}, static SASS {len(rows)}")
print("--- by stall samples")
for key, n in stall.most_common(topn//2):
    print(f"{100*per[key]/tot:5.1f}% exec {100*n/max(stot,1):5.1f}% stall {static[key]:5d} sass  {key}  {text(key)}")
print("--- by executed instructions")
for key, n in per.most_common(topn):
    print(f"{100*n/tot:5.1f}% exec {100*stall[key]/max(stot,1):5.1f}% stall {static[key]:5d} sass  {key}  {text(key)}")
