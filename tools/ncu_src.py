"""Summarise an ncu --set full report per CUDA source line (SASS rows joined with nvdisasm -g line
info by instruction index): share of issued warp-instructions, average active threads, stall samples.
usage: ncu_src.py report.ncu-rep kernel_name launch_skip [top] [lib.so]"""
import csv, subprocess, sys, io, os, re, tempfile, collections
rep, kname, skip = sys.argv[1], sys.argv[2], sys.argv[3]
# kname may be "ncu_regex:disasm_substring" (e.g. "k_trace:k_traceILb1ELb0ELb0")
dname = kname.split(":")[1] if ":" in kname else kname
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[5] if len(sys.argv) > 5 else os.path.join(ROOT, "raytracer-odin_b200/csrc/libodinrt_b200.so")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kname.split(chr(58))[0]}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
seen, uniq = set(), []  # ncu may print the SASS table more than once: keep the first row of every address
for r in data:
    if r[0] in seen or not r[0].startswith("0x"):
        continue
    seen.add(r[0]); uniq.append(r)
data = uniq
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
# line info
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
cubs = [os.path.join(d, x) for x in os.listdir(d) if x.endswith(".cubin")]
cub = next((c for c in cubs if os.path.basename(c).startswith("odinrt")), max(cubs, key=os.path.getsize))
dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
lines, cur, infn = [], None, False
for l in dis.split("\n"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: infn = dname in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l): lines.append(cur)
print(f"sass rows {len(data)}  disasm instrs {len(lines)}")
src = {}
def srcline(fl):
    if fl is None: return "?"
    fn, ln = fl
    if fn not in src:
        p = os.path.join(ROOT, "raytracer-odin_b200/csrc", fn)
        src[fn] = open(p).read().split("\n") if os.path.exists(p) else []
    return src[fn][ln - 1].strip() if 0 < ln <= len(src[fn]) else ""
ti = sum(f(r, "Instructions Executed") for r in data); tt = sum(f(r, "Thread Instructions Executed") for r in data)
ts = sum(f(r, "# Samples") for r in data)
print(f"kernel {kname} skip {skip}: warp-inst {ti:.3e}  avg threads/inst {tt / ti:.2f}  samples {ts:.0f}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = sorted(((sum(f(r, s) for r in data), s) for s in stalls), reverse=True)[:7]
print("stalls:", ", ".join(f"{s[6:]} {v / ts * 100:.1f}%" for v, s in agg))
per = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
for i, r in enumerate(data):
    k = lines[i] if i < len(lines) else None
    a = per[k]; a[0] += f(r, "Instructions Executed"); a[1] += f(r, "Thread Instructions Executed"); a[2] += f(r, "# Samples"); a[3] += 1
for k, a in sorted(per.items(), key=lambda x: -x[1][0])[:top]:
    if a[0] == 0: break
    print(f"{str(k[1]) if k else '?':>5s} inst {a[0] / ti * 100:5.2f}% thr {a[1] / a[0]:5.1f} samp {a[2] / ts * 100:5.2f}% n={a[3]:3d} | {srcline(k)[:100]}")
