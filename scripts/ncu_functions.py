"""Executed warp-instructions and stall samples of one profiled kernel, aggregated by the (inlined) source function:
joins the SASS page of an .ncu-rep with nvdisasm's line info of the library that was profiled.

    python scripts/ncu_functions.py <prof.ncu-rep> <libcoup_b200.so used> <kernel substring> <warps x steps of the launch>
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, lib, pattern = sys.argv[1:4]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
cols = rows[hdr]
ci, cs = cols.index("Instructions Executed"), cols.index("Warp Stall Sampling (All Samples)")
insts = [(int(r[ci] or 0), int(r[cs] or 0)) for r in rows[hdr + 1:] if len(r) > ci]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
dis = "".join(subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
              for f in os.listdir(tmp) if f.endswith(".cubin"))
m = re.search(r"\.section\s+\.text\.(\S*%s[^,\s]*)," % re.escape(pattern), dis)
body = dis[m.start() + 10:]
nxt = re.search(r"\n\s*\.section\s", body)
body = body[: nxt.start()] if nxt else body
line, lines = ("?", 0), []
for ln in body.split("\n"):
    mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if mm:
        line = (os.path.basename(mm.group(1)), int(mm.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
        lines.append(line)


def functions(path):
    out = []
    for i, l in enumerate(open(path), 1):
        mm = re.match(r"\s*(?:COUP_FN|__device__ __forceinline__|__global__|template).*?\b([a-zA-Z_0-9]+)\(", l)
        if mm and not l.strip().startswith("//"):
            out.append((i, mm.group(1)))
    return out


src = {f: functions(os.path.join(ROOT, "open_spiel_coup_b200", "csrc", f)) for f in sorted(os.listdir(os.path.join(ROOT, "open_spiel_coup_b200", "csrc"))) if f.endswith(".cuh")}


def function_of(f, ln):
    best = f
    for start, name in src.get(f, []):
        if start <= ln:
            best = name
        else:
            break
    return best


agg = collections.defaultdict(lambda: [0, 0])
for (ex, st), (f, l) in zip(insts, lines):
    a = agg[function_of(f, l)]
    a[0] += ex
    a[1] += st
tot, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print("kernel %s: %d warp-instructions, %.1f per unit, %d stall samples" % (m.group(1), tot, tot / units, ts))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print("%-28s %8.1f per unit %5.1f%%   stalls %5.1f%%" % (k, a[0] / units, 100 * a[0] / tot, 100 * a[1] / max(ts, 1)))
