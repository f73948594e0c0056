"""Per-source-line hot spots of one profiled kernel: joins the SASS page of an .ncu-rep (instructions executed,
stall samples per SASS instruction) with nvdisasm's line info for the same kernel in the built library.

    python scripts/ncu_hotspots.py gpurun_out/prof.ncu-rep open_spiel_coup_b200/libcoup_b200.so <mangled-substring> [top]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, lib, pattern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
cols = rows[hdr]
ci, cs, cx = cols.index("Instructions Executed"), cols.index("Warp Stall Sampling (All Samples)"), cols.index("Source")
insts = [(r[cx].strip(), int(r[ci] or 0), int(r[cs] or 0)) for r in rows[hdr + 1:] if len(r) > ci]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
dis = ""
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
m = re.search(r"\.section\s+\.text\.(\S*%s[^,\s]*)," % re.escape(pattern), dis)
name = m.group(1)
start = m.start()
body = dis[start + 10:]
nxt = re.search(r"\n\s*\.section\s", body)
body = body[: nxt.start()] if nxt else body
line = "?"
lines = []
for ln in body.split("\n"):
    mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if mm:
        line = "%s:%s" % (os.path.basename(mm.group(1)), mm.group(2))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
        lines.append(line)
print("kernel %s: %d SASS in report, %d in library" % (name, len(insts), len(lines)))
n = min(len(insts), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0])
for (txt, ex, st), l in zip(insts[:n], lines[:n]):
    a = agg[l]; a[0] += ex; a[1] += st; a[2] += 1
tot_ex = sum(a[0] for a in agg.values()); tot_st = sum(a[1] for a in agg.values())
print("total warp-instructions executed %d, stall samples %d" % (tot_ex, tot_st))
print("%-28s %12s %6s %10s %6s %5s" % ("source line", "inst exec", "%", "stalls", "%", "sass"))
for l, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-28s %12d %5.1f%% %10d %5.1f%% %5d" % (l, a[0], 100 * a[0] / tot_ex, a[1], 100 * a[1] / max(1, tot_st), a[2]))
