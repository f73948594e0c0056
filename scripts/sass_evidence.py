"""Writes profiles/sass_evidence.txt: per kernel of the built library, the SASS mnemonics that back the claims made in
DESIGN.md / profiles/README.md -- bulk (TMA) stores (UBLKCP), warp votes / shuffles / reductions, and the ABSENCE of any
tensor-core instruction (HMMA / IMMA / UTCMMA ...: there is no contraction on this path).

    python scripts/sass_evidence.py            # needs only cuobjdump; no GPU
"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "open_spiel_coup_b200", "libcoup_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass_evidence.txt")
WATCH = ("UBLKCP", "UTMA", "VOTE", "SHFL", "REDUX", "MATCH", "ATOM", "RED", "BAR", "LDL", "STL", "BRA", "BSSY")
TENSOR = re.compile(r"\b(HMMA|IMMA|DMMA|BMMA|QMMA|OMMA|UTCHMMA|UTCIMMA|UTCQMMA|UTCMMA|HGMMA|IGMMA|WGMMA)\b")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
parts = re.split(r"\n\s*Function : ", sass)
lines = ["SASS evidence for %s (cuobjdump -sass, arch %s)" % (os.path.relpath(LIB, ROOT), sorted(set(re.findall(r"sm_\d+a?", sass)))),
         "columns: instructions | " + " ".join(WATCH) + " | tensor-core instructions", ""]
tensor_total = 0
for part, name in zip(parts[1:], names):
    ops = collections.Counter()
    n = 0
    for l in part.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            n += 1
            ops[m.group(1)] += 1
    tensor = len(TENSOR.findall(part))
    tensor_total += tensor
    short = re.sub(r"coup::", "", name)
    short = re.sub(r"\(.*", "", short)
    lines.append("%-58s %5d | %s | %d" % (short[:58], n, " ".join("%s=%d" % (w, sum(v for k, v in ops.items() if k.startswith(w))) for w in WATCH), tensor))
lines += ["", "kernels: %d, tensor-core instructions in the whole library: %d" % (len(parts) - 1, tensor_total),
          "kernels with bulk (TMA) stores: " + ", ".join(sorted({re.sub(r"\(.*", "", re.sub(r"coup::", "", n)) for p, n in zip(parts[1:], names) if "UBLKCP" in p}))]
open(OUT, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-3:]))
