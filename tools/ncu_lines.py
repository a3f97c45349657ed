"""Warp-stall samples per CUDA source line: joins the SASS view of an .ncu-rep with the line table of the built
library (nvdisasm --print-line-info), matched by instruction order inside the kernel.
usage: python tools/ncu_lines.py gpurun_out/prof.ncu-rep <kernel-name-regex> [top N]"""
import csv
import re
import subprocess
import sys
import tempfile
import os

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "thewalrus_b200", "libwalrus_b200.so")

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
sass = [(r[iSrc].strip(), int(r[iS]), int(r[iI])) for r in rows[2:] if len(r) > iI]

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, check=True, capture_output=True)
    lines = None
    for f in sorted(os.listdir(td)):
        out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, f)], capture_output=True, text=True).stdout
        # split into functions
        cur, name, funcs = [], None, {}
        for ln in out.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                if name:
                    funcs[name] = cur
                name, cur = m.group(1), []
                continue
            cur.append(ln)
        if name:
            funcs[name] = cur
        for fn, body in funcs.items():
            if re.search(pat, fn):
                seq, line = [], None
                for ln in body:
                    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
                    if m:
                        line = (os.path.basename(m.group(1)), int(m.group(2)))
                        continue
                    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
                    if m:
                        seq.append((m.group(1).strip(), line))
                if abs(len(seq) - len(sass)) <= 2:
                    lines = seq
                    print(f"# matched {fn} ({len(seq)} instructions; ncu has {len(sass)})")
                    break
        if lines:
            break
if not lines:
    raise SystemExit("no function with a matching instruction count (is the library the build that was profiled?)")
agg, tot = {}, sum(s for _, s, _ in sass)
toti = sum(i for _, _, i in sass)
for (txt, smp, ins), (_, line) in zip(sass, lines):
    a = agg.setdefault(line, [0, 0])
    a[0] += smp
    a[1] += ins
src_cache = {}
def src(line):
    if not line:
        return "?"
    path = None
    for d in ("thewalrus_b200/csrc", "include"):
        p = os.path.join(ROOT, d, line[0])
        if os.path.exists(p):
            path = p
    if not path:
        return line[0]
    if path not in src_cache:
        src_cache[path] = open(path).read().splitlines()
    return src_cache[path][line[1] - 1].strip()[:100]
print(f"# total samples {tot}, warp instructions {toti}")
for line, (smp, ins) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * smp / tot:5.1f}% samples {100 * ins / toti:5.1f}% instr  {line[0] if line else '?'}:{line[1] if line else 0:<4d} {src(line)}")
