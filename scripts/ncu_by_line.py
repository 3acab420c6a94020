"""Join an ncu SASS source page (csv) with nvdisasm line info -> stall samples per CUDA source line.

usage: ncu_by_line.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> [top_n]
"""
import csv, os, re, subprocess, sys, tempfile, collections

rep, lib, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
# the library holds one cubin per translation unit: take the one that defines the kernel
start = None
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    start = next((i for i, l in enumerate(sass) if l.startswith(".text.") and kname in l), None)
    if start is not None:
        break
if start is None:
    sys.exit(f"kernel {kname} not found in {lib}")
lines = []  # source line per instruction in order
cur = None
for l in sass[start + 1:]:
    if l.startswith("\t.section") or l.startswith("//-----"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l):
        lines.append(cur)
src_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src_csv.splitlines()))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
idx = {n: hdr.index(n) for n in ["# Samples", "Source", "Instructions Executed", "stall_barrier", "stall_short_sb",
                                "stall_wait", "stall_long_sb", "stall_selected", "stall_math", "stall_branch_resolving",
                                "L1 Wavefronts Shared"]}
inst = rows[hi + 1:]
print(f"# sass instructions: ncu={len(inst)} nvdisasm={len(lines)}")
agg = collections.defaultdict(lambda: collections.Counter())
for i, r in enumerate(inst):
    key = lines[i] if i < len(lines) else None
    for n, j in idx.items():
        if n == "Source":
            continue
        try:
            agg[key][n] += int(r[j])
        except ValueError:
            pass
tot = sum(c["# Samples"] for c in agg.values())
srcfile = {}
def srcline(key):
    if key is None: return "?"
    f, ln = key
    if f not in srcfile:
        for root in ("sushie_b200/csrc", "include", "."):
            p = os.path.join(root, f)
            if os.path.exists(p):
                srcfile[f] = open(p).read().splitlines(); break
        else:
            srcfile[f] = []
    s = srcfile[f]
    return s[ln - 1].strip()[:90] if 0 < ln <= len(s) else ""
print(f"total samples {tot}")
print("samples   pct   line  barrier short_sb  wait long_sb  selected     inst_exec  smem_wavefronts | source")
for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:topn]:
    ln = f"{key[0]}:{key[1]}" if key else "?"
    print(f"{c['# Samples']:8d} {100*c['# Samples']/tot:5.1f} {ln:>22} {c['stall_barrier']:7d} {c['stall_short_sb']:7d} {c['stall_wait']:6d} {c['stall_long_sb']:6d} {c['stall_selected']:8d} {c['Instructions Executed']:12d} {c['L1 Wavefronts Shared']:12d} | {srcline(key)}")
