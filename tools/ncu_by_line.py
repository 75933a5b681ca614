"""Aggregate an ncu SASS source-page CSV by CUDA source line (uses nvdisasm -g line markers).

usage: python tools/ncu_by_line.py <report.ncu-rep> <kernel-mangled-substring> [top]
"""
import csv, re, subprocess, sys, tempfile, os, collections
rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
so = os.environ.get("B2SIM_LIB") or os.path.join(os.path.dirname(__file__), "..", "mjlab_b200", "csrc", "libb2sim.so")  # must be the build the report was captured with
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# offset -> line
off2line = {}
cur = None; infunc = False
for ln in dis.splitlines():
  if ln.startswith("\t.section") or ln.startswith("//-----"):
    infunc = kname in ln if ".text." in ln else infunc
  m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
  if m:
    cur = (os.path.basename(m.group(1)), int(m.group(2)))
  m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", ln)
  if m and infunc:
    off2line[int(m.group(1), 16)] = cur
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
stall_agg = collections.defaultdict(lambda: collections.Counter())
for r in rows[2:]:
  if len(r) < len(hdr): continue
  a = int(r[ia], 16)
  if base is None: base = a
  key = off2line.get(a - base)
  agg[key][0] += int(r[ii] or 0); agg[key][1] += int(r[isamp] or 0); agg[key][2] += 1
  for c in stall_cols:
    v = int(r[c] or 0)
    if v: stall_agg[key][hdr[c]] += v
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f"total inst {tot_i}  samples {tot_s}  sass instrs {sum(v[2] for v in agg.values())}")
src = {}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
  if k and k[0] not in src:
    try: src[k[0]] = open(os.path.join(os.path.dirname(__file__), "..", "mjlab_b200", "csrc", k[0])).read().splitlines()
    except Exception: src[k[0]] = []
  text = src.get(k[0], [""])[k[1] - 1].strip()[:90] if k and src.get(k[0]) and k[1] - 1 < len(src[k[0]]) else ""
  st = ",".join(f"{n[6:]}:{c}" for n, c in stall_agg[k].most_common(3))
  print(f"{str(k):28s} samp {100*v[1]/max(tot_s,1):5.1f}%  inst {100*v[0]/max(tot_i,1):5.1f}%  sass {v[2]:5d}  {st:40s} | {text}")
if os.environ.get("NCU_BY_FILE"):  # exact totals per source file and per function range of b2_convex.h
  byf = collections.defaultdict(lambda: [0, 0])
  for k, v in agg.items():
    f = k[0] if k else None
    byf[f][0] += v[0]; byf[f][1] += v[1]
  for f, v in sorted(byf.items(), key=lambda kv: -kv[1][1]):
    print(f"FILE {str(f):32s} samp {100*v[1]/max(tot_s,1):6.2f}%  inst {100*v[0]/max(tot_i,1):6.2f}%")
  cs = open(os.path.join(os.path.dirname(__file__), "..", "mjlab_b200", "csrc", "b2_convex.h")).read().splitlines()
  marks = [(i, m.group(2)) for i, ln in enumerate(cs, 1) for m in [re.match(r"B2C_(FN|INL) \w[\w ]*?(\w+)\(", ln)] if m]
  fn = collections.defaultdict(lambda: [0, 0])
  for k, v in agg.items():
    if not k or k[0] != "b2_convex.h": continue
    lab = "?"
    for ml, name in marks:
      if ml <= k[1]: lab = name
    fn[lab][0] += v[0]; fn[lab][1] += v[1]
  for f, v in sorted(fn.items(), key=lambda kv: -kv[1][1]):
    print(f"CONVEX {f:28s} samp {100*v[1]/max(tot_s,1):6.2f}%  inst {100*v[0]/max(tot_i,1):6.2f}%")
