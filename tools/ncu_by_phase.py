"""Aggregate tools/ncu_by_line.py output by kernel phase / helper function.

  python tools/ncu_by_phase.py <report.ncu-rep> [kernel-substring]
Ranges come from the *current* b2_kernel.cuh (out-of-line helpers by their definition lines, kernel phases by
their PHASE_MARK lines), so the library and source must be the ones the report was captured with."""
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "mjlab_b200", "csrc", "b2_kernel.cuh")).read().splitlines()
marks = []  # (line, label)
for i, ln in enumerate(src, 1):
  m = re.match(r"__device__ __noinline__ \w[\w ]*?(\w+)\(", ln)
  if m:
    marks.append((i, "fn " + m.group(1)))
  m = re.match(r"__device__ __forceinline__ \w[\w ]*?(\w+)\(", ln)
  if m:
    marks.append((i, "inline helpers"))
  m = re.match(r"\s*PHASE_MARK\((\d+)\);", ln)
  if m:
    marks.append((i, "after mark " + m.group(1)))
  if ln.startswith("template <bool STEP>"):
    marks.append((i, "kernel prologue / TMA load"))
names = {"after mark 0": "1 kinematics", "after mark 1": "1b geom/site poses", "after mark 2": "2 com/cinert/cdof",
         "after mark 3": "3 CRB", "after mark 4": "4 velocities/RNE/actuation", "after mark 5": "5 collision",
         "after mark 6": "6 limits/groups/aref", "after mark 7": "7 M factor + qacc_smooth",
         "after mark 8": "8 solver: init + update (J^T f, cost)", "after mark 12": "8 solver: H assembly",
         "after mark 13": "8 solver: factor/solve calls", "after mark 14": "8 solver: line-search setup",
         "after mark 15": "8 solver: line search", "after mark 16": "8 solver: move + bookkeeping",
         "after mark 9": "9 contact forces / sensors", "after mark 10": "10 integrate", "after mark 11": "11 stores"}
marks.sort()
out = subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_by_line.py"), sys.argv[1],
                      sys.argv[2] if len(sys.argv) > 2 else "b2_step_kernelILi1", "100000"],
                     capture_output=True, text=True).stdout
tot = {}
for line in out.splitlines():
  m = re.match(r"\('([^']+)', (\d+)\)\s+samp\s+([\d.]+)%\s+inst\s+([\d.]+)%", line)
  if not m:
    if line.startswith("total"):
      print(line)
    continue
  f, l, sp, ip = m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4))
  if f != "b2_kernel.cuh":
    key = "cuda intrinsics headers (shfl, ...)"
  else:
    key = "?"
    for ml, lab in marks:
      if ml <= l:
        key = names.get(lab, lab)
  t = tot.setdefault(key, [0.0, 0.0])
  t[0] += sp
  t[1] += ip
print(f"{'phase / function':44s} samples%  inst%")
for k, v in sorted(tot.items(), key=lambda x: -x[1][0]):
  print(f"{k:44s} {v[0]:7.1f} {v[1]:7.1f}")
