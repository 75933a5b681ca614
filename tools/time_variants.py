"""Time step_n(4) for several builds of libb2sim (B2SIM_LIB override), one process per variant."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
vdir = os.path.join(root, "mjlab_b200", "csrc", "variants")
for v in [None] + sorted(f for f in os.listdir(vdir) if f.startswith("v_")):
  env = dict(os.environ)
  if v: env["B2SIM_LIB"] = os.path.join(vdir, v)
  out = subprocess.run([sys.executable, os.path.join(root, "tools", "profile_step.py"), "4096", "time"],
                       capture_output=True, text=True, env=env).stdout.splitlines()
  print(v or "default", "|", " | ".join(l.strip() for l in out[:2]))
