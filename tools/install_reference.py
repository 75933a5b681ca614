"""Install the UNMODIFIED reference package (pure Python) into baseline/_ref so that its own modules
(entity/data.py, envs/mdp/events.py, sim/sim_data.py ...) can be imported by the drop-in tests on the
GPU box, where /root/reference does not exist.  baseline/_ref is git-ignored (no reference source enters
the history) but not gpurun-ignored (it travels with the snapshot, like the built .so files).

Recipe: the reference's build backend `uv_build` is not in the image, so the install is made
from a copy under /tmp whose `[build-system]` table is switched to setuptools (src layout, *.py + *.xml
package data; meshes are left out: 29 MB of STL files nothing here reads).  `--no-deps`: mujoco, mujoco_warp
and warp are absent from the image and the wheelhouse; the tests provide stand-ins for exactly those three.
No source file of the package is edited.
"""
import re
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
TARGET = ROOT / "baseline" / "_ref"


def install(force: bool = False) -> bool:
  if not REF.exists():
    return TARGET.joinpath("mjlab").exists()
  if TARGET.joinpath("mjlab", "entity", "data.py").exists() and not force:
    return True
  tmp = Path("/tmp/b2_refcopy")
  shutil.rmtree(tmp, ignore_errors=True)
  shutil.copytree(REF, tmp, ignore=shutil.ignore_patterns(".git", "*.stl", "*.STL", "*.obj", "*.png", "docs"))
  pp = tmp / "pyproject.toml"
  txt = pp.read_text()
  txt = re.sub(r"\[build-system\].*?(?=\n\[)", '[build-system]\nrequires = ["setuptools"]\nbuild-backend = "setuptools.build_meta"\n',
               txt, count=1, flags=re.S)
  txt = re.sub(r"\n\[tool\.uv[^\]]*\].*?(?=\n\[|\Z)", "\n", txt, flags=re.S)
  txt = re.sub(r"license-files = .*\n", "", txt)
  txt = re.sub(r'\s*"License :: [^"]*",\n', "\n", txt)  # setuptools >= 77 refuses licence classifiers next to `license`
  txt += ('\n[tool.setuptools.packages.find]\nwhere = ["src"]\n'
          '\n[tool.setuptools.package-data]\n"*" = ["*.xml", "*.typed", "*.yaml", "*.json"]\n')
  pp.write_text(txt)
  shutil.rmtree(TARGET, ignore_errors=True)
  TARGET.parent.mkdir(exist_ok=True)
  r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                      "--find-links", "/opt/wheelhouse", "--target", str(TARGET), str(tmp)],
                     capture_output=True, text=True)
  if r.returncode != 0:
    print(r.stdout[-1500:], r.stderr[-1500:])
    return False
  return TARGET.joinpath("mjlab", "entity", "data.py").exists()


if __name__ == "__main__":
  ok = install(force="--force" in sys.argv)
  print("baseline/_ref:", "installed" if ok else "NOT installed")
  sys.exit(0 if ok else 1)
