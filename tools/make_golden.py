"""Generate the committed golden vectors under tests/golden/ from the fp64 CPU oracle.

There are no golden vectors in the reference for this path and the reference cannot
run here, so these fixtures pin the *oracle itself* (regression guard) and give the GPU tests a
second, file-based target.  Re-run after an intentional change of the restated algorithm:
    python tools/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from mjlab_b200.asset_zoo import load_compiled  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from util import load_oracle, make_states, terrain_states  # noqa: E402

OUT = ROOT / "tests" / "golden"
FIELDS = ["qpos", "qvel", "qacc", "qacc_smooth", "qfrc_smooth", "qfrc_constraint", "xpos", "xquat",
          "subtree_com", "cvel", "actuator_force", "sensordata", "contact_dist", "contact_force"]


def main():
  OUT.mkdir(exist_ok=True)
  for name, seed in (("g1_flat", 101), ("go1_flat", 102), ("g1_tracking_flat", 103), ("go1_stairs_small", 104)):
    m = load_compiled(name)
    n = 8
    # box-terrain scenes: robots dropped around random sub-terrain origins (grid-static broadphase + box primitives)
    st = terrain_states(m, n, seed, 1.4) if "terrain_origins" in m.arrays else make_states(m, n, seed=seed)
    o = Oracle(m, nworld=n, maxcon=48)
    load_oracle(o, st)
    o.forward()
    fwd = {f"fwd_{f}": o.field(f).copy() for f in FIELDS[2:]}
    fwd["fwd_ncon"] = o.ncon.copy()
    fwd["fwd_nefc"] = o.nefc.copy()
    fwd["fwd_contact_geom"] = o.contact_geom.copy()
    load_oracle(o, st)
    for _ in range(3):
      o.step()
    stp = {f"step3_{f}": o.field(f).copy() for f in ("qpos", "qvel", "qacc_warmstart")}
    np.savez_compressed(OUT / f"{name}_seed{seed}.npz", seed=seed, n=n,
                        **{f"in_{k}": v for k, v in st.items()}, **fwd, **stp)
    print("wrote", name, "ncon", fwd["fwd_ncon"].ravel())


if __name__ == "__main__":
  main()
