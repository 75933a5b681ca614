"""Compile the reference's G1 / Go1 MJCF into the flat model blobs shipped in
``mjlab_b200/asset_zoo/compiled`` (run in the authoring container; needs /root/reference).

Task sim options: dt 0.005, Newton 10 iterations, 20 line-search iterations, implicitfast,
pyramidal (reference ``tasks/velocity/velocity_env_cfg.py:248-256``).
"""

import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from mjlab_b200.asset_zoo import COMPILED_DIR, g1, go1, reference_xml  # noqa: E402
from mjlab_b200.asset_zoo.scene import compile_scene  # noqa: E402
from mjlab_b200.sim.sim import MujocoCfg  # noqa: E402

TASK = MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20)


def main():
  COMPILED_DIR.mkdir(exist_ok=True)
  g1_xml, go1_xml = reference_xml("g1"), reference_xml("go1")
  out = {
    "g1_flat": compile_scene(g1.robot_cfg(g1_xml, g1.velocity_sensors()), TASK),
    "g1_tracking_flat": compile_scene(g1.robot_cfg(g1_xml, g1.tracking_sensors()), TASK),
    "go1_flat": compile_scene(go1.robot_cfg(go1_xml, go1.velocity_sensors()), TASK),
  }
  for name, m in out.items():
    m.save(COMPILED_DIR / f"{name}.npz")
    print(
      f"{name}: nq={int(m.nq)} nv={int(m.nv)} nu={int(m.nu)} nbody={int(m.nbody)} "
      f"njnt={int(m.njnt)} ngeom={int(m.ngeom)} nsite={int(m.nsite)} npair={int(m.npair)} "
      f"nsensordata={int(m.nsensordata)} meaninertia={float(m.stat_meaninertia):.5f}"
    )


if __name__ == "__main__":
  main()
