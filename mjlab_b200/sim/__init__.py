from mjlab_b200.sim.sim import MujocoCfg, Simulation, SimulationCfg
from mjlab_b200.sim.sim_data import Bridge, TorchArray

__all__ = ["MujocoCfg", "Simulation", "SimulationCfg", "Bridge", "TorchArray"]
