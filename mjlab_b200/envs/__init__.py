from mjlab_b200.envs.tracking_env import TrackingEnvCfg, TrackingFlatEnv
from mjlab_b200.envs.velocity_env import VelocityEnvCfg, VelocityFlatEnv

__all__ = ["TrackingEnvCfg", "TrackingFlatEnv", "VelocityEnvCfg", "VelocityFlatEnv"]
