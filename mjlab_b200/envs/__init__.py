from mjlab_b200.envs.velocity_env import VelocityEnvCfg, VelocityFlatEnv

__all__ = ["VelocityEnvCfg", "VelocityFlatEnv"]
