// b2_env.cuh — fused MDP glue of the velocity-tracking task around the physics step (SURVEY.md §8f items
// 1-3): the caller-side work of ManagerBasedRlEnv.step (reference src/mjlab/envs/manager_based_rl_env.py:106-147
// with the terms of tasks/velocity/velocity_env_cfg.py) as two kernels instead of ~90 tiny torch launches:
//   pre  : ctrl = default_joint_pos + action_scale * action                     (joint_actions.py:85-103)
//   post : episode counter, terminations (bad_orientation, time_out), rewards, masked reset
//          (events.py:43-124), interval push (events.py:127-143), command resample, observations.
// The torch implementation in mjlab_b200/envs/velocity_env.py is the reference; both consume the same
// per-step uniform numbers U[n][10] so they can be compared exactly (tests/test_boundary_gpu.py).
#pragma once
#include "../../include/b2sim.h"
#include "b2_types.h"

// B2VelEnvArgs is declared in include/b2sim.h (it is part of the C ABI).

__device__ __forceinline__ void b2e_rot_inv(const float* q, const float* v, float* r) {
  // r = R(q)^T v
  float t0 = 2.f * (q[2] * v[2] - q[3] * v[1]);
  float t1 = 2.f * (q[3] * v[0] - q[1] * v[2]);
  float t2 = 2.f * (q[1] * v[1] - q[2] * v[0]);
  r[0] = v[0] - q[0] * t0 + (q[2] * t2 - q[3] * t1);
  r[1] = v[1] - q[0] * t1 + (q[3] * t0 - q[1] * t2);
  r[2] = v[2] - q[0] * t2 + (q[1] * t1 - q[2] * t0);
}

__global__ void b2_velenv_pre_kernel(DevData dd, int nu, const float* __restrict__ action,
                                     const float* __restrict__ default_joint_pos,
                                     const float* __restrict__ action_scale) {
  long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (long long)dd.nworld * nu) return;
  int w = (int)(tid / nu), a = (int)(tid % nu);
  dd.ctrl.p[(size_t)w * dd.ctrl.stride + a] = default_joint_pos[a] + action_scale[a] * action[tid];
}

__global__ void b2_velenv_post_kernel(DevData dd, int nq, int nv, int nu, B2VelEnvArgs A) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= dd.nworld) return;
  float* qpos = dd.qpos.p + (size_t)w * dd.qpos.stride;
  float* qvel = dd.qvel.p + (size_t)w * dd.qvel.stride;
  float* ctrl = dd.ctrl.p + (size_t)w * dd.ctrl.stride;
  const float* act = A.action + (size_t)w * nu;
  const float* U = A.U + (size_t)w * 10;
  float* last = A.last_action + (size_t)w * nu;
  float* cmd = A.command + 3 * (size_t)w;
  int ep = A.episode_length[w] + 1;
  const float down[3] = {0.f, 0.f, -1.f};
  float q[4] = {qpos[3], qpos[4], qpos[5], qpos[6]}, gb[3], lb[3];
  b2e_rot_inv(q, down, gb);
  bool term = acosf(fminf(fmaxf(-gb[2], -1.f), 1.f)) > A.fall_angle;  // bad_orientation
  bool trunc = ep >= A.max_episode_length;                              // time_out
  // rewards (velocity_env_cfg.py:183-214)
  float lin[3] = {qvel[0], qvel[1], qvel[2]};
  b2e_rot_inv(q, lin, lb);
  float e0 = cmd[0] - lb[0], e1 = cmd[1] - lb[1], e2 = cmd[2] - qvel[5];
  float r_lin = expf(-(e0 * e0 + e1 * e1) / 0.25f), r_ang = expf(-(e2 * e2) / 0.25f);
  float pose = 0.f, lim = 0.f, rate = 0.f;
  for (int a = 0; a < nu; a++) {
    float jp = qpos[7 + a], dj = jp - A.default_joint_pos[a];
    pose += dj * dj;
    lim += fmaxf(A.soft_lo[a] - jp, 0.f) + fmaxf(jp - A.soft_hi[a], 0.f);
    float da = act[a] - last[a];
    rate += da * da;
  }
  float r_pose = expf(-(pose / (float)nu) / 0.09f);
  A.reward[w] = (r_lin + r_ang + r_pose - lim - 0.1f * rate) * A.step_dt;
  bool done = term || trunc;
  A.terminated[w] = term; A.truncated[w] = trunc; A.done[w] = done;
  // masked reset (reset_root_state_uniform + reset_joints_by_scale), command resample
  if (done) {
    for (int i = 0; i < nq; i++) qpos[i] = A.default_qpos[i];
    qpos[0] += (U[0] - 0.5f) + A.env_origins[3 * (size_t)w];
    qpos[1] += (U[1] - 0.5f) + A.env_origins[3 * (size_t)w + 1];
    qpos[2] += A.env_origins[3 * (size_t)w + 2];  // terrain spawn height (0 on the flat scene)
    float yaw = (U[2] * 2.f - 1.f) * 3.14f;
    qpos[3] = cosf(0.5f * yaw); qpos[4] = A.default_qpos[4]; qpos[5] = A.default_qpos[5]; qpos[6] = sinf(0.5f * yaw);
    for (int a = 0; a < nu; a++) {
      qpos[7 + a] = fminf(fmaxf(qpos[7 + a], A.soft_lo[a]), A.soft_hi[a]);
      ctrl[a] = A.default_joint_pos[a];
      last[a] = 0.f;
    }
    for (int i = 0; i < nv; i++) qvel[i] = 0.f;
    cmd[0] = U[3] * 2.f - 1.f; cmd[1] = U[4] - 0.5f; cmd[2] = U[5] * 2.f - 1.f;
    ep = 0;
  } else {
    for (int a = 0; a < nu; a++) last[a] = act[a];
  }
  A.episode_length[w] = ep;
  // interval event: push_by_setting_velocity
  float tl = A.push_time_left[w] - A.step_dt;
  if (tl <= 0.f) {
    qvel[0] = (U[6] * 2.f - 1.f) * A.push_vel;
    qvel[1] = (U[7] * 2.f - 1.f) * A.push_vel;
    tl = U[8] * (A.push_hi - A.push_lo) + A.push_lo;
  }
  A.push_time_left[w] = tl;
  // observations from the (possibly reset / pushed) state
  float* o = A.obs + (size_t)w * (9 + 3 * nu + 3);
  float q2[4] = {qpos[3], qpos[4], qpos[5], qpos[6]}, lin2[3] = {qvel[0], qvel[1], qvel[2]};
  b2e_rot_inv(q2, lin2, lb);
  b2e_rot_inv(q2, down, gb);
  o[0] = lb[0]; o[1] = lb[1]; o[2] = lb[2];
  o[3] = qvel[3]; o[4] = qvel[4]; o[5] = qvel[5];
  o[6] = gb[0]; o[7] = gb[1]; o[8] = gb[2];
  for (int a = 0; a < nu; a++) {
    o[9 + a] = qpos[7 + a] - A.default_joint_pos[a];
    o[9 + nu + a] = qvel[6 + a];
    o[9 + 2 * nu + a] = last[a];
  }
  o[9 + 3 * nu] = cmd[0]; o[10 + 3 * nu] = cmd[1]; o[11 + 3 * nu] = cmd[2];
}
