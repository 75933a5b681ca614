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
  const int NUU = B2_VELENV_NU(nu);
  const float* U = A.U + (size_t)w * NUU;
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
  A.log_row[3 * (size_t)w] = A.reward[w]; A.log_row[3 * (size_t)w + 1] = term ? 1.f : 0.f;
  A.log_row[3 * (size_t)w + 2] = trunc ? 1.f : 0.f;
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
    ep = 0;
  } else {
    for (int a = 0; a < nu; a++) last[a] = act[a];
  }
  A.episode_length[w] = ep;
  // command term (CommandTerm.compute + UniformVelocityCommand, velocity_command.py:64-110): resample on
  // reset and when the timer runs out, then heading control and standing envs every step
  float ctl = A.cmd_time_left[w] - A.step_dt;
  if (done || ctl <= 0.f) {
    cmd[0] = U[3] * 2.f - 1.f; cmd[1] = U[4] - 0.5f; cmd[2] = U[5] * 2.f - 1.f;
    A.heading_target[w] = (U[6] * 2.f - 1.f) * 3.14159265358979f;
    A.is_standing[w] = U[7] <= 0.1f;
    ctl = 3.f + 5.f * U[8];
  }
  A.cmd_time_left[w] = ctl;
  {
    // heading_w = atan2 of the body x axis in the world (entity/data.py:480-484), from the post-reset pose
    float qw = qpos[3], qx = qpos[4], qy = qpos[5], qz = qpos[6];
    float fx = 1.f - 2.f * (qy * qy + qz * qz), fy = 2.f * (qx * qy + qw * qz);
    float err = A.heading_target[w] - atan2f(fy, fx);
    err -= 6.28318530717959f * floorf((err + 3.14159265358979f) / 6.28318530717959f);  // wrap_to_pi
    cmd[2] = fminf(fmaxf(0.5f * err, -1.f), 1.f);
    if (A.is_standing[w]) { cmd[0] = 0.f; cmd[1] = 0.f; cmd[2] = 0.f; }
  }
  // interval event: push_by_setting_velocity
  float tl = A.push_time_left[w] - A.step_dt;
  if (tl <= 0.f) {
    qvel[0] = (U[9] * 2.f - 1.f) * A.push_vel;
    qvel[1] = (U[10] * 2.f - 1.f) * A.push_vel;
    tl = U[11] * (A.push_hi - A.push_lo) + A.push_lo;
  }
  A.push_time_left[w] = tl;
  // observations from the (possibly reset / pushed) state
  // policy group with uniform noise (velocity_env_cfg.py:86-118), critic group without (:120-125)
  float* o = A.obs + (size_t)w * (9 + 3 * nu + 3);
  float* cr = A.critic + (size_t)w * (9 + 3 * nu + 3);
  const float* Z = U + 16;  // noise draws
  float q2[4] = {qpos[3], qpos[4], qpos[5], qpos[6]}, lin2[3] = {qvel[0], qvel[1], qvel[2]};
  b2e_rot_inv(q2, lin2, lb);
  b2e_rot_inv(q2, down, gb);
  for (int k = 0; k < 3; k++) {
    cr[k] = lb[k]; cr[3 + k] = qvel[3 + k]; cr[6 + k] = gb[k];
    o[k] = lb[k] + (Z[k] * 2.f - 1.f) * 0.1f;
    o[3 + k] = qvel[3 + k] + (Z[3 + k] * 2.f - 1.f) * 0.2f;
    o[6 + k] = gb[k] + (Z[6 + k] * 2.f - 1.f) * 0.05f;
  }
  for (int a = 0; a < nu; a++) {
    float jp = qpos[7 + a] - A.default_joint_pos[a], jv = qvel[6 + a];
    cr[9 + a] = jp; cr[9 + nu + a] = jv; cr[9 + 2 * nu + a] = last[a];
    o[9 + a] = jp + (Z[9 + a] * 2.f - 1.f) * 0.01f;
    o[9 + nu + a] = jv + (Z[9 + nu + a] * 2.f - 1.f) * 1.5f;
    o[9 + 2 * nu + a] = last[a];
  }
  for (int k = 0; k < 3; k++) { o[9 + 3 * nu + k] = cmd[k]; cr[9 + 3 * nu + k] = cmd[k]; }
}
