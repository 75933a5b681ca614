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

__device__ __forceinline__ float b2e_wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per environment (a thread per environment left 4096 threads = 128 warps on 148 SMs walking ~400
// dependent loads each: 50 us of pure latency on the critical path of every env step).  Scalar terms are computed
// redundantly by every lane from values read before any lane writes; per-joint terms are spread over the lanes.
__global__ void b2_velenv_post_kernel(DevData dd, int nq, int nv, int nu, B2VelEnvArgs A) {
  const int lane = threadIdx.x & 31;
  const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
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
  float c0 = cmd[0], c1 = cmd[1], c2 = cmd[2];
  float ctl = A.cmd_time_left[w] - A.step_dt, tl = A.push_time_left[w] - A.step_dt;
  float htarget = A.heading_target[w];
  bool standing = A.is_standing[w] != 0;
  const float v0 = qvel[0], v1 = qvel[1], v2 = qvel[2], wz = qvel[5];
  b2e_rot_inv(q, down, gb);
  bool term = acosf(fminf(fmaxf(-gb[2], -1.f), 1.f)) > A.fall_angle;  // bad_orientation
  bool trunc = ep >= A.max_episode_length;                              // time_out
  // rewards (velocity_env_cfg.py:183-214)
  float lin[3] = {v0, v1, v2};
  b2e_rot_inv(q, lin, lb);
  float e0 = c0 - lb[0], e1 = c1 - lb[1], e2 = c2 - wz;
  float r_lin = expf(-(e0 * e0 + e1 * e1) / 0.25f), r_ang = expf(-(e2 * e2) / 0.25f);
  float pose = 0.f, lim = 0.f, rate = 0.f;
  for (int a = lane; a < nu; a += 32) {
    float jp = qpos[7 + a], dj = jp - A.default_joint_pos[a];
    pose += dj * dj;
    lim += fmaxf(A.soft_lo[a] - jp, 0.f) + fmaxf(jp - A.soft_hi[a], 0.f);
    float da = act[a] - last[a];
    rate += da * da;
  }
  pose = b2e_wsum(pose); lim = b2e_wsum(lim); rate = b2e_wsum(rate);
  __syncwarp();  // every lane holds the pre-reset state it needs; writes may start
  float r_pose = expf(-(pose / (float)nu) / 0.09f);
  const float reward = (r_lin + r_ang + r_pose - lim - 0.1f * rate) * A.step_dt;
  bool done = term || trunc;
  if (lane == 0) {
    A.reward[w] = reward;
    A.terminated[w] = term; A.truncated[w] = trunc; A.done[w] = done;
    A.log_row[3 * (size_t)w] = reward; A.log_row[3 * (size_t)w + 1] = term ? 1.f : 0.f;
    A.log_row[3 * (size_t)w + 2] = trunc ? 1.f : 0.f;
  }
  // masked reset (reset_root_state_uniform + reset_joints_by_scale), command resample
  if (done) {
    for (int i = lane; i < nq; i += 32) {
      float v = A.default_qpos[i];
      if (i == 0) v += (U[0] - 0.5f) + A.env_origins[3 * (size_t)w];
      if (i == 1) v += (U[1] - 0.5f) + A.env_origins[3 * (size_t)w + 1];
      if (i == 2) v += A.env_origins[3 * (size_t)w + 2];  // terrain spawn height (0 on the flat scene)
      if (i == 3 || i == 6) {
        float yaw = (U[2] * 2.f - 1.f) * 3.14f;
        v = i == 3 ? cosf(0.5f * yaw) : sinf(0.5f * yaw);
      }
      if (i >= 7) v = fminf(fmaxf(v, A.soft_lo[i - 7]), A.soft_hi[i - 7]);
      qpos[i] = v;
    }
    for (int a = lane; a < nu; a += 32) { ctrl[a] = A.default_joint_pos[a]; last[a] = 0.f; }
    for (int i = lane; i < nv; i += 32) qvel[i] = 0.f;
    ep = 0;
  } else {
    for (int a = lane; a < nu; a += 32) last[a] = act[a];
  }
  // command term (CommandTerm.compute + UniformVelocityCommand, velocity_command.py:64-110): resample on
  // reset and when the timer runs out, then heading control and standing envs every step
  if (done || ctl <= 0.f) {
    c0 = U[3] * 2.f - 1.f; c1 = U[4] - 0.5f; c2 = U[5] * 2.f - 1.f;
    htarget = (U[6] * 2.f - 1.f) * 3.14159265358979f;
    standing = U[7] <= 0.1f;
    ctl = 3.f + 5.f * U[8];
  }
  __syncwarp();  // the reset state is visible to every lane
  {
    // heading_w = atan2 of the body x axis in the world (entity/data.py:480-484), from the post-reset pose
    float qw = qpos[3], qx = qpos[4], qy = qpos[5], qz = qpos[6];
    float fx = 1.f - 2.f * (qy * qy + qz * qz), fy = 2.f * (qx * qy + qw * qz);
    float err = htarget - atan2f(fy, fx);
    err -= 6.28318530717959f * floorf((err + 3.14159265358979f) / 6.28318530717959f);  // wrap_to_pi
    c2 = fminf(fmaxf(0.5f * err, -1.f), 1.f);
    if (standing) { c0 = 0.f; c1 = 0.f; c2 = 0.f; }
  }
  // interval event: push_by_setting_velocity
  float pv0 = qvel[0], pv1 = qvel[1];
  const bool push = tl <= 0.f;
  if (push) {
    pv0 = (U[9] * 2.f - 1.f) * A.push_vel;
    pv1 = (U[10] * 2.f - 1.f) * A.push_vel;
    tl = U[11] * (A.push_hi - A.push_lo) + A.push_lo;
  }
  __syncwarp();
  if (lane == 0) {
    A.episode_length[w] = ep;
    cmd[0] = c0; cmd[1] = c1; cmd[2] = c2;
    A.heading_target[w] = htarget; A.is_standing[w] = standing; A.cmd_time_left[w] = ctl;
    A.push_time_left[w] = tl;
    if (push) { qvel[0] = pv0; qvel[1] = pv1; }
  }
  // observations from the (possibly reset / pushed) state
  // policy group with uniform noise (velocity_env_cfg.py:86-118), critic group without (:120-125)
  float* o = A.obs + (size_t)w * (9 + 3 * nu + 3);
  float* cr = A.critic + (size_t)w * (9 + 3 * nu + 3);
  const float* Z = U + 16;  // noise draws
  float q2[4] = {qpos[3], qpos[4], qpos[5], qpos[6]}, lin2[3] = {pv0, pv1, qvel[2]};
  b2e_rot_inv(q2, lin2, lb);
  b2e_rot_inv(q2, down, gb);
  if (lane < 3) {
    const int k = lane;
    const float av = qvel[3 + k];
    cr[k] = lb[k]; cr[3 + k] = av; cr[6 + k] = gb[k];
    o[k] = lb[k] + (Z[k] * 2.f - 1.f) * 0.1f;
    o[3 + k] = av + (Z[3 + k] * 2.f - 1.f) * 0.2f;
    o[6 + k] = gb[k] + (Z[6 + k] * 2.f - 1.f) * 0.05f;
    const float ck = k == 0 ? c0 : (k == 1 ? c1 : c2);
    o[9 + 3 * nu + k] = ck; cr[9 + 3 * nu + k] = ck;
  }
  for (int a = lane; a < nu; a += 32) {
    float jp = qpos[7 + a] - A.default_joint_pos[a], jv = qvel[6 + a], la = last[a];
    cr[9 + a] = jp; cr[9 + nu + a] = jv; cr[9 + 2 * nu + a] = la;
    o[9 + a] = jp + (Z[9 + a] * 2.f - 1.f) * 0.01f;
    o[9 + nu + a] = jv + (Z[9 + nu + a] * 2.f - 1.f) * 1.5f;
    o[9 + 2 * nu + a] = la;
  }
}
