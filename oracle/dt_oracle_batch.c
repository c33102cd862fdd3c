/* dt_oracle_batch.c — CPU ORACLE (test infrastructure): Simulator.step() = physics + done/reward +
 * render for a batch of independent envs, one env per OpenMP task.  This is what bench.py times as
 * the CPU baseline ("port") and what `bench.py --impl reference` runs: the way the reference would
 * be scaled on host cores with one env per worker. */
#include <string.h>

#include "dt_oracle.h"

typedef struct {
  orc_dyn_state s;
  int32_t step_count;
  double px, pz;          /* last simulator-frame position */
  double px0, pz0, ang0;  /* spawn pose: a finished episode restarts here (no host reset in the sample) */
} orc_env;

void orc_env_init(const orc_map* m, orc_env* e, double px, double pz, double ang) {
  memset(e, 0, sizeof *e);
  orc_cartesian_from_weird(m, px, pz, ang, &e->s);
  e->px = e->px0 = px; e->pz = e->pz0 = pz; e->ang0 = ang;
}

void orc_full_step_batch(const orc_map* m, const orr_scene* sc, const orc_dyn_params* dp, int n, orc_env* envs,
                         const float* actions, int action_mode, double wheel_dist, const double env5[5],
                         int frame_skip, double dt, int max_steps, double robot_speed, const orr_episode* eps, int W,
                         int H, int render, uint8_t* obs, float* reward, uint8_t* done, int threads) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int k = 0; k < n; k++) {
    orc_env* e = &envs[k];
    orc_step_out o;
    const double act[2] = {(double)actions[2 * k], (double)actions[2 * k + 1]};
    orc_step(m, dp, &e->s, &e->step_count, &e->px, &e->pz, act, action_mode, wheel_dist, env5, frame_skip, dt,
             max_steps, robot_speed, &o);
    reward[k] = (float)o.reward;
    done[k] = o.done;
    double rx = o.pos_x, rz = o.pos_z, ra = o.angle;
    if (o.done) {  /* auto-reset to the spawn pose; obs is the first frame of the new episode */
      orc_env_init(m, e, e->px0, e->pz0, e->ang0);
      rx = e->px0; rz = e->pz0; ra = e->ang0;
    }
    if (render) orr_render(sc, rx, rz, ra, &eps[k], W, H, 0, 0, 0, obs + (size_t)k * W * H * 3);
  }
}
