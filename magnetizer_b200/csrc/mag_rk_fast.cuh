// mag_rk_fast.cuh -- the time-stepping loop of one snapshot (dynamo.f90:168-212):
// nsteps x { update_floor_sign ; rk (3 stages) ; impose_bc }.
#pragma once
#include "mag_galaxy.cuh"

namespace magdev {

// per-warp shared memory
struct RkShared {
  RngState rng;
};

// update_floor_sign (floor_field.f90:99-114)
__device__ __forceinline__ void update_floor_sign(Gal& G, double time) {
  if (time - G.t_last_sign_choice > __dmul_rn(G.a->P.p_floor_kappa, G.tau_min)) {
    (void)random_sign(G);  // Bfloor_sign: its square cancels in the target field
    G.t_last_sign_choice = time;
    G.refresh_sign = true;
  }
}

// Returns ok (false: the field blew up, check_f_error).
__device__ inline bool run_timesteps(Gal& G, RkShared* rs) {
  (void)rs;
  const double dtg = __dmul_rn(G.dt, G.a->U.t0_Gyr);
  for (int jt = 1; jt <= G.nsteps; jt++) {
    const double this_t = __dadd_rn(G.t_Gyr, __dmul_rn(dtg, (double)jt));
    update_floor_sign(G, this_t);
    const bool ok = rk_generic(G);
    impose_bc(G, G.A(A_F0), G.A(A_F1), G.A(A_F2));
    if (!ok) return false;
  }
  return true;
}

}  // namespace magdev
