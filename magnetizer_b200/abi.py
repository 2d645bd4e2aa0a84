"""ctypes mirror of include/magnetizer_b200.h (struct layouts and enums only).

The names follow the reference: quantities are those of ts_arrays.f90:31-66, inputs
those of input_parameters.f90:158-170, parameters those of global_input_parameters.f90.
"""
from __future__ import annotations

import ctypes as C

MAG_ABI_VERSION = 3
MAG_NINPUT = 13
MAG_INVALID = -99999.0

PROFILE_NAMES = ("Br", "Bp", "alp_m", "Bzmod", "h", "Omega", "Omega_b", "Omega_d", "Omega_h", "Shear", "l", "v",
                 "etat", "tau", "alp_k", "alp", "Uz", "Ur", "n", "Beq", "rkpc")
SCALAR_NAMES = ("t_Gyr", "dt", "rmax", "Bmax", "Bmax_idx", "Bavg", "Beavg")
PROFILE_ID = {n: i for i, n in enumerate(PROFILE_NAMES)}
SCALAR_ID = {n: i for i, n in enumerate(SCALAR_NAMES)}

MAG_OK = 0
MAG_ERR_NO_DEVICE = -1
MAG_ERR_BAD_ARGUMENT = -2
MAG_ERR_UNSUPPORTED = -3
MAG_ERR_CUDA = -4
MAG_ERR_NX_OVERFLOW = -5
ERROR_NAMES = {0: "MAG_OK", -1: "MAG_ERR_NO_DEVICE", -2: "MAG_ERR_BAD_ARGUMENT", -3: "MAG_ERR_UNSUPPORTED",
               -4: "MAG_ERR_CUDA", -5: "MAG_ERR_NX_OVERFLOW"}

MAG_RNG_XOSHIRO256SS = 0
MAG_RNG_XORSHIFT1024S = 1

# p_outflow_type (outflow.f90:63-95)
OUTFLOW_TYPES = ("no_outflow", "vturb", "wind", "superbubble_simple", "superbubble")
# p_seed_choice (seed_field.f90:37-62); 'floor' is not implemented
SEED_CHOICES = ("random", "fraction", "decaying")

MAG_FLAG_UNINIT_R0 = 1
MAG_FLAG_SECANT = 2
MAG_FLAG_RETRY = 4
MAG_FLAG_STALE_PROFILES = 8


class MagParams(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("nsteps_0", C.c_int32),
        ("p_nsteps_max", C.c_int32),
        ("p_MAX_FAILS", C.c_int32),
        ("p_random_seed", C.c_int32),
        ("p_no_magnetic_fields_test_run", C.c_int32),
        ("p_courant_v", C.c_double),
        ("p_courant_eta", C.c_double),
        ("p_nx_ref", C.c_int32),
        ("p_nx_MAX", C.c_int32),
        ("p_rmax_over_rdisk", C.c_double),
        ("frac_seed", C.c_double),
        ("C_alp", C.c_double),
        ("C_floor", C.c_double),
        ("p_floor_kappa", C.c_double),
        ("fmag", C.c_double),
        ("R_kappa", C.c_double),
        ("p_ISM_sound_speed_km_s", C.c_double),
        ("p_ISM_kappa", C.c_double),
        ("p_ISM_xi", C.c_double),
        ("p_ISM_gamma", C.c_double),
        ("p_ISM_turbulent_length", C.c_double),
        ("p_stellarHeightToRadiusScale", C.c_double),
        ("p_molecularHeightToRadiusScale", C.c_double),
        ("p_gasScaleRadiusToStellarScaleRadius_ratio", C.c_double),
        ("p_Rmol_alpha", C.c_double),
        ("p_Rmol_P0", C.c_double),
        ("p_rreg_to_rdisk", C.c_double),
        ("p_minimum_density", C.c_double),
        ("p_rdisk_min", C.c_double),
        ("Mgas_disk_min", C.c_double),
        ("rmin_over_rmax", C.c_double),
        ("p_ignore_satellite_DM_haloes", C.c_int32),
        ("p_use_Pdm", C.c_int32),
        ("p_use_Pbulge", C.c_int32),
        ("rng_kind", C.c_int32),
        ("p_variable_timesteps", C.c_int32),
        ("p_simplified_pressure", C.c_int32),
        ("p_halo_contraction", C.c_int32),
        ("p_enable_P2", C.c_int32),
        ("Dyn_quench", C.c_int32),
        ("Alg_quench", C.c_int32),
        ("Damp", C.c_int32),
        ("lFloor", C.c_int32),
        ("p_space_varying_floor", C.c_int32),
        ("p_time_varying_floor", C.c_int32),
        ("p_limit_turbulent_scale", C.c_int32),
        ("p_allow_positive_shears", C.c_int32),
        ("seed_choice", C.c_int32),
        ("outflow_type", C.c_int32),
        ("p_neumann_boundary_condition_rmax", C.c_int32),
        ("p_use_fixed_physical_grid", C.c_int32),
        ("p_nn_seed", C.c_int32),
        ("reserved", C.c_int32 * 3),
        ("p_outflow_Lsn", C.c_double),
        ("p_outflow_fOB", C.c_double),
        ("p_outflow_etaSN", C.c_double),
        ("p_tOB", C.c_double),
        ("p_N_SN1OB", C.c_double),
        ("p_outflow_hot_gas_density", C.c_double),
        ("p_outflow_nu0", C.c_double),
        ("p_outflow_alphahot", C.c_double),
        ("p_outflow_Vhot", C.c_double),
        ("p_r_seed_decay", C.c_double),
    ]

    def as_dict(self):
        return {n: (list(getattr(self, n)) if n == "reserved" else getattr(self, n)) for n, _ in self._fields_}


class MagOutputs(C.Structure):
    _fields_ = [
        ("profiles", C.c_void_p),
        ("scalars", C.c_void_p),
        ("status", C.c_void_p),
        ("nsteps", C.c_void_p),
        ("error", C.c_void_p),
        ("point_stages", C.c_void_p),
        ("flags", C.c_void_p),
    ]
