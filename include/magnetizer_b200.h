/*
 * magnetizer_b200.h -- C ABI of the B200-native per-galaxy galactic-dynamo solve.
 *
 * This is the drop-in boundary for ONE path of luizfelippesr/magnetizer: the
 * operator that `jobs_distribute(work_routine, write_routine, ...)` farms out
 * (reference source/distributor.f90:39-53,180; invoked at source/Magnetizer.f90:36
 * with work_routine = dynamo_run, source/dynamo.f90:39).  In the reference the
 * operator communicates through module globals (ts_arrays::ts_data,
 * ts_arrays::ts_status_code, input_params::nsteps; inputs are pulled with
 * IO_read_dataset_scalar, source/input_parameters.f90:148-170).  Here the same data
 * crosses the boundary as plain caller-owned buffers so that a Fortran
 * ISO_C_BINDING shim or the C++ host driver can fill
 * ts_data / ts_status_code / nsteps and call the unchanged write_output
 * (source/output.f90:23).
 *
 * All entry points are extern "C", take plain pointers and sizes, never retain or
 * free caller memory, and return 0 on success or a negative MAG_ERR_* code for
 * CUDA / configuration errors.  Physics failures are never errors: they are status
 * characters per snapshot, exactly as in the reference (doc/status_codes.md,
 * source/messages.f90:112-151).
 *
 * There is no CPU fallback: every solve entry point fails with MAG_ERR_NO_DEVICE
 * when no CUDA device is usable.
 */
#ifndef MAGNETIZER_B200_H
#define MAGNETIZER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAG_ABI_VERSION 3

/* Number of per-galaxy input time series, in the order read by
 * load_input_parameters (source/input_parameters.f90:158-170). */
#define MAG_NINPUT 13
enum mag_input_column {
  MAG_IN_R_DISK = 0, MAG_IN_V_DISK, MAG_IN_R_BULGE, MAG_IN_V_BULGE, MAG_IN_R_HALO,
  MAG_IN_V_HALO, MAG_IN_NFW_CS1, MAG_IN_MGAS_DISK, MAG_IN_MSTARS_DISK, MAG_IN_SFR,
  MAG_IN_MHALO, MAG_IN_MSTARS_BULGE, MAG_IN_CENTRAL
};

/* Profile quantities, in the order of profile_names (source/ts_arrays.f90:39-66).
 * Values are in CODE units, as stored in ts_data; the unit conversion of
 * write_output (source/output.f90:63-201) stays on the host side. */
#define MAG_NPROFILE 21
enum mag_profile {
  MAG_Q_BR = 0, MAG_Q_BP, MAG_Q_ALP_M, MAG_Q_BZMOD, MAG_Q_H, MAG_Q_OMEGA, MAG_Q_OMEGA_B,
  MAG_Q_OMEGA_D, MAG_Q_OMEGA_H, MAG_Q_SHEAR, MAG_Q_L, MAG_Q_V, MAG_Q_ETAT, MAG_Q_TAU,
  MAG_Q_ALP_K, MAG_Q_ALP, MAG_Q_UZ, MAG_Q_UR, MAG_Q_N, MAG_Q_BEQ, MAG_Q_RKPC
};

/* Scalar quantities, in the order of scalar_names (source/ts_arrays.f90:31-38).
 * t_Gyr is never set by the reference and therefore always MAG_INVALID. */
#define MAG_NSCALAR 7
enum mag_scalar {
  MAG_S_T_GYR = 0, MAG_S_DT, MAG_S_RMAX, MAG_S_BMAX, MAG_S_BMAX_IDX, MAG_S_BAVG, MAG_S_BEAVG
};

/* Value of never-written entries (source/tsDataObj.f90:5). */
#define MAG_INVALID (-99999.0)

/* Error codes (return values). */
#define MAG_OK 0
#define MAG_ERR_NO_DEVICE (-1)     /* no usable CUDA device / extension not built for it */
#define MAG_ERR_BAD_ARGUMENT (-2)
#define MAG_ERR_UNSUPPORTED (-3)   /* parameter combination outside the implemented path */
#define MAG_ERR_CUDA (-4)          /* CUDA runtime error; see mag_last_error() */
#define MAG_ERR_NX_OVERFLOW (-5)   /* grid extension beyond p_nx_MAX (reference: stop, grid.f90:222-226) */

/* RNG flavour of the libgfortran the reference would have been built with
 * (SURVEY.md App. B): intrinsic random_number is xoshiro256** for GCC >= 9 and
 * xorshift1024* for GCC 7/8.  The default is xorshift1024*: it is the one that
 * reproduces the Bmax values published in the reference's tutorial notebook
 * (doc/magnetizer_tutorial.ipynb:673) and the one whose stream is pinned against a
 * real libgfortran (tests/golden/rng_golden.json). */
#define MAG_RNG_XOSHIRO256SS 0
#define MAG_RNG_XORSHIFT1024S 1

/*
 * POD mirror of the namelist values used on the path
 * (source/global_input_parameters.f90).  Fill with mag_params_default() -- which
 * applies the float-rounded default literals of the reference (e.g. 0.09 ->
 * 0.0900000035762787) -- then overwrite whatever the namelist file sets (values
 * read from a namelist are parsed as doubles).
 */
typedef struct mag_params {
  int32_t abi_version;            /* = MAG_ABI_VERSION */
  /* run_parameters (:66-85) */
  int32_t nsteps_0;               /* 1000 */
  int32_t p_nsteps_max;           /* 20000 */
  int32_t p_MAX_FAILS;            /* 6 */
  int32_t p_random_seed;          /* 17 */
  int32_t p_no_magnetic_fields_test_run; /* 0 */
  double p_courant_v;             /* 0.09f */
  double p_courant_eta;           /* 0.09f */
  /* grid_parameters (:93-120) */
  int32_t p_nx_ref;               /* 101 */
  int32_t p_nx_MAX;               /* 10000 */
  double p_rmax_over_rdisk;       /* 2.75 */
  /* dynamo_parameters (:127-202) */
  double frac_seed;               /* 0.01 */
  double C_alp;                   /* 1 */
  double C_floor;                 /* 1 */
  double p_floor_kappa;           /* 10 */
  double fmag;                    /* 0.5 */
  double R_kappa;                 /* 1 */
  /* ISM_and_disk_parameters (:205-299) */
  double p_ISM_sound_speed_km_s;  /* 10 */
  double p_ISM_kappa;             /* 1 */
  double p_ISM_xi;                /* 0.25 */
  double p_ISM_gamma;             /* 5/3 */
  double p_ISM_turbulent_length;  /* 0.1f */
  double p_stellarHeightToRadiusScale;   /* 1/7.3f */
  double p_molecularHeightToRadiusScale; /* 0.032f */
  double p_gasScaleRadiusToStellarScaleRadius_ratio; /* 1 */
  double p_Rmol_alpha;            /* 0.92 */
  double p_Rmol_P0;               /* 4.787e-12 */
  double p_rreg_to_rdisk;         /* 0.1f */
  double p_minimum_density;       /* 1e-30 */
  double p_rdisk_min;             /* 0.5 */
  double Mgas_disk_min;           /* 1e4 */
  double rmin_over_rmax;          /* 0.001f */
  int32_t p_ignore_satellite_DM_haloes; /* 0 */
  int32_t p_use_Pdm;              /* 1 */
  int32_t p_use_Pbulge;           /* 1 */
  /* which libgfortran RNG to reproduce */
  int32_t rng_kind;               /* MAG_RNG_XORSHIFT1024S */
  /*
   * Non-default physics switches.  The B200 path implements the reference's default
   * branch; mag_create() returns MAG_ERR_UNSUPPORTED when any of these differs
   * from the stated default (SURVEY.md 8(f3) lists them as "next").
   */
  int32_t p_variable_timesteps;   /* 1 */
  int32_t p_simplified_pressure;  /* 1 */
  int32_t p_halo_contraction;     /* 1 */
  int32_t p_enable_P2;            /* 0 */
  int32_t Dyn_quench;             /* 1; 0 (nvar = 2, no alp_m equation, no 'alp_m' rows: var module, gutsdynamo.f90:28-43) runs on the generic path */
  int32_t Alg_quench;             /* 0; 1 is implemented on the generic (any-nx) path: alp = alp_k/(1+B^2/Beq^2), gutsdynamo.f90:202-203 */
  int32_t Damp;                   /* 0 */
  int32_t lFloor;                 /* 1 */
  int32_t p_space_varying_floor;  /* 1 */
  int32_t p_time_varying_floor;   /* 1 */
  int32_t p_limit_turbulent_scale;/* 1 */
  int32_t p_allow_positive_shears;/* 0 */
  int32_t seed_choice;            /* p_seed_choice, seed_field.f90:37-62: MAG_SEED_RANDOM (0, default), MAG_SEED_FRACTION, MAG_SEED_DECAYING;
                                     'floor' is not implemented */
  /* outflow_parameters (:303-330), source/outflow.f90:24-121 -- all five prescriptions are
   * implemented: Uz enters the hoisted coefficients of the RHS (vertical advection of Br, Bp,
   * alp_m; floor-field source; critical dynamo number) and the 'Uz' output row */
  int32_t outflow_type;           /* MAG_OUTFLOW_*: 0 = 'no_outflow' */
  int32_t p_neumann_boundary_condition_rmax; /* 0; 1: dBr/dr = dBp/dr = 0 at r = R (impose_bc, gutsdynamo.f90:63-74), generic path */
  int32_t p_use_fixed_physical_grid;         /* 0; 1: r_max = p_rmax_over_rdisk x the largest r_disk of the history, adjust_grid does
                                                nothing (grid.f90:69-73, :159-162) */
  int32_t p_nn_seed;                         /* 2 ('decaying' seed) */
  int32_t reserved[3];
  double p_outflow_Lsn;           /* 1d38 */
  double p_outflow_fOB;           /* 0.7f */
  double p_outflow_etaSN;         /* 9.4d-3 */
  double p_tOB;                   /* 3 */
  double p_N_SN1OB;               /* 40 */
  double p_outflow_hot_gas_density; /* 1.7d-27 */
  double p_outflow_nu0;           /* 0.5 */
  double p_outflow_alphahot;      /* -3.2f */
  double p_outflow_Vhot;          /* 425 */
  double p_r_seed_decay;          /* 15 kpc ('decaying' seed, global_input_parameters.f90:151) */
} mag_params_t;

/* p_seed_choice (source/seed_field.f90:37-62) */
#define MAG_SEED_RANDOM 0
#define MAG_SEED_FRACTION 1
#define MAG_SEED_DECAYING 2

/* p_outflow_type (source/global_input_parameters.f90:306, source/outflow.f90:63-95) */
#define MAG_OUTFLOW_NONE 0
#define MAG_OUTFLOW_VTURB 1
#define MAG_OUTFLOW_WIND 2
#define MAG_OUTFLOW_SUPERBUBBLE_SIMPLE 3
#define MAG_OUTFLOW_SUPERBUBBLE 4

/* Writes the reference defaults (with float-rounded literals) into *p. */
void mag_params_default(mag_params_t *p);

/* Caller-owned result buffers of one solve_batch call.  Any pointer may be NULL
 * to skip that output. */
typedef struct mag_outputs {
  /* [n_profile_q][ngal][nz][p_nx_ref], radial index fastest.  Entry for snapshot iz
   * (0-based) of galaxy g = what ts_data%value(iz+1, :) holds in the reference after
   * dynamo_run (source/ts_arrays.f90:76-203).  Never-written snapshots hold
   * MAG_INVALID. */
  double *profiles;
  /* [n_scalar_q][ngal][nz] */
  double *scalars;
  /* [ngal][nz] status characters (ts_status_code, source/ts_arrays.f90:70) */
  char *status;
  /* [ngal] input_params::nsteps after the run (written as Log/nsteps, output.f90:43) */
  int32_t *nsteps;
  /* [ngal] 1 where dynamo_run would return error=.true./runtime<0, i.e. nothing is
   * written for the galaxy (source/dynamo.f90:69-73); else 0 */
  int32_t *error;
  /* [ngal] device-side cost of the galaxy in RK point-stages:
   * sum over snapshots of (time steps actually taken)*3*nx   */
  int64_t *point_stages;
  /* [ngal] diagnostic flags, MAG_FLAG_* */
  uint32_t *flags;
} mag_outputs_t;

/* A radius OTHER than the centre point hit the reference's use of an uninitialised r0
 * in the halo contraction (source/rotationCurves.f90:266-269); this implementation uses
 * r0 = r there.  (At the centre point the value is overwritten by regularize(),
 * profiles.f90:318-330, so it cannot matter and is not flagged.) */
#define MAG_FLAG_UNINIT_R0 1u
/* The secant fallback of the hydrostatic solve was used (pressureEquilibrium.f90:180) */
#define MAG_FLAG_SECANT 2u
/* At least one time-stepping retry happened (status 'm', dynamo.f90:228) */
#define MAG_FLAG_RETRY 4u
/* The galaxy's first valid snapshot had negligible gas: profile arrays that the
 * reference leaves stale from the previous galaxy are zero here */
#define MAG_FLAG_STALE_PROFILES 8u

typedef struct mag_ctx mag_ctx_t;

/* Creates a solver context on CUDA device `device` (-1: current device).
 * Replaces the per-rank module state of the reference. */
int mag_create(const mag_params_t *params, int device, mag_ctx_t **ctx_out);
void mag_destroy(mag_ctx_t *ctx);

/*
 * Solves `ngal` galaxy histories: the batched equivalent of calling
 * runtime = dynamo_run(gal_id, test_run, error) (source/dynamo.f90:39) for each id
 * and keeping the module globals write_output reads.
 *
 *  gal_ids  [ngal]              1-based galaxy ids (only used for the RNG seed,
 *                               source/random.f90:43), int32
 *  nz                           number of snapshots (number_of_redshifts)
 *  t_gyr    [nz]                'Input/t' (source/input_parameters.f90:148)
 *  inputs   [MAG_NINPUT][ngal][nz]  HOST buffer, snapshot index fastest
 *  profile_q[n_profile_q], scalar_q[n_scalar_q]  requested quantities
 *  out                          HOST buffers
 *
 * Host<->device copies happen inside the call.  If out->profiles / scalars / status are
 * pinned (page-locked, mapped) host memory -- cudaHostAlloc, cudaHostRegister, or e.g.
 * torch pin_memory() -- the kernels store the rows directly into them while the solve runs
 * (every row is stored exactly once) and no staging copy follows; pageable arrays, such as
 * the Fortran allocatables of the reference's modules, are staged in HBM and copied after
 * the solve.  MAG_ZERO_COPY=0 in the environment forces the staged path.
 */
int mag_solve_batch(mag_ctx_t *ctx, int32_t ngal, const int32_t *gal_ids, int32_t nz,
                    const double *t_gyr, const double *inputs,
                    int32_t n_profile_q, const int32_t *profile_q,
                    int32_t n_scalar_q, const int32_t *scalar_q,
                    const mag_outputs_t *out);

/*
 * Same, with `inputs`, `gal_ids`, `t_gyr` and every pointer of `out` being DEVICE
 * pointers (already resident in HBM), launched on `cuda_stream` (a cudaStream_t
 * passed as void*; NULL = default stream).  Returns after enqueueing; the caller
 * synchronises the stream.
 */
int mag_solve_batch_device(mag_ctx_t *ctx, int32_t ngal, const int32_t *d_gal_ids, int32_t nz,
                           const double *d_t_gyr, const double *d_inputs,
                           int32_t n_profile_q, const int32_t *profile_q,
                           int32_t n_scalar_q, const int32_t *scalar_q,
                           const mag_outputs_t *d_out, void *cuda_stream);

/*
 * Range forms, for callers that hold the whole catalogue in one set of arrays and solve it
 * piecewise (the multi-GPU partition below, a driver that streams a catalogue larger than a
 * GPU's record memory): `gal_ids`, `inputs` and the arrays of `out` are FULL-catalogue arrays
 * ([..][ngal_total][..], the layouts of mag_solve_batch); only the rows of catalogue galaxies
 * [g0, g0+n) are read and written, in place -- no packing copy on either side.
 *
 * mag_solve_range_async (HOST arrays) enqueues the copies and kernels and returns; mag_wait
 * blocks until the call is complete, repeats it with a larger record arena when a grid grew
 * beyond the budget, and returns the call's status.  One call may be pending per context.
 * mag_solve_batch(...) == mag_solve_range_async(ngal, 0, ngal, ...) + mag_wait.
 */
int mag_solve_range_async(mag_ctx_t *ctx, int32_t ngal_total, int32_t g0, int32_t n,
                          const int32_t *gal_ids, int32_t nz, const double *t_gyr,
                          const double *inputs, int32_t n_profile_q, const int32_t *profile_q,
                          int32_t n_scalar_q, const int32_t *scalar_q, const mag_outputs_t *out);
int mag_wait(mag_ctx_t *ctx);

/* The same with DEVICE arrays (mag_solve_batch_device == range [0, ngal)). */
int mag_solve_range_device(mag_ctx_t *ctx, int32_t ngal_total, int32_t g0, int32_t n,
                           const int32_t *d_gal_ids, int32_t nz, const double *d_t_gyr,
                           const double *d_inputs, int32_t n_profile_q, const int32_t *profile_q,
                           int32_t n_scalar_q, const int32_t *scalar_q,
                           const mag_outputs_t *d_out, void *cuda_stream);

/* Device-side status of the last mag_solve_batch_device / mag_solve_range_device call on this
 * context, to be called after the caller has synchronised its stream: MAG_OK,
 * MAG_ERR_NX_OVERFLOW (a grid grew beyond p_nx_MAX / the workspace; the reference stops,
 * grid.f90:222-226) or MAG_ERR_CUDA (record arena exhausted: results incomplete).  The
 * host-buffer calls report the same conditions through their return value. */
int mag_last_device_status(mag_ctx_t *ctx);

/* Page-locked host memory that every GPU of the box can store into (cudaHostAlloc, portable +
 * mapped).  Output arrays allocated here are written by the kernels directly while the solve
 * runs; any other page-locked memory (cudaHostRegister, torch pin_memory) works as well. */
void *mag_host_alloc(size_t bytes);
void mag_host_free(void *p);

/*
 * Multi-GPU form: replaces the MPI master/worker farm of jobs_distribute
 * (source/distributor.f90:253-384) for this operator.  The catalogue is cut into chunks of
 * `chunk` galaxies (chunk <= 0: one chunk per device, cut into equal pieces of at most 25 000); GPU number d of `devices[ndev]` owns the chunks k with k % ndev == d
 * (static share) and steals unclaimed chunks from the tails of the other GPUs' shares when its
 * own is exhausted.  Every GPU is driven by TWO host threads with one solver context each, so
 * that the profile phase and the head of one chunk fill the SMs that the tail of the previous
 * chunk leaves idle; inside a chunk the galaxies are scheduled longest-first with the exact
 * cost known after the profile phase, the longest ones on four-warp teams.  Inputs are read
 * from and results stored into the caller's full-catalogue HOST arrays in place (layouts of
 * mag_solve_batch): that IS the host-side gather, in galaxy order.  Page-locked output arrays
 * (mag_host_alloc) are stored by the kernels directly; pageable ones are staged in HBM.  There
 * is no collective: galaxies are independent.
 *
 * mag_partition_create builds the contexts and worker threads once; mag_partition_solve runs
 * one catalogue; stats (may be NULL) receives per device chunks_done[ndev], chunks_stolen[ndev].
 * mag_solve_partitioned = create + solve + destroy.  The same device may be listed more than once.
 */
typedef struct mag_partition mag_partition_t;
int mag_partition_create(const mag_params_t *params, int32_t ndev, const int32_t *devices,
                         mag_partition_t **part_out);
void mag_partition_destroy(mag_partition_t *part);
int mag_partition_solve(mag_partition_t *part, int32_t ngal, const int32_t *gal_ids, int32_t nz,
                        const double *t_gyr, const double *inputs,
                        int32_t n_profile_q, const int32_t *profile_q,
                        int32_t n_scalar_q, const int32_t *scalar_q, const mag_outputs_t *out,
                        int32_t chunk, int32_t *chunks_done, int32_t *chunks_stolen);
/*
 * The same partition with the catalogue RESIDENT IN HBM: d_gal_ids[d], d_t_gyr[d], d_inputs[d]
 * are device d's own full copies of the inputs, d_out[d] its own full-size output arrays; the
 * rows of a galaxy end up on the device that solved it, owner[g] (HOST, [ngal], may be NULL)
 * says which (index into devices).
 */
int mag_partition_solve_device(mag_partition_t *part, int32_t ngal, const int32_t *const *d_gal_ids,
                               int32_t nz, const double *const *d_t_gyr, const double *const *d_inputs,
                               int32_t n_profile_q, const int32_t *profile_q,
                               int32_t n_scalar_q, const int32_t *scalar_q, const mag_outputs_t *d_out,
                               int32_t chunk, int32_t *chunks_done, int32_t *chunks_stolen, int32_t *owner);
/* Of the last solve on this partition: the chunk size used (chunk <= 0 in a solve call selects it
 * automatically), kernel launches per device and the device time
 * (ms, CUDA events) spent inside the kernels of each device's chunks -- where two chunks of a
 * device overlap, both count. */
int mag_partition_stats(mag_partition_t *part, int32_t *chunk_used, int64_t *launches, double *busy_ms);
int mag_solve_partitioned(const mag_params_t *params, int32_t ndev, const int32_t *devices,
                          int32_t ngal, const int32_t *gal_ids, int32_t nz, const double *t_gyr,
                          const double *inputs, int32_t n_profile_q, const int32_t *profile_q,
                          int32_t n_scalar_q, const int32_t *scalar_q, const mag_outputs_t *out,
                          int32_t chunk, int32_t *chunks_done, int32_t *chunks_stolen);
const char *mag_partition_last_error(void);

/* Per-phase device time (ms, CUDA events on the solve stream) of the last solve call on this
 * context, after the stream is synchronised: [0] profile phase (k_profiles, k_chains, queue
 * sort; first chunk), [1] all kernels (profile phase + evolve kernels), [2] H2D, [3] D2H.
 * Returns the number of kernel launches of that call. */
int mag_last_timing(mag_ctx_t *ctx, double ms_out[4]);

/* FP64 DFMA micro-benchmark on the context's device: returns achieved FLOP/s
 * (2 flops per DFMA) -- the FP64 roofline denominator, which MEASURED_PEAKS.json
 * does not carry. */
double mag_measure_fp64_peak(mag_ctx_t *ctx, int repeats);

/* Text of the last error on this thread. */
const char *mag_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* MAGNETIZER_B200_H */
