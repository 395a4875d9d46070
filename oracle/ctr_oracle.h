/* ctr_oracle.h — CPU oracle of the reference forward kinematics.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libctrfk.so) never links, loads or calls it.
 *
 * PARITY PINNED AGAINST THE REFERENCE'S OWN CODE, modulo Erl: the reference ships no tests, golden
 * vectors or fixtures, but its sources compile here against stand-in headers for its three absent
 * third-party dependencies (oracle/shim; recipe: oracle/Makefile target `ref` -> oracle/_ref/
 * libctr_ref.so).  This file's outputs equal that library's BIT FOR BIT — sample counts, steps,
 * every frame of every sample, radii, Jacobians, stability verdicts, on random and edge inputs, all
 * 14 section layouts (tests/test_reference_build.py; committed outputs of the reference build:
 * tests/golden/reference_golden.npz).  Still assumed, because Erl's source is unavailable (unpinned
 * gitlab HEAD, install.sh:41): Transform/Rotmat constructor argument order, the SE(3) product,
 * Erl::equal's tolerance, nearestOrthogonal (the [ASSUMED] items of oracle/shim/mini_erl.h).
 * Further pins: an independent pure-Python restatement (oracle/ctr_oracle_py.py), the sanity
 * vectors of SURVEY.md §9.2 and closed-form known-answer tests (tests/test_oracle.py).
 */
#ifndef CTR_ORACLE_H
#define CTR_ORACLE_H

#include <stdint.h>
#include "../include/ctr_fk.h"   /* ctr_robot_desc (POD only) */

#ifdef __cplusplus
extern "C" {
#endif

/* One configuration.  All outputs nullable.
 *   tip12      : m_CentreLineTransform[m_Samples-1], column-major 3x4
 *   backbone   : [max_samples][12], entries [0, n_out) written
 *   radius     : [max_samples]
 *   curv       : [max_samples][3]  m_CurvatureInnermostTube (u_x,u_y,u_z per sample)
 *   alpha_samp : [max_samples]     m_AlphaInnermostTube(:,0)
 *   alpha_base : [ntubes]          m_SampleTubeAlpha after the sweep
 * Returns CTR_FK_STATUS_* (0 = OK) or a negative CTR_FK_E_* code. */
int ctr_oracle_fk(const ctr_robot_desc* desc, const double* jv, int mode, double step,
                  int32_t nsamples, double* tip12, double* backbone, double* radius,
                  int32_t* n_out, double* step_out, double* curv, double* alpha_samp,
                  double* alpha_base);

/* B configurations with `nthreads` OpenMP threads (one scratch per thread, mirroring the
 * one-Robot-per-thread usage of the reference).  backbone [B][max_samples][12] etc. */
int ctr_oracle_fk_batch(const ctr_robot_desc* desc, const double* jv, int64_t B, int mode,
                        double step, int32_t nsamples, int nthreads, double* tip,
                        double* backbone, double* radius, int32_t* n_out, double* step_out,
                        double* alpha_base, int32_t* status);

/* Finite-difference tip Jacobian, robot_i.h:917-960 + :41-56.  jac[n_jv][6] column-major 6 x n_jv. */
int ctr_oracle_jacobian(const ctr_robot_desc* desc, const double* jv, int32_t nsamples,
                        double fd_trans, double fd_rot, double* jac, double* tip12);

/* Every calcJacobian form (robot_i.h:901-1072): mode STEP = the ArcLengthStep form (:901-907: base pose from
 * calcKinematic(JV, step), perturbed solves in NSamples mode with N = m_Samples); i_sample >= 0 = the form that
 * also returns the Jacobian of sample i's frame (:908-915, :1012-1072).  jac_*: 6 x n_jv column-major. */
int ctr_oracle_jacobian_ex(const ctr_robot_desc* desc, const double* jv, int mode, double step, int32_t nsamples,
                           int32_t i_sample, double fd_trans, double fd_rot, double* jac_tip, double* jac_sample,
                           double* tip12, double* sample12, int32_t* n_out);

/* calcStability, robot_i.h:756-818.  Returns 1 stable, 0 unstable, <0 error. */
int ctr_oracle_stability(const ctr_robot_desc* desc, const double* jv, double angle_step,
                         double arc_step);

/* calcDegreeStability, robot_i.h:819-899. */
int ctr_oracle_degree_stability(const ctr_robot_desc* desc, const double* jv, double angle_step,
                                double arc_step, double min_degree, double* degree, int* stable_out);

/* Batch forms (OpenMP over configurations).  rc_out[i] = return code of configuration i (Jacobian); stable[i] < 0 =
 * error code (stability). */
int ctr_oracle_jacobian_batch(const ctr_robot_desc* desc, const double* jv, int64_t B, int mode, double step,
                              int32_t nsamples, int32_t i_sample, double fd_trans, double fd_rot, int nthreads,
                              double* jac_tip, double* jac_sample, double* tip, int32_t* n_out, int32_t* rc_out);
int ctr_oracle_stability_batch(const ctr_robot_desc* desc, const double* jv, int64_t B, double angle_step,
                               double arc_step, double min_degree, int nthreads, int32_t* stable, double* degree);

int ctr_oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
