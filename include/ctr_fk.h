/* ctr_fk.h — C ABI of the B200 batched forward-kinematics engine for concentric
 * tube robots (CTR).
 *
 * This is the drop-in boundary.  Everything above it (the C++14 `ctrb200::Robot`
 * host object, Python ctypes bindings, a maintainer's patch to the reference's
 * `CTR::Robot<Real>`) talks to the CUDA kernels only through these entry points:
 * plain pointers and sizes, no C++ or torch types.
 *
 * Reference interfaces replaced (all paths relative to /root/reference/ctr_source):
 *   - RobotBase::calcKinematic(VectorJ, Real step, Real& stepRet)      robot_i.h:223, :690-721
 *   - RobotBase::calcKinematic(VectorJ, size_t NSamples, Real& stepRet) robot_i.h:226, :722-755
 *   - Robot::calcKinematic(Real* JV, const Real& step, Real* Tr)        ctr_robot.cpp:166-186
 *     (the raw-pointer form: JV array in, 3x4 column-major transform out) — ctr_fk_batch is
 *     this call over B joint vectors at once.
 *   - getNSamples/getTransformSample/getRadiusSample/getTip*            robot_i.h:520-533
 *     -> the optional backbone / radius / n_out outputs.
 *   - copy2opencl(local_constant_mem) robot descriptor                  robot_i.h:658-686
 *     -> ctr_robot_desc (same fields, InitTrafo column-major).
 *   - calcJacobian (finite differences), all four forms                 ctr_robot.h:121-126,
 *                                                                       robot_i.h:901-1072
 *     -> ctr_fk_jacobian_batch (NSamples form), ctr_fk_jacobian_batch_ex (every form).
 *   - calcStability / calcDegreeStability                               robot_i.h:756-899
 *     -> ctr_fk_stability_batch.
 *   - getRandomJointValues                                              robot_i.h:1308-1317,
 *                                                                       section_i.h:328-337
 *     -> ctr_fk_sample_joints (counter-based instead of mt19937; same distribution).
 *   - a live CTR::Robot<double> of the reference                        ctr_robot.h:76-173
 *     -> ctr_fk_reference_adapter.hpp (header-only): its ctr_robot_desc through the public interface.
 * Checked against the reference's own sources, compiled with stand-in third-party headers
 * (oracle/_ref, tests/test_reference_build.py).
 *
 * Conventions
 *   - Joint vector layout (robot_i.h:1126, section_i.h:204-209):
 *       jv[0] = RigidTheta; per section s: jv[1+s+t0] = Phi_s, jv[2+s+t0+k] = AlphaTip of
 *       tube k of the section (t0 = number of tubes in earlier sections).
 *       n_jv = 1 + nsections + ntubes.
 *   - Transforms are 12 doubles, column-major 3x4: r00 r10 r20 r01 r11 r21 r02 r12 r22 tx ty tz
 *     (Erl::Transform::getData(.., ColMajor), ctr_robot.cpp:171).
 *   - mode CTR_FK_MODE_STEP   : calcKinematic(JV, ArcLengthStep)  — N = floor(sum(Phi)/step)
 *     mode CTR_FK_MODE_NSAMPLE: calcKinematic(JV, NSamples)       — N = min(NSamples, MaxSamples)
 *   - Return value: 0 = OK, <0 = CTR_FK_E_* (bad argument / descriptor), >0 = cudaError_t.
 *     No exceptions cross this boundary.  ctr_fk_last_error() gives a thread-local message.
 *   - There is NO CPU fallback: every compute entry point fails with a CUDA error code when
 *     no device is usable.
 */
#ifndef CTR_FK_H
#define CTR_FK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTR_FK_MAX_SECTIONS 3   /* CTR_MAXSECTION, CMakeLists.txt:26-27 */
#define CTR_FK_MAX_TUBES    6   /* 2 tubes per section at most, ctr_static.h:162 */
#define CTR_FK_MAX_JV      (1 + CTR_FK_MAX_SECTIONS + CTR_FK_MAX_TUBES)

#define CTR_FK_MODE_STEP    0
#define CTR_FK_MODE_NSAMPLE 1

/* flags for ctr_fk_batch* */
#define CTR_FK_FLAG_STRICT     1u  /* reference operation order, libm-grade sin/cos/div/sqrt   */
#define CTR_FK_FLAG_NO_BINNING 2u  /* step mode: do not sort configurations by sample count    */
#define CTR_FK_FLAG_SOA        4u  /* backbone laid out [k][12][B] instead of [B][MaxS][12]     */

/* error codes (<0) */
#define CTR_FK_E_NULL      -1   /* null handle / required pointer                               */
#define CTR_FK_E_DESC      -2   /* invalid robot descriptor                                     */
#define CTR_FK_E_ARG       -3   /* invalid argument (mode, step<=0, nsamples<1, B<0 ...)         */
#define CTR_FK_E_XML       -4   /* XML file unreadable or malformed                             */
#define CTR_FK_E_NOMEM     -5   /* host allocation failed                                       */
#define CTR_FK_E_RANGE     -6   /* ctr_fk_single only: some Phi outside [0, Length]; outputs NaN */

/* per-configuration status bits written to the optional `status` output */
#define CTR_FK_STATUS_OK        0
#define CTR_FK_STATUS_PHI_RANGE 1   /* some Phi outside [0, Length] (reference: assert, section_i.h:205);
                                       that configuration's tip, step_out, alpha_base, frame 0 of its
                                       backbone row and radius 0 are NaN, its n_out is 0      */

/* One section = 1 tube (CC) or 2 tubes (VC), section_i.h:365-369. */
typedef struct ctr_section_desc {
    int32_t ntube;          /* 1 or 2 */
    int32_t reserved;
    double  length;         /* m_Length                                     */
    double  stiffness[2];   /* m_Stiffness      (entry [1] unused for CC)   */
    double  curvature[2];   /* m_Curvature      (only [0] enters the FK, section_i.h:320-325) */
    double  radius[2];      /* m_Radius         ([0] is the outer radius, section_i.h:151)    */
    double  poisson[2];     /* m_PoissonRatio                               */
} ctr_section_desc;

/* Robot descriptor; mirrors RobotCCPP of copy2opencl(local_constant_mem), robot_i.h:658-686. */
typedef struct ctr_robot_desc {
    int32_t nsections;      /* 1..3 */
    int32_t max_samples;    /* m_MaxSamples (>=1) */
    double  init_trafo[12]; /* m_InitTrafo, column-major 3x4, already orthogonalised            */
    double  zero_tol;       /* tolerance of Erl::equal(norm,0) in curvature2Transform
                               (robot_i.h:1245); 0 selects the default 1e-12                    */
    ctr_section_desc sec[CTR_FK_MAX_SECTIONS];
} ctr_robot_desc;

typedef struct ctr_fk_engine* ctr_fk_handle;

/* ---- descriptor helpers (host only, no CUDA) --------------------------------------------- */

/* Parse a ctr_resources XML file (V1 schema ctr_static.h:132-208, V2 schema :461-532, version
 * sniff :594-601).  phi_xml (nullable, CTR_FK_MAX_SECTIONS doubles) receives the per-section
 * <phi> values used for the construction-time FK (ctr_static.h:179,188,248-252). */
int ctr_fk_desc_from_xml(const char* path, int32_t max_samples, ctr_robot_desc* out, double* phi_xml);
/* Validate a descriptor (section count, tube counts, positive stiffness sums ...). */
int ctr_fk_desc_validate(const ctr_robot_desc* desc);
int ctr_fk_desc_ntubes(const ctr_robot_desc* desc);
int ctr_fk_desc_njoints(const ctr_robot_desc* desc);          /* getNJoints, robot_i.h:522 */
double ctr_fk_desc_max_length(const ctr_robot_desc* desc);    /* getMaxLength, robot_i.h:537-543 */
/* Nearest rotation (polar factor) of the 3x3 part of a column-major 3x4 transform, in place;
 * stands in for Erl::Transform::nearestOrthogonal (ctr_static.h:206).  Returns CTR_FK_E_ARG if
 * max|R^T R - I| exceeds `verify_tol` before projection (verifyOrthogonal, ctr_static.h:202). */
int ctr_fk_orthogonalize(double* trafo12, double verify_tol);

/* ---- engine lifetime ------------------------------------------------------------------- */

int ctr_fk_create(const ctr_robot_desc* desc, int device, ctr_fk_handle* out);
int ctr_fk_destroy(ctr_fk_handle h);
int ctr_fk_get_desc(ctr_fk_handle h, ctr_robot_desc* out);

/* ---- the hot path ---------------------------------------------------------------------- */

/* Batched FK, DEVICE pointers (all on the handle's device), asynchronous on `stream`
 * (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   jv        [B][n_jv]               in
 *   tip       [B][12]                 out, nullable: m_CentreLineTransform[m_Samples-1]
 *   backbone  [B][max_samples][12]    out, nullable: m_CentreLineTransform[0..n_out)
 *                                     (or [max_samples][12][B] with CTR_FK_FLAG_SOA);
 *                                     entries k >= n_out are unspecified (used as scratch)
 *   radius    [B][max_samples]        out, nullable: m_CentreLineRadius (needs backbone != NULL
 *                                     or may be requested alone)
 *   n_out     [B]                     out, nullable: m_Samples
 *   step_out  [B]                     out, nullable: _ArcLengthStepReturn
 *   alpha_base[B][ntubes]             out, nullable: m_SampleTubeAlpha after the sweep (the tube
 *                                     angles at the base; what calcStability reads, robot_i.h:810)
 *   status    [B]                     out, nullable: CTR_FK_STATUS_* bits
 * jv, tip and backbone must be 16-byte aligned (CTR_FK_E_ARG otherwise); a 32-byte aligned
 * backbone buffer lets the frames go out as full-sector 256-bit stores.
 */
int ctr_fk_batch(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step,
                 int32_t nsamples, uint32_t flags,
                 double* tip, double* backbone, double* radius, int32_t* n_out,
                 double* step_out, double* alpha_base, int32_t* status, void* stream);

/* Same computation with state arithmetic in FP32 (sample count, arc positions and the section
 * masks stay FP64 so that n_out matches the FP64 path exactly).  Inputs/outputs stay double. */
int ctr_fk_batch_f32(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step,
                     int32_t nsamples, uint32_t flags,
                     double* tip, int32_t* n_out, double* step_out, int32_t* status, void* stream);

/* Batched FK with HOST pointers (pageable or pinned).  Synchronous: returns when all outputs are in
 * host memory.  Two forms:
 *   - staged (default): the batch is cut into chunks of one wave of the solver kernel that are copied in, solved and
 *     copied out on rotating streams, so that both copy engines and the SMs work at the same time (page-locked
 *     buffers — cudaHostAlloc / cudaHostRegister / torch pin_memory() / ctr_fk_host_alloc — make the copies
 *     asynchronous; pageable memory works, slower).
 *   - zero-copy (CTR_FK_FLAG_HOST_ZEROCOPY, or CTRFK_HOST_ZEROCOPY=1 in the environment at ctr_fk_create; every
 *     buffer page-locked and 16-byte aligned, FAST arithmetic, no step-mode binning — otherwise the call falls back
 *     to the staged form): ONE kernel launch reads the joint vectors from and writes the poses to host memory
 *     directly, whole lines at a time; no device staging memory.  Results are bit-identical in both forms.
 *     CTR_FK_FLAG_HOST_STAGED overrides the environment.
 * Only tip / n_out / step_out / status are offered here (backbone output is 26 KB per configuration
 * and is meant to stay on the device; see ctr_fk_batch_host_backbone). */
#define CTR_FK_FLAG_HOST_STAGED   0x800u
#define CTR_FK_FLAG_HOST_ZEROCOPY 0x4000u
/* ctr_fk_batch_host only: the FP32 state arithmetic of ctr_fk_batch_f32 (chunked copy pipeline; doubles at the
 * interface) — what ctrb200::RobotF, the shape of the reference's CTR::Robot<float>, calls */
#define CTR_FK_FLAG_F32         0x2000u
int ctr_fk_batch_host(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step,
                      int32_t nsamples, uint32_t flags,
                      double* tip, int32_t* n_out, double* step_out, int32_t* status);

/* Page-locked, device-mapped host buffers for the calls above, placed for the engine's GPU: the pages are faulted in
 * while the calling thread is bound to the CPUs local to the GPU's PCIe root, so they land on the GPU's NUMA node
 * (on a multi-socket host a buffer on the far socket sends every transfer across the inter-socket link as well).
 * CTR_FK_HOST_WRITE_COMBINED: for buffers the host only ever writes (joint vectors); host reads of such memory are
 * slow.  ctr_fk_host_local_cpus writes those CPUs as a cpulist string ("0-15,32-47"; empty if unknown) for callers
 * that want to bind their own worker threads. */
#define CTR_FK_HOST_WRITE_COMBINED 1u
int ctr_fk_host_alloc(ctr_fk_handle h, size_t bytes, uint32_t flags, void** ptr);
int ctr_fk_host_free(ctr_fk_handle h, void* ptr);
int ctr_fk_host_local_cpus(ctr_fk_handle h, char* buf, size_t len);

/* HOST pointers, WITH the backbone: backbone [B][max_samples][12] (required; whole rows are copied back: the first
 * n_out[i] frames of row i are getTransformSample of robot_i.h:527, the entries behind them are unspecified scratch,
 * and a configuration with status != 0 has n_out 0 and NaN in frame 0), radius [B][max_samples] (nullable,
 * getRadiusSample :526; same rule), the rest
 * as ctr_fk_batch_host.  104 * max_samples bytes per configuration come back over PCIe, which bounds this call. */
int ctr_fk_batch_host_backbone(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step,
                               int32_t nsamples, uint32_t flags, double* tip, double* backbone,
                               double* radius, int32_t* n_out, double* step_out, int32_t* status);

/* Solve + all-gather in ONE kernel (the optional all-gather of the poses over NVLink).  Every rank
 * calls this on its own slice; each thread stores its pose into row `first_row + i` of EVERY destination buffer —
 * the gathered [world * B][12] buffers of all ranks, mapped into this process (CUDA IPC: ctr_fk_peer_alloc /
 * ctr_fk_peer_open below, or torch symmetric memory), so the transfer is NVLink peer stores issued by the solver
 * warps while the others keep the FP64 pipe busy; no separate collective, no staging copy.
 *   tip_dst   HOST array of `ndst` (1..8) DEVICE pointers, one per destination rank (this rank's own buffer included)
 *   flags     CTR_FK_FLAG_GATHER_MULTICAST: tip_dst[0] is an NVSwitch multicast address (ndst = 1); one
 *             multimem.st per 16 bytes, replicated to every GPU by the switch.  CTR_FK_FLAG_NO_BINNING as usual.
 *   first_row this rank's first row in the gathered buffer (ctr_fk shards contiguously: rank * B)
 * Stores to peers are complete when the kernel is; the caller synchronises the RANKS before reading the gathered
 * buffer (a barrier).  FAST arithmetic, tip poses only.  n_out/step_out/status: local [B], nullable. */
#define CTR_FK_FLAG_GATHER_MULTICAST 0x200u
#define CTR_FK_FLAG_GATHER_ROWWISE   0x400u   /* one thread stores its own row (16-byte NVLink writes) instead of the
                                                  CTA writing its 128 rows as whole lines; for comparison */
int ctr_fk_batch_gather(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step,
                        int32_t nsamples, uint32_t flags, void* const* tip_dst, int32_t ndst,
                        int64_t first_row, int32_t* n_out, double* step_out, int32_t* status,
                        void* stream);
/* Peer-visible device buffers for it, one process per GPU: ctr_fk_peer_alloc allocates `bytes` on the engine's
 * device and returns a 64-byte CUDA IPC handle to hand to the other ranks (any byte transport);
 * ctr_fk_peer_open maps a peer's buffer into this process (peer access is enabled on first use);
 * ctr_fk_peer_close / ctr_fk_peer_free undo them. */
int ctr_fk_peer_alloc(ctr_fk_handle h, size_t bytes, void** dptr, unsigned char* handle64);
int ctr_fk_peer_open(ctr_fk_handle h, const unsigned char* handle64, void** dptr);
int ctr_fk_peer_close(ctr_fk_handle h, void* dptr);
int ctr_fk_peer_free(ctr_fk_handle h, void* dptr);

/* One configuration, HOST pointers, synchronous: the reference's single-shot call
 * (robot_i.h:690-755) served by the same kernels with B = 1.  The handle keeps a small device
 * scratch for it; calls on one handle are serialised.  backbone [max_samples][12], radius
 * [max_samples]; every output nullable.  Returns CTR_FK_E_RANGE (outputs NaN, *n_out = 0) when the
 * configuration's status is CTR_FK_STATUS_PHI_RANGE — a code of its own, so that it cannot be taken for
 * cudaError_t 1. */
int ctr_fk_single(ctr_fk_handle h, const double* jv, int mode, double step, int32_t nsamples,
                  uint32_t flags, double* tip, double* backbone, double* radius, int32_t* n_out,
                  double* step_out, double* alpha_base);

/* Batched finite-difference tip Jacobian, NSamples form (robot_i.h:917-960): 6 x n_jv,
 * column-major per configuration: jac[B][n_jv][6].  DEVICE pointers.  tip nullable. */
int ctr_fk_jacobian_batch(ctr_fk_handle h, const double* jv, int64_t B, int32_t nsamples,
                          double fd_trans, double fd_rot, double* jac, double* tip, void* stream);

/* The same with HOST pointers (synchronous; staged through the device). */
int ctr_fk_jacobian_host(ctr_fk_handle h, const double* jv, int64_t B, int32_t nsamples,
                         double fd_trans, double fd_rot, double* jac, double* tip);

/* Every calcJacobian form of the reference (ctr_robot.h:121-126, robot_i.h:901-1072), DEVICE pointers:
 *   mode = CTR_FK_MODE_NSAMPLE, i_sample < 0   calcKinematic(JV, N) + calcJacobian(JV, Tip, N, ..)   (:917-960)
 *   mode = CTR_FK_MODE_STEP,    i_sample < 0   calcJacobian(JV, ArcLengthStep, ..) (:901-907): the base pose comes
 *                                              from calcKinematic(JV, step), the perturbed solves run in NSamples
 *                                              mode with N = m_Samples of that call — the reference differences two
 *                                              discretisations here; reproduced, not repaired
 *   i_sample >= 0                              calcJacobian(JV, N, .., iSample, JTip, JiSample) (:908-915,
 *                                              :1012-1072): jac_sample[B][n_jv][6] receives the Jacobian of the
 *                                              frame of sample i_sample, `sample` (nullable) that frame
 *   flags & CTR_FK_FLAG_JAC_GIVEN_POSES        calcJacobian(JV, TTip, N, ..) (:917) and (JV, TTip, TiSample, N, ..,
 *                                              iSample, ..) (:1012): `tip` (and `sample`) are INPUTS, the base
 *                                              poses the caller already has; no base solve is run
 * tip, sample [B][12]; n_out [B] = m_Samples of the base solve; all three nullable (unless given).  A
 * configuration whose i_sample is not below its sample count gets NaN in jac_sample/sample (the reference would
 * read a stale frame). */
#define CTR_FK_FLAG_JAC_GIVEN_POSES 0x100u
/* flags: CTR_FK_FLAG_JAC_GIVEN_POSES, CTR_FK_FLAG_NO_BINNING, CTR_FK_FLAG_STRICT (every solve in the reference's
 * operation order: the Jacobian then differs from the reference's own only through libm and the frame product's
 * association). */
int ctr_fk_jacobian_batch_ex(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step,
                             int32_t nsamples, int32_t i_sample, double fd_trans, double fd_rot,
                             uint32_t flags, double* jac_tip, double* jac_sample, double* tip,
                             double* sample, int32_t* n_out, void* stream);
int ctr_fk_jacobian_host_ex(ctr_fk_handle h, const double* jv, int64_t B, int mode, double step,
                            int32_t nsamples, int32_t i_sample, double fd_trans, double fd_rot,
                            uint32_t flags, double* jac_tip, double* jac_sample, double* tip,
                            double* sample, int32_t* n_out);

/* Batched calcStability / calcDegreeStability (robot_i.h:756-818, :819-899), DEVICE pointers.
 * For every tube 1..T-1 the tip angle is scanned in `angle_step` increments from 0 to its joint
 * value (or from the joint value to 2*pi when it exceeds pi); each scan point runs the tip->base
 * sweep (step mode, `arc_step`) and the configuration is unstable as soon as the tube's base
 * angle decreases between two scan points.  JV[0] is ignored (set to 0, robot_i.h:762).
 *   stable [B]  out, nullable: 1 stable, 0 unstable, -1 invalid joint values
 *   degree [B]  out, nullable: calcDegreeStability's return value, atan2(angle_step, smallest
 *               base-angle increment); `min_degree` is its _MinDegreeStability argument (pass -1
 *               for the reference's default) */
int ctr_fk_stability_batch(ctr_fk_handle h, const double* jv, int64_t B, double angle_step,
                           double arc_step, double min_degree, int32_t* stable, double* degree,
                           void* stream);
/* The same with HOST pointers (synchronous; staged through the device). */
int ctr_fk_stability_host(ctr_fk_handle h, const double* jv, int64_t B, double angle_step,
                          double arc_step, double min_degree, int32_t* stable, double* degree);
/* Both with flags: CTR_FK_FLAG_STRICT runs every sweep of the scan in the reference's operation order.  The verdict
 * is the SIGN of a base-angle difference between two scan points (robot_i.h:805, :873); FAST arithmetic moves that
 * difference by ~1e-13 rad, so a configuration whose smallest difference lies that close to zero can flip — STRICT
 * exists for callers who need the reference's verdict there too.
 * angle_step must advance a scan at 2 pi (2 pi + angle_step > 2 pi), else CTR_FK_E_ARG. */
int ctr_fk_stability_batch_ex(ctr_fk_handle h, const double* jv, int64_t B, double angle_step,
                              double arc_step, double min_degree, uint32_t flags, int32_t* stable,
                              double* degree, void* stream);
int ctr_fk_stability_host_ex(ctr_fk_handle h, const double* jv, int64_t B, double angle_step,
                             double arc_step, double min_degree, uint32_t flags, int32_t* stable,
                             double* degree);

/* Base-angle-actuated forward kinematics (not in the reference): the alpha
 * slots of `jv_base` hold the tube angles AT THE BASE (what the actuators set) instead of at the
 * tip.  A damped Newton shooting solve finds the tip angles whose sweep ends at those base angles (Jacobian from the
 * tangent-linear sweep — the variational equations of the discrete recurrence; backtracking line search on the
 * residual max-norm; step limiter 1 rad), then the ordinary solve produces the pose.
 * Sweep counts differ per configuration; the kernel is persistent and every lane draws its
 * next configuration from a global ticket counter as soon as it finishes (a work queue that absorbs the divergence).
 *   jv_tip [B][n_jv]  out, REQUIRED: jv_base with the alpha slots replaced by the tip angles found
 *                     (feeding it to ctr_fk_batch reproduces `tip`); tube 0's angle passes through.  On a status other
 *                     than OK it holds the best estimate reached (lowest residual accepted)
 *   tip    [B][12]    out, nullable
 *   iters  [B]        out, nullable: sweeps run
 *   status [B]        out, nullable: 0 converged (max |base - wanted| <= tol), 1 Phi out of range,
 *                     2 singular Jacobian (bifurcation point), 3 max_iter sweeps spent, 4 stalled: the line search
 *                     could not lower the residual (a fold of the base-angle map — no root on this branch)
 * These robots are unstable for a large share of random inputs (calcStability) and the base-angle map then has
 * several roots.  Without a guess the solve starts from x = beta and, when a start stalls, starts again from
 * reproducible pseudo-random angles until a root is found or max_iter sweeps are spent (so statuses 2 and 4 only come
 * back together with a guess): on the shipped robots 99.3-100 % of random inputs with a root are solved within 100
 * sweeps at a mean of 6-11 (a cap of 100 is a good default; the work queue absorbs the tail) — but WHICH root comes
 * back is arbitrary.  A robot driven at the base follows a branch: pass the previous configuration's solution as the
 * guess (ctr_fk_batch_base_ex); the solve then never leaves that branch and reports STALLED where it ends.
 * DEVICE pointers. */
#define CTR_FK_SHOOT_OK        0
#define CTR_FK_SHOOT_PHI_RANGE 1
#define CTR_FK_SHOOT_SINGULAR  2
#define CTR_FK_SHOOT_MAX_ITER  3
#define CTR_FK_SHOOT_STALLED   4
int ctr_fk_batch_base(ctr_fk_handle h, const double* jv_base, int64_t B, int mode, double step,
                      int32_t nsamples, double tol, int32_t max_iter, double* jv_tip, double* tip,
                      int32_t* iters, int32_t* status, void* stream);
/* The same with
 *   jv_guess [B][n_jv]  in, nullable: the alpha slots hold the starting tip angles (warm start / branch tracking);
 *                       NULL starts from x = beta
 *   resid    [B]        out, nullable: max |base - wanted| at the returned tip angles
 *   flags               CTR_FK_FLAG_RK4: shoot on the CONTINUUM rod model integrated tip -> base by
 *                       classical fixed-step RK4 with the variational equations carried through every stage, instead
 *                       of on the reference's first-order sweep.  A different discretisation: its roots differ from
 *                       the reference sweep's by O(h) (halving with h); `tip` is still the reference FK at the tip
 *                       angles found.  Checked against oracle/ctr_shoot_rk4_py.py, not against the reference. */
#define CTR_FK_FLAG_RK4 0x1000u
int ctr_fk_batch_base_ex(ctr_fk_handle h, const double* jv_base, const double* jv_guess, int64_t B, int mode,
                         double step, int32_t nsamples, double tol, int32_t max_iter, uint32_t flags,
                         double* jv_tip, double* tip, int32_t* iters, int32_t* status, double* resid,
                         void* stream);

/* ---- inputs ---------------------------------------------------------------------------- */

/* Fill jv[B][n_jv] (DEVICE pointer) with the distribution of getRandomJointValues
 * (robot_i.h:1308-1317): theta, alpha = nextafter(2*pi,0)*u, phi_s = Length_s*u, u~U[0,1).
 * u for (config i, slot j) is splitmix64-derived from (seed, first_config + i, j) so any shard
 * of a batch can be generated independently and reproducibly. */
int ctr_fk_sample_joints(ctr_fk_handle h, uint64_t seed, int64_t first_config, int64_t B,
                         double* jv, void* stream);
/* The same generator on the host (no CUDA). */
int ctr_fk_sample_joints_host(const ctr_robot_desc* desc, uint64_t seed, int64_t first_config,
                              int64_t B, double* jv);

/* Sampling with the extension-length scale policies of transrotscale.h
 * (getRandomJointValues(.., TransRotScale), section_i.h:338-347):
 *   CTR_FK_SCALE_NONE          phi_s = Length_s * u
 *   CTR_FK_SCALE_GAMMA         phi_s = Length_s * u^p0 (GammaCorrection, :106-122; lut_size > 1
 *                              selects its LookUpSize-entry interpolated table)
 *   CTR_FK_SCALE_PROPORTIONAL  ProportionalScale (:273-291): total extension steered towards
 *                              MaxLength * U[p0, p1]
 * jv is a DEVICE pointer for ctr_fk_sample_joints_scaled, a host pointer for the _host form. */
#define CTR_FK_SCALE_NONE         0
#define CTR_FK_SCALE_GAMMA        1
#define CTR_FK_SCALE_PROPORTIONAL 2
int ctr_fk_sample_joints_scaled(ctr_fk_handle h, uint64_t seed, int64_t first_config, int64_t B,
                                int policy, double p0, double p1, int32_t lut_size, double* jv,
                                void* stream);
int ctr_fk_sample_joints_scaled_host(const ctr_robot_desc* desc, uint64_t seed, int64_t first_config,
                                     int64_t B, int policy, double p0, double p1, int32_t lut_size,
                                     double* jv);

/* ---- measurement helpers ----------------------------------------------------------------- */

/* Sample count (m_Samples before the position-driven forward loop), host side, bit-identical to
 * the kernel: used by bench.py to compute the algorithmic flop count of a step-mode batch. */
int ctr_fk_count_samples_host(const ctr_robot_desc* desc, const double* jv, int64_t B, int mode,
                              double step, int32_t nsamples, int32_t* n);
/* FP64 FMA-pipe peak probe: runs independent DFMA chains on every SM for `iters` iterations on
 * `stream` and returns the number of flops issued (2 per DFMA per lane).  Time it with events. */
int ctr_fk_fp64_probe(int device, int32_t iters, double* flops_issued, void* stream);
/* The same for the FP32 FMA pipe in its packed form (fma.rn.f32x2, SASS FFMA2 — what the FP32 variant's
 * solver issues): 4 flops per instruction per lane. */
int ctr_fk_fp32_probe(int device, int32_t iters, double* flops_issued, void* stream);
/* Number of kernel launches issued by this library since load (all threads). */
int64_t ctr_fk_launch_count(void);

const char* ctr_fk_last_error(void);
const char* ctr_fk_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CTR_FK_H */
