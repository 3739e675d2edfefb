/*
 * rbx.h -- C ABI of the B200-native rigid-body hot path (librbx.so).
 *
 * This is the drop-in boundary for the one hot path of
 * dineshadepu/rigid_body_2d_3d_pysph: one time step of the rigid-body scheme
 * (SURVEY.md section 3.3, 8a).  The reference has no FFI of its own -- PySPH
 * transpiles the Python Equation / IntegratorStep methods to Cython at run
 * time -- so each entry point below names the reference method(s) whose
 * generated loop it replaces (paths relative to /root/reference/code/).
 *
 * Conventions
 *  - Every pointer inside the descriptor structs is a DEVICE pointer owned by
 *    the caller (torch tensors on the Python side).  The library allocates
 *    nothing persistent and keeps no global state.
 *  - All calls are asynchronous on the given stream (a cudaStream_t passed as
 *    void*), capture-safe (no syncs, no allocations), and re-entrant across
 *    streams and devices.
 *  - Return value: RBX_OK or a negative RbxError; no C++ exception crosses
 *    the ABI.  Device-side capacity problems (slot table full, neighbour
 *    list full, grid too large) are OR-ed into the device status word
 *    RbxScene.status, which the host reads at a sync point of its choosing.
 *  - Real data are FP64, ids int32, exactly as PySPH's `double` / `int`.
 */
#ifndef RBX_H_
#define RBX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBX_VERSION 404 /* 0.4.4 */

typedef enum {
  RBX_OK = 0,
  RBX_ERR_INVALID = -1,   /* bad argument / null pointer / size */
  RBX_ERR_WORKSPACE = -2, /* workspace too small */
  RBX_ERR_LAUNCH = -3,    /* CUDA launch error (cudaGetLastError) */
  RBX_ERR_NO_DEVICE = -4
} RbxError;

/* bits of the device status word */
#define RBX_STATUS_SLOT_OVERFLOW 1u  /* > 32 source-body runs in one particle's
                                        list (or the diagnostics hold >
                                        RBX_MAX_KEYS)                         */
#define RBX_STATUS_HIST_OVERFLOW 2u  /* > ks simultaneous contacts            */
#define RBX_STATUS_LIST_OVERFLOW 4u  /* per-particle neighbour list full       */
#define RBX_STATUS_GRID_COARSENED 8u /* cell size enlarged to fit cap_cells    */
#define RBX_STATUS_LVC_OVERFLOW 16u  /* LVC tangential history `limit` hit     */
#define RBX_STATUS_PAIR_OVERFLOW 32u /* RbxDiag.pairs buffer full              */

#define RBX_MAX_KEYS 8   /* slots per particle the optional RbxDiag arrays hold */

/* Uniform grid description, written on the device by rbx_cells_build. */
typedef struct {
  double x0, y0, z0;  /* origin = min corner of the binned points */
  double cell;        /* cell edge >= radius_scale * h_max        */
  double inv_cell;
  int32_t nx, ny, nz;
  int32_t ncells;     /* nx*ny*nz <= cap_cells                    */
  int32_t npoints;    /* points binned                            */
  int32_t pad_;
} RbxGridInfo;

/* A set of points to bin: the particles `index[k]` (k < n) of the scene's
 * global SoA arrays, or particles 0..n-1 when index is NULL. */
typedef struct {
  int32_t n;
  int32_t pad_;
  const int32_t *index;
  const double *x, *y, *z, *h;
  const int32_t *dem_id;
} RbxPoints;

/* Cell list, caller-allocated.  After rbx_cells_build:
 *   cell_start[c] .. cell_start[c+1]  = sorted range of cell c
 *   (c = (cz*ny + cy)*nx + cx, x fastest), ascending global index inside a
 *   cell (deterministic);  gidx/sx/sy/sz/sh/sdem = sorted SoA copies.      */
typedef struct {
  int32_t cap_cells;   /* cell_start has cap_cells + 1 entries */
  int32_t cap_points;
  RbxGridInfo *info;   /* device */
  int32_t *cell_start; /* [cap_cells + 1] */
  int32_t *cell_of;    /* [cap_points] scratch: cell id per binned point */
  int32_t *rank;       /* [cap_points] scratch: arrival rank inside cell */
  int32_t *gidx;       /* [cap_points] sorted -> global particle index   */
  double *sx, *sy, *sz, *sh; /* [cap_points] sorted copies              */
  int32_t *sdem;       /* [cap_points] */
  const uint32_t *cond; /* device word or NULL: when it reads 0 the build is
                           skipped (neighbour lists still valid, see
                           RbxScene.rebuild)                              */
} RbxCells;

/* The scene: global SoA over all particles (rigid-body particles first,
 * static boundary particles after), per-body state in the reference's own
 * layout (3*b+j, 9*b+j), and the sparse contact history.                  */
typedef struct {
  int32_t n_total;   /* all particles                                     */
  int32_t n_rigid;   /* destination (rigid-body) particles: [0, n_rigid)  */
  int32_t n_bodies;  /* rigid bodies (global numbering across arrays)     */
  int32_t n_chunks;  /* work items: <= 128 consecutive particles of 1 body */
  int32_t dim;       /* kernel dimension (2 or 3)                          */
  int32_t ks;        /* history slots per particle                         */
  int32_t eta_mode;  /* 0: none, 1: dense table rows, 2: uniform scalar    */
  int32_t planar;    /* 1: GTVFRigidBody2DStep semantics                   */
  int32_t list_cap;  /* neighbour-list entries per particle                */
  int32_t pad_;
  /* per particle [n_total] */
  double *x, *y, *z, *u, *v, *w;
  const double *h, *m, *rho;
  const int32_t *dem_id;
  /* per rigid particle [n_rigid] */
  double *fx, *fy, *fz;
  const double *dx0, *dy0, *dz0;
  const int32_t *body;        /* global body index                       */
  const int32_t *is_boundary; /* may be NULL: no normal rotation         */
  const double *normal0;      /* stride 3, may be NULL                   */
  double *normal;             /* stride 3, may be NULL                   */
  /* work items */
  const int32_t *chunk_start; /* [n_chunks + 1] particle ranges          */
  const int32_t *chunk_body;  /* [n_chunks]                              */
  const int32_t *body_chunk;  /* [n_bodies + 1] chunk ranges per body    */
  /* Neighbour lists.  nbr_pos[list_cap][n_rigid] (column = particle): global
   * index of every gated source within reach + skin of the particle when the
   * list was built, the sources of one body contiguous, bit 31 set on the
   * first entry of a body; nbr_cnt[n_rigid] entries per particle (bit 30:
   * a body may own more than one run of this list).  Built by
   * rbx_contact_neighbours when *rebuild != 0, reused otherwise.
   * rbx_contact_slots reads the transposed copy made at the same time:
   * work item t <-> particle nbr_order[t], the particles of every window of
   * 256 ordered by descending list length (so that the lanes of a warp run
   * lists of equal length); nbr_cnt_srt[t] entries in column t of
   * nbr_srt[list_cap][n_rigid], bit 31 set on the LAST entry of a body.  It
   * applies the exact neighbour predicate to every entry with the current
   * positions, so the pair set is independent of the skin.               */
  int32_t *nbr_pos, *nbr_cnt;
  int32_t *nbr_order, *nbr_cnt_srt, *nbr_srt;
  /* per body */
  const double *total_mass, *izz, *spacing0; /* [n_bodies]               */
  double *xcm, *vcm, *ang_mom, *omega;       /* [3 n_bodies]             */
  double *force, *torque;                    /* [3 n_bodies]             */
  double *R, *R_prev;                        /* [9 n_bodies]             */
  const double *iinv_b;                      /* [9 n_bodies]             */
  double *iinv_g;                            /* [9 n_bodies]             */
  double *xcm0, *vcm0, *ang_mom0, *R0;       /* RK2 saved state          */
  /* damping table (rigid_body_common.py:925): eta_mode 1 -> value
   * eta[eta_row[body] + source_dem_id]; eta_mode 2 -> eta[0]            */
  const double *eta;
  const int64_t *eta_row;
  /* sparse history, slot-major [ks][n_rigid]; key = source dem_id, -1 end */
  const int32_t *hist_key_in;
  const double *hist_dlt_in, *hist_fn_in; /* [3][ks][n_rigid] */
  int32_t *hist_key_out;
  double *hist_dlt_out, *hist_fn_out;
  /* device status word + 8 counters: [0] gated in-range pairs, [1] active
   * (in-contact) slots, [2] candidate distance tests, [3] neighbour-list
   * entries written, [7] chunk dispenser of k_neighbours (internal)      */
  uint32_t *status;
  unsigned long long *counters;
  /* list reuse: rebuild[0] != 0 <=> lists must be rebuilt this evaluation.
   * Set by the drift / RK2 kernels when a body has moved more than skin/2
   * since the last build (|xcm - xcm_ref| + |R - R_ref|_F * rmax), by the
   * host after it changed positions; cleared after a rebuild.             */
  uint32_t *rebuild;
  double *xcm_ref, *R_ref;    /* [3 n_bodies], [9 n_bodies] at last build */
  const double *rmax;         /* [n_bodies] max |body-frame position|     */
  /* [n_bodies] first body (global index) of the particle array the body
   * belongs to, or NULL = one array.  RK2RigidBody3DStep.py_initialize saves
   * the angular momentum of the FIRST body of every array only
   * (rigid_body_3d.py:415, quirk Q7).                                      */
  const int32_t *body_first;
  /* Two-precision contact evaluation (optional: both NULL -> every slot is
   * evaluated in FP64).  pos32[n_total] = {x - origin[0], y - origin[1],
   * z - origin[2], h} as floats, kept current by rbx_pose_particles (rigid
   * particles), rbx_halo_unpack, rbx_pos32_refresh, and for the static
   * particles by rbx_contact_neighbours on a rebuild.  A first pass sums
   * every (particle, source body) slot in FP32 with a running error bound
   * and proves most of them out of contact; clist[4 n_rigid] receives
   * {work item, bit mask of the runs that could not be excluded, first entry
   * | ordinal << 20 of the first such run, last entry of the last one} and
   * only those are evaluated by the exact FP64 code.  counters[6] = entries. */
  float *pos32;
  int32_t *clist;
  double origin[3];
  /* Sparse outputs of the contact evaluation (all optional, NULL = every
   * output is written for every particle at every evaluation).  At config 5
   * about 1 % of the particles are in contact; everything else carries
   * fx, fy, fz = m g and an empty history, step after step.  The evaluation
   * then only writes what changes:
   *   alist_out / acount_out: particles that came out of THIS evaluation with
   *     at least one slot in contact (they own rows of hist_*_out and a
   *     non-trivial force); alist_prev / acount_prev: the same of the previous
   *     evaluation.  Before the pair kernels run, fx, fy, fz of alist_prev are
   *     reset to m g and hist_key_out of the old content of alist_out (two
   *     evaluations ago: the last owner of these rows) to "empty".  The
   *     caller swaps the two together with the history buffers.
   *   RBX_PARAM_DENSE_OUT makes one evaluation write every particle (first
   *     evaluation, or after the host changed fx / m / the history).
   *   body_tag[n_bodies]: set to 1 for a body with a particle in contact;
   *     rbx_reduce_bodies sums the particles of tagged bodies only (and clears
   *     (the next evaluation clears the tags of alist_prev), every other body
   *     gets force = total_mass * gravity, torque = 0.
   *   aux32[2 n_rigid] = {m / rho, spacing0 of the body} as floats for the
   *     FP32 first pass (instead of m, rho, body, spacing0[body]).
   *   h_uniform > 0: every particle has this h (pos32.w without loading h).  */
  int32_t *alist_out, *alist_prev;
  uint32_t *acount_out, *acount_prev;
  int32_t *body_tag;
  const float *aux32;
  double h_uniform;
  double gravity[3];  /* = RbxParams.gx, gy, gz (force of an untagged body) */
  /* [3 (n_total - n_rigid)] x, y, z of the static particles (walls, halo) when
   * the neighbour lists were last built, or NULL.  rbx_static_update and
   * rbx_halo_unpack raise `rebuild` when one of them has moved more than
   * skin / 2 since -- the rule the drift kernel applies to the bodies.       */
  double *static_ref;
  /* A second stream of the caller's (a cudaStream_t), or NULL.  When
   * rbx_gtvf_step is called while `stream` is being captured into a CUDA
   * graph, the list rebuild (cell list + neighbour lists, ~13 kernels that
   * otherwise all launch and return at once on the 9 steps of 10 that reuse
   * the lists) is captured through it into the body of a conditional IF node
   * on *rebuild: a step that reuses the lists then does not launch them.    */
  void *aux_stream;
} RbxScene;

typedef struct {
  double radius_scale; /* 3.0 for QuinticSpline */
  double kr, kf, fric_coeff;
  double gx, gy, gz;
  double dt;
  double reach;     /* radius_scale * h_max over ALL arrays = minimum cell edge */
  double h_uniform; /* > 0: every particle has this h (skips the h loads)  */
  double skin;      /* >= 0: extra radius of the neighbour lists; 0 rebuilds
                       them at every evaluation                            */
  int32_t flags;    /* RBX_PARAM_* */
  int32_t pad_;
} RbxParams;

#define RBX_PARAM_EXACT 1  /* evaluate every slot in FP64 (no FP32 first pass) */
#define RBX_PARAM_BODY_VEL 2 /* velocities of rigid particles are not read from
                                u, v, w but formed where needed as
                                vcm + omega x (R_prev r0) (stage 1 of the GTVF
                                step, rigid_body_3d.py:62-95, fused after the
                                drift): rbx_gtvf_step sets it and writes u, v,
                                w once, at the end of the step              */
#define RBX_PARAM_DENSE_OUT 4 /* see RbxScene.alist_out                       */

/* Optional per-slot diagnostics of one contact evaluation, slot-major
 * [RBX_MAX_KEYS][n_rigid]; any pointer may be NULL.  Used by parity tests
 * to rebuild the reference's dense (particle, body) slot arrays.          */
typedef struct {
  int32_t *key;     /* source dem_id, -1 = unused                        */
  int32_t *closest; /* global index of the closest source particle       */
  double *nx, *ny, *nz, *dist, *overlap, *ftx, *fty, *ftz;
  /* neighbour pairs as the contact kernel sees them: every list entry that
   * passes the gate of rigid_body_common.py:678-679 and the exact NNPS
   * predicate is appended as {destination, source} (global indices, in no
   * particular order) to pairs[2 pair_cap]; pair_count[0] = pairs found
   * (may exceed pair_cap: RBX_STATUS_PAIR_OVERFLOW).  NULL: not wanted.   */
  int32_t *pairs;
  unsigned long long *pair_count;
  int64_t pair_cap;
} RbxDiag;

/* The DEMScheme scene (dem.py): granular (destination) particles first,
 * boundary particles after.  History entries are the reference's own:
 * tng_idx = index of the source particle INSIDE ITS ARRAY, tng_dem = its
 * dem_id, [n_dest * limit], -1 = empty.                                   */
typedef struct {
  int32_t n_total, n_dest, n_arrays, limit;
  double *x, *y, *z, *u, *v, *w, *wx, *wy, *wz;   /* [n_total] */
  const double *h, *m, *rad_s;                    /* [n_total] */
  const double *moi;                              /* [n_dest]  */
  const int32_t *dem_id;                          /* [n_total] */
  const int32_t *arr_off;    /* [n_total] first global index of its array */
  const int32_t *arr_start;  /* [n_arrays + 1] array ranges, scheme order */
  const int32_t *tbl_row;    /* [n_dest] row offset into kn/kt/alpha/mu   */
  const double *kn, *kt, *alpha, *mu;  /* by row + source dem_id          */
  double *fx, *fy, *fz, *torx, *tory, *torz;      /* [n_dest] */
  int32_t *tng_idx, *tng_dem;                     /* [n_dest * limit] */
  double *tng_x, *tng_y, *tng_z;                  /* [n_dest * limit] */
  int32_t *total_tng;                             /* [n_dest] */
  uint32_t *status;
} RbxDemScene;

/* Inputs of the Canelas Hertz contact that RbxScene does not carry, all
 * [n_total]: sphere radius, Young modulus and Poisson ratio of every particle
 * (the reference reads the array constants d_E[0], s_E[0], ...), and the
 * global body index of the particle (-1: not part of a rigid body -> the
 * RigidWall law).                                                         */
typedef struct {
  const double *rad_s, *E, *nu;
  const int32_t *src_body;
  double Cn;         /* damping constant, 1.4e-5 in the reference */
} RbxCanelas;

int rbx_version(void);
const char *rbx_strerror(int code);
/* sizeof(RbxGridInfo, RbxPoints, RbxCells, RbxScene, RbxParams, RbxDiag) for
 * which = 0..5, RbxDemScene for 6, RbxCanelas for 7: lets a binding check its
 * struct mirrors.                                                         */
size_t rbx_sizeof(int which);

/* Scratch bytes needed by rbx_cells_build for a cell list of this capacity. */
size_t rbx_cells_workspace_bytes(int32_t cap_cells, int32_t cap_points);

/* [upstream] NNPS.update (LinkedListNNPS binning).
 * Counting-sort cell-list build: bounds -> cell ids + histogram -> prefix
 * scan -> scatter -> per-cell index sort -> gather of the sorted SoA.
 * min_cell = radius_scale * h_max over all arrays.                       */
int rbx_cells_build(const RbxPoints *pts, const RbxCells *cells,
                    double min_cell, uint32_t *status, void *workspace,
                    size_t workspace_bytes, void *stream);
/* (min_cell = RbxParams.reach) */

/* [upstream] NNPS.get_nearest_particles: for every point of `dst`, the
 * source particles j of the cell list with r2 < (k h_i)^2 or r2 < (k h_j)^2.
 * Parity mode.  With idx == NULL only counts[i] is written; otherwise
 * offsets[i] (exclusive scan of counts, made by the caller) places row i.  */
int rbx_pairs_dump(const RbxPoints *dst, const RbxCells *cells,
                   double radius_scale, int32_t *counts,
                   const int64_t *offsets, int32_t *idx, void *stream);

/* ComputeContactForceNormals + ComputeContactForceDistanceAndClosestPoint +
 * BodyForce + ComputeContactForce (rigid_body_common.py:631-1032, :115-125),
 * fused, with sparse slots.  Reads history *_in, writes *_out and fx,fy,fz. */
int rbx_contact_mofidi(const RbxScene *scene, const RbxCells *cells,
                       const RbxParams *params, const RbxDiag *diag,
                       void *stream);

/* The two launches of rbx_contact_mofidi, separately (profiling, tests):
 * neighbours = exact NNPS predicate + source gate (rigid_body_common.py:
 * 678-679) -> neighbour lists; slots = pair sums, force law, partials.     */
int rbx_contact_neighbours(const RbxScene *scene, const RbxCells *cells,
                           const RbxParams *params, void *stream);
int rbx_contact_slots(const RbxScene *scene, const RbxCells *cells,
                      const RbxParams *params, const RbxDiag *diag,
                      void *stream);

/* BodyForce.initialize + RigidBodyCanelasRigidRigid.loop +
 * RigidBodyCanelasRigidWall.loop (rigid_body_common.py:115-125, 244-628;
 * tangential part disabled upstream) over the cell list `cells` of the
 * source particles: fx, fy, fz of every rigid particle.                     */
int rbx_contact_canelas(const RbxScene *scene, const RbxCells *cells,
                        const RbxParams *params, const RbxCanelas *canelas,
                        void *stream);

/* SumUpExternalForces.reduce (rigid_body_common.py:128-175): fx,fy,fz,x,y,z
 * -> force[3nb], torque[3nb], one warp per body, fixed order
 * (deterministic).  With RbxScene.body_tag: tagged bodies only (see there). */
int rbx_reduce_bodies(const RbxScene *scene, void *stream);

/* GTVFRigidBody{3D,2D}Step.py_stage1 / py_stage3 (rigid_body_3d.py:41-60,
 * 171-190; rigid_body_2d.py:41-55, 157-170): half kick of vcm, ang_mom,
 * omega.                                                                    */
int rbx_gtvf_kick(const RbxScene *scene, double dt, void *stream);

/* GTVFRigidBody{3D,2D}Step.py_stage2 + normalize_R_orientation
 * (rigid_body_3d.py:97-132, rigid_body_common.py:178-203): drift of xcm, R
 * (R_prev keeps the pre-drift orientation), inertia update.  Also raises
 * RbxScene.rebuild when a body has moved more than skin/2 since the
 * neighbour lists were built.                                              */
int rbx_gtvf_drift(const RbxScene *scene, double dt, double skin,
                   void *stream);

/* stage1/stage3 (velocities, rigid_body_3d.py:62-95, 192-225) and stage2
 * (positions + boundary normals, :134-169) of the body particles.
 * flags: RBX_POSE_* below.                                                 */
#define RBX_POSE_POS 1       /* x,y,z = xcm + R r0                       */
#define RBX_POSE_VEL 2       /* u,v,w = vcm + omega x (R r0)             */
#define RBX_POSE_VEL_PREV 4  /* velocities use R_prev (stage1 fused after drift) */
#define RBX_POSE_NORMALS 8   /* rotate normal0 -> normal where is_boundary */
int rbx_pose_particles(const RbxScene *scene, int flags, void *stream);

/* pos32 of the particles [first, first + n) from x, y, z, h (after the caller
 * wrote positions itself; no counterpart in the reference).                 */
int rbx_pos32_refresh(const RbxScene *scene, int32_t first, int32_t n,
                      void *stream);

/* Host-driven boundaries (an Application's post_step moving a wall,
 * stack_of_cylinders.py:438-445): new state of the static particles
 * [first, first + n) from device buffers (any of x..w may be NULL = keep),
 * pos32 refreshed, `rebuild` raised if a particle is now further than
 * skin / 2 from where it was when the lists were built (static_ref).        */
int rbx_static_update(const RbxScene *scene, int32_t first, int32_t n,
                      const double *x, const double *y, const double *z,
                      const double *u, const double *v, const double *w,
                      double skin, void *stream);

/* Multi-GPU halo payload (no counterpart in the reference, which is single
 * process): rows of 8 doubles {x, y, z, u, v, w, h, dem_id}.  pack gathers
 * the particles index[0..n) into rows[n][8] (body_vel != 0: the velocity of
 * a body particle is its stage-1 velocity formed from the body state, as
 * under RBX_PARAM_BODY_VEL); unpack writes rows[n][8] into the particles
 * [first, first + n) (a static source array of the scene) and raises
 * `rebuild` if one of them has moved more than skin / 2 since the lists
 * were built.                                                              */
int rbx_halo_pack(const RbxScene *scene, const int64_t *index, int32_t n,
                  double *rows, int body_vel, void *stream);
int rbx_halo_unpack(const RbxScene *scene, int32_t first, int32_t n,
                    const double *rows, double skin, void *stream);

/* RK2RigidBody3DStep (rigid_body_3d.py:406-575): stage 0 = py_initialize
 * (fix_q7 != 0 saves ang_mom0 of every body), 1 and 2 =
 * py_stage1 / py_stage2.  Particle update: rbx_pose_particles(POS|VEL).    */
int rbx_rk2_stage(const RbxScene *scene, int stage, double dt, int fix_q7,
                  double skin, void *stream);

/* UpdateTangentialContactsLVCDisplacement.initialize_pair (dem.py:208-293),
 * BodyForce.initialize, LVCDisplacement.loop (dem.py:35-205) over a cell list
 * of ALL particles.  torx/tory/torz accumulate across calls, as in the
 * reference (nothing resets them).                                          */
int rbx_contact_lvc(const RbxDemScene *scene, const RbxCells *cells,
                    const RbxParams *params, void *stream);

/* DEMStep.stage1/2/3 (dem.py:595-625); stage in {1, 2, 3}.                  */
int rbx_dem_step(const RbxDemScene *scene, int stage, double dt,
                 void *stream);

/* Setup path (one-off): boundary-particle identification of ONE array whose
 * particles are all binned in `cells`: ComputeNormals -> SmoothNormals
 * [upstream wall_normal; boundary_particles.py:71-135] ->
 * IdentifyBoundaryParticleCosAngle (boundary_particles.py:22-68).  m, rho,
 * normal_tmp[3n], normal[3n], is_boundary[n] are indexed like the points.   */
int rbx_boundary_identify(const RbxPoints *pts, const RbxCells *cells, int dim,
                          double radius_scale, const double *m,
                          const double *rho, double *normal_tmp,
                          double *normal, int32_t *is_boundary, void *stream);

/* Setup path (one-off): set_total_mass, set_center_of_mass,
 * set_moment_of_inertia_izz, set_moment_of_inertia_and_its_inverse and
 * set_body_frame_position_vectors (rigid_body_common.py:21-107) of one rigid
 * array whose particles are grouped by body: body b owns the particles
 * [body_start[b], body_start[b + 1]).  Writes total_mass[nb], xcm[3 nb],
 * izz[nb] (may be NULL), the inertia tensor about the centre of mass and its
 * inverse [9 nb] (body frame = global frame at t = 0), dx0, dy0, dz0[n].    */
int rbx_setup_bodies(int32_t n_bodies, const int32_t *body_start,
                     const double *x, const double *y, const double *z,
                     const double *m, double *total_mass, double *xcm,
                     double *izz, double *inertia, double *inertia_inverse,
                     double *dx0, double *dy0, double *dz0, void *stream);

/* Whole GTVF step [upstream GTVFIntegrator.one_timestep, SURVEY App. C-6]:
 * kick, drift, pose, cells_build, contact, reduce, kick, velocities.
 * `src` = the source points (contact_force_is_boundary == 1) to bin.
 * flags: bit0 = skip the final particle-velocity / boundary-normal write
 * (inside a batch of steps nothing reads them: stage 1 of the next step
 * overwrites the velocities, and the normals are a function of R alone);
 * bit1 = only the part before the force evaluation (kick, drift, positions),
 * bit2 = only the part from the cell list on -- a multi-GPU caller puts its
 * halo exchange between the two                                             */
int rbx_gtvf_step(const RbxScene *scene, const RbxPoints *src,
                  const RbxCells *cells, const RbxParams *params,
                  void *workspace, size_t workspace_bytes, int flags,
                  void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RBX_H_ */
