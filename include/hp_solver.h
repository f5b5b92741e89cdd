/* hp_solver.h -- C ABI of libhpsolver (B200 / sm_100a) : the transient r-z conduction hot path of
 * kubaniak/heat-pipe-solver.
 *
 * The reference has no FFI: its hot path is the FiPy call `eq.sweep(var=T, dt=dt)` inside a module-level
 * Python loop (reference src/2d/main.py:148-150, 200-202, 267-271).  This library is what a replacement
 * of that call binds to (ctypes from Python; the binding a maintainer would add is in INTEGRATION.md).
 * Each entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.  `stream` arguments are `cudaStream_t` passed
 *     as `void*` (0 = legacy default stream).
 *   - one host thread per handle; one handle per GPU (the current CUDA device at hp_create time).
 *   - all work is enqueued on the handle's stream; calls return without synchronising unless stated.
 *   - return value 0 = success, negative = error (text via hp_last_error()).  Nothing throws; the
 *     library never frees caller memory; the handle owns all scratch.
 *   - fields are fp64, FiPy cell order: id = j*nr + i, radial index i fastest (CylindricalGrid2D with
 *     dr->x, dz->y; reference src/2d/mesh.py:109, src/2d/main.py:30-31).
 *   - "ghost-inclusive" per-row arrays have nz+2 entries: entry 0 / nz+1 describe the row below / above
 *     the owned block (a neighbouring slab's row, or a dummy at a physical end).  Field pointers handed
 *     to hp_set_T / hp_get_T cover the nz owned rows only.
 */
#ifndef HP_SOLVER_H
#define HP_SOLVER_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hp_solver_s* hp_handle;

/* physical constants + mode switches (reference config/faghri.json, main.py:37-46,103-150) */
typedef struct hp_phys {
    double R_vc, M_g, R_gas, porosity, T_melt, dT_melt, sigma;
    double simple_rho, simple_cp, simple_k;      /* main.py:41-46 */
    int32_t properties;   /* 0 temperature_dependent (main.py:103-117), 1 simple */
    int32_t gamma;        /* 0 snapshot at create time (main.py:133,145), 1 lazy */
    int32_t source;       /* 0 q/k (main.py:150), 1 q */
    int32_t face_k;       /* 0 region law at arithmetic face T (main.py:103-106), 1 harmonic mean of cell k */
    int32_t radiation;    /* 0 off, 1 on: g = h*T_amb + sigma*eps*(T_amb^4 - T_f^4) (verification_examples.py:81-91) */
    int32_t _pad;
} hp_phys;

/* structured r-z grid + region codes.  Replaces the FiPy mesh object built by
 * mesh.generate_composite_mesh (reference src/2d/mesh.py:42-149) and the position masks of
 * main.py:64-72,86-90,123-124, which the host evaluates with the reference's float recipe. All HOST pointers. */
typedef struct hp_mesh_desc {
    int32_t nr;                 /* radial cells per row (fast index) */
    int32_t nz;                 /* owned axial rows */
    int32_t cartesian;          /* 0: CylindricalGrid2D (mesh.py:109), per-radian areas/volumes; 1: Grid2D (mesh.py:12-32) */
    int32_t _pad;
    const double* r_v;          /* [nr+1]   radial vertices */
    const double* z_v;          /* [nz+3]   axial vertices, ghost-inclusive: row g spans z_v[g]..z_v[g+1] */
    const int8_t* cell_band;    /* [nr]     0 vc, 1 wick, 2 wall, 3 none     (cells, main.py:64-70) */
    const int8_t* rface_band;   /* [nr]     band of the radial face at r_v[i+1]; i=nr-1 is the outer surface (main.py:86-90) */
    const int8_t* zface_band;   /* [nr]     band of axial faces (centre r_c[i], face inequalities) */
    const uint8_t* zc_flags;    /* [nz+2]   HP_ZF_* bits at row centres */
    const uint8_t* zv_flags;    /* [nz+2]   HP_ZF_* bits at the face between row g and g+1 */
    const uint8_t* axial_open;  /* [nz+2]   1 if the face between row g and g+1 conducts (interior face) */
    /* outer-surface boundary data per row (main.py:36-37,130; ensembles vary them per member) */
    const double* q_line;       /* [nz+2]   evaporator heat flux  W/m^2 */
    const double* h_line;       /* [nz+2]   condenser film coefficient W/m^2K */
    const double* Tamb_line;    /* [nz+2]   ambient temperature K (also the snapshot temperature) */
    const double* eps_line;     /* [nz+2]   emissivity (radiation mode) */
} hp_mesh_desc;

typedef struct hp_solve_opts {
    int32_t precond;        /* 0 Jacobi, 1 radial-line tridiagonal, 2 'mgline': V(1,1) multigrid, axial semi-coarsening,
                               damped radial-line smoother (the stopping rule stays the line-preconditioned residual) */
    int32_t criterion;      /* 0 max|M^-1 r| <= tol [K]; 1 max|r_i/a_ii| <= tol [K]; 2 ||r||_2 <= tol*||b||_2 */
    double tol;
    int32_t max_iter;
    int32_t loop_mode;      /* 0 host-polled chunks of poll_chunk iterations; 1 CUDA graph, device-side WHILE */
    int32_t poll_chunk;
    int32_t refactor_every; /* re-factor the line preconditioner every k-th sweep (1 = every sweep) */
    int32_t extrapolate;    /* 0 off; k = 1..3: hp_step/hp_run start the first solve of a step from the field extrapolated from
                               the last k+1 step values (1: T + (T - T_prev); 2 / 3: quadratic / cubic in the step index, used
                               once enough steps of history exist).  The system is still assembled at T and `res` is still
                               FiPy's pre-solve residual; only the PCG start changes */
    int32_t mg_levels;      /* mgline: maximum number of levels (>= 2) */
    int32_t mg_coarse_sweeps; /* mgline: extra smoothing sweeps on the coarsest level */
    int32_t reuse_converged;  /* 1: a sweep that exactly repeats a sweep which converged before its first iteration (T, T_old, dt
                                 untouched since) only redoes the bookkeeping -- same residual / statistics, no assembly.
                                 Single GPU only.  0 (default): every sweep assembles, like the reference */
    double mg_omega;        /* mgline: damping of the line smoother (default 0.6) */
    int32_t while_unroll;   /* loop_mode 1: PCG iterations per pass of the device-side WHILE node (1..8, default 1) */
    int32_t peer_timeout_ms;/* peer-memory slabs: how long a kernel waits for a neighbour's flag before it gives up and raises
                               the handle's fault (default 60000).  A fault is sticky: every later call on the handle returns an
                               error until hp_comm_clear_fault */
    double mg_zscale;       /* mgline: axial coupling of a coarse operator relative to the Galerkin (row-pair aggregation) value;
                               1 = Galerkin, 0.5 (default) = what a discretisation on the coarse rows gives */
    int32_t tma_rows;       /* 1: k_cg_B stages its operand rows in shared memory with bulk asynchronous copies (cp.async.bulk +
                               mbarrier, the next rows in flight during the scans of the current one) where the row shape allows
                               (512 < nr <= 1024, nr % 4 == 0); 0: register-direct 256-bit loads */
    int32_t pipe_kernel;    /* 1 (default): meshes made of small uncoupled pipes (9 radial cells, <= 512 rows each: the native
                               Faghri mesh, ensembles of it) with precond = 1 run each pipe's whole time step in ONE CTA, on
                               chip (every pipe its own PCG to `tol`); 0: always the row kernels */
    int32_t _pad;
    double ext_guard_tols;  /* extrapolate = 3: the cubic term of the predicted start (third difference of the step history) is
                               applied cell by cell with a weight rising from 0 at ext_guard_tols * tol to 1 at twice that --
                               below the floor it is solver noise (default 300; 0 = always apply in full) */
} hp_solve_opts;

typedef struct hp_sweep_stat {
    double residual_l2;     /* FiPy's `res`: ||L(T) T - b(T)||_2 before the solve (main.py:202) */
    double rhs_l2;          /* ||b||_2 */
    double crit;            /* value of the stopping measure when the solve stopped */
    int32_t iters;          /* PCG iterations */
    int32_t flags;          /* bit0 converged, bit1 hit max_iter, bit2 breakdown (p.Ap <= 0 or NaN) */
} hp_sweep_stat;

const char* hp_last_error(void);
int hp_version(void);

/* life cycle.  Replaces mesh + CellVariable + equation construction, main.py:29-34,103-150.
 * The initial field is Tamb_line broadcast over each row (main.py:34,152); Gamma / b snapshots are taken here. */
int hp_create(hp_handle* out, const hp_mesh_desc* mesh, const hp_phys* phys, void* stream);
int hp_destroy(hp_handle h);
int hp_set_stream(hp_handle h, void* stream);
int hp_set_opts(hp_handle h, const hp_solve_opts* opts);
/* the options in force (after the library's own adjustments: clamped values, 'mgline' replaced by the line preconditioner
 * on a mesh too short to coarsen) */
int hp_get_opts(hp_handle h, hp_solve_opts* out);

/* field access (`T.setValue`, `T.value`, main.py:152,208).  *_dev: device pointers, *_host: host pointers
 * (pinned for true asynchrony).  nz*nr doubles. */
int hp_set_T_dev(hp_handle h, const double* d_T);
int hp_get_T_dev(hp_handle h, double* d_T);
int hp_set_T_host(hp_handle h, const double* h_T);
int hp_get_T_host(hp_handle h, double* h_T);        /* synchronises the stream */
const double* hp_T_ptr(hp_handle h);                /* device pointer to owned row 0 of the live field */

/* `T.updateOld()` (main.py:269) */
int hp_update_old(hp_handle h);

/* `res = eq.sweep(var=T, dt=dt)` (main.py:202,271).  has_old = 0 reproduces hasOld=False (main.py:34):
 * the transient term uses the current iterate as old value.  If host_residual != NULL the stream is
 * synchronised and the pre-solve residual returned (FiPy semantics); otherwise nothing is synchronised. */
int hp_sweep(hp_handle h, double dt, int has_old, double* host_residual);

/* one time step of the loop body: [updateOld] + `sweeps` sweeps (main.py:200-202 / 267-271), launched
 * as one CUDA graph when loop_mode = 1 */
int hp_step(hp_handle h, double dt, int sweeps, int has_old);
/* n_steps steps back to back */
int hp_run(hp_handle h, double dt, int n_steps, int sweeps, int has_old);
/* end-to-end with HOST buffers: H2D(T_in) + n_steps steps + D2H(T_out) + stream sync */
int hp_run_host(hp_handle h, const double* h_T_in, double* h_T_out, double dt, int n_steps, int sweeps, int has_old);

/* statistics of the most recent sweeps (synchronises). Returns how many were written (<= max_n). */
int hp_get_stats(hp_handle h, hp_sweep_stat* out, int max_n);
int64_t hp_total_sweeps(hp_handle h);
int64_t hp_total_iters(hp_handle h);        /* synchronises */
int64_t hp_kernel_launches(hp_handle h);    /* kernels of this library launched so far (graph nodes counted per launch; WHILE bodies per executed iteration; synchronises) */

/* parity hooks: assemble at the current iterate and copy out the 5-point system (device pointers,
 * nz*nr doubles each; any may be NULL).  cR[j,i] couples (j,i)-(j,i+1), cZ[j,i] couples (j,i)-(j+1,i). */
int hp_assemble_debug(hp_handle h, double dt, int has_old, double* d_cR, double* d_cZ, double* d_diag, double* d_rhs);
/* elementwise property evaluation on the device (material_properties.py functions); id = hp::PropId */
int hp_eval_property(hp_handle h, int prop_id, const double* d_T, double* d_out, int64_t n);
/* same without a solver handle (material_properties.py called on CUDA tensors); stream = cudaStream_t */
int hp_eval_property_raw(const hp_phys* phys, int prop_id, const double* d_T, double* d_out, int64_t n, void* stream);
/* rho, cp per cell and face-averaged k per cell at the current iterate (profile capture, main.py:212-254) */
int hp_eval_cell_fields(hp_handle h, double* d_rho, double* d_cp, double* d_kavg);

/* per-kernel device times by CUDA events (bench.py roofline).  Between begin and end every sweep runs in
 * stream-launch mode with an event pair around each kernel launch (at most max_launches are recorded).
 * kinds: 0 other, 1 coef, 2 diag, 3 factor, 4 cg_init, 5 cg_A, 6 cg_B, 7 mg_res (fine level), 8 mg_line (fine level),
 * 9 (unused since the distributed hierarchy replaced the replicated coarse levels), 10 mg setup (coarsening + coarse pivots),
 * 11..15 mg kernels of local coarse level 1..5 (15 = level 5 and deeper) */
typedef struct hp_kernel_times {
    int64_t count[16];          /* every launch of the kind */
    double total_ms[16];
    int64_t full_count[16];     /* launches that did the work of the kind: duration >= half the longest one.  Launches   */
    double full_ms[16];         /* that return at once (sweep already converged, iterations past the end of a polled   */
                                /* chunk) would otherwise pull the per-launch average -- the roofline numerator -- down */
} hp_kernel_times;
int hp_profile_begin(hp_handle h, int max_launches);
int hp_profile_end(hp_handle h, hp_kernel_times* out);

/* axial-slab decomposition (FiPy parallelComm / PETSc ghost scatter + allreduce, main.py:11,339;
 * tests/test_parallel.py).  hp_comm_unique_id fills 128 bytes on rank 0; every rank then calls hp_comm_init.
 * After it, sweeps exchange one ghost row with each neighbour and all-reduce the CG scalars over NCCL. */
int hp_comm_unique_id(void* id128);
int hp_comm_init(hp_handle h, int rank, int nranks, const void* id128);
/* peer-memory (NVLink) exchange fused into the PCG kernels, replacing the NCCL calls inside a sweep: every rank
 * exports 4 CUDA-IPC handles (256 bytes: mailbox, (q,z) block, T, replicated global coarse levels of mgline), the host
 * gathers them (nranks*256 bytes, rank
 * order) and every rank imports them together with the owned-row count of every rank.  After it k_cg_B stores its
 * boundary rows of z and T straight into the neighbours' ghost rows and the last CTA of every reducing kernel stores
 * its sums + a sequence flag into every rank's mailbox; a one-warp kernel waits for the flags (bounded) and combines
 * the sums in rank order.  The whole time step is then one CUDA graph per rank (loop_mode 1). */
int hp_peer_export(hp_handle h, const int32_t* nz_of_rank, void* out256);
int hp_peer_import(hp_handle h, const void* all_handles, const int32_t* nz_of_rank);
/* 1 if a kernel of this handle gave up waiting for a peer (flags bit3 of hp_sweep_stat).  While the fault is up hp_sweep /
 * hp_step / hp_run / hp_run_host / hp_get_T_* / hp_get_stats return an error instead of results of a frozen simulation. */
int hp_comm_fault(hp_handle h);
/* collective recovery: after EVERY rank has synchronised its stream and passed a host-side barrier, every rank calls this
 * (and passes another barrier before the next sweep); clears the fault, the mailbox flags and the sequence counters. */
int hp_comm_clear_fault(hp_handle h);

#ifdef __cplusplus
}
#endif
#endif /* HP_SOLVER_H */
