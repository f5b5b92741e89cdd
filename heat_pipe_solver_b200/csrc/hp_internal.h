// Internal declarations shared by hp_kernels.cu (device code + launchers) and hp_api.cu (C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "hp_solver.h"

#define HP_STATS_CAP 4096
#define HP_MAX_PARTIALS 4096     // max CTAs of any reducing kernel
#define HP_NRED 4                // reduced values per kernel

#define HP_MAX_RANKS 16
#define HP_MAX_STAGES 40         // stages of one V-cycle that hand boundary rows to the neighbouring slabs
enum { HP_COMM_DIAG = 0, HP_COMM_A = 1, HP_COMM_B = 2, HP_COMM_MG = 3, HP_COMM_STAGE0 = 4, HP_COMM_KINDS = 4 + HP_MAX_STAGES };

// Peer-memory mailbox of one rank (cudaMalloc'ed, IPC-mapped into every peer of the box over NVLink).
// red[kind][src] / flag[kind][src] are WRITTEN BY rank `src` (remote stores from its reducing kernel) and read locally.
struct HpComm {
    double red[HP_COMM_KINDS][HP_MAX_RANKS][HP_NRED];
    unsigned long long flag[HP_COMM_KINDS][HP_MAX_RANKS];
};

// Device-resident scalar state of the running solve (one per handle).
struct HpState {
    double alpha, beta, rz, pq, crit, rr, res2, b2;
    int32_t iter;        // PCG iterations of the current sweep
    int32_t done;        // 1 -> remaining (A,B) launches of this sweep are no-ops
    int32_t mg_skip;     // 1 -> the V-cycle kernels of this iteration return immediately (k_cg_B found convergence)
    int32_t _pad0;
    int32_t first;       // 1 -> next direction update uses beta = 0 (p := z)
    int32_t parity;      // which p buffer holds the current search direction
    int32_t flags;       // hp_sweep_stat.flags of the current sweep
    int32_t mg_init;     // 1 -> the next V-cycle is the first of its sweep (its last kernel starts the rz recurrence: no beta)
    unsigned int ticket; // last-CTA election
    unsigned int _pad2;
    long long pipe_iters;       // PCG iterations run inside k_pipe_step (not separate launches)
    unsigned long long cycle;   // V-cycles started so far (sequence number of the stage flags between neighbouring slabs)
    long long sweep_count;
    long long total_iters;
    // A sweep that converged at its INIT decision (no iteration, no shifted start) left T, r, diag and the decision inputs
    // untouched: an identical sweep right after it (same dt / has_old, T_old untouched) is pure bookkeeping
    int32_t noop_ok;
    int32_t noop_has_old;
    double noop_dt;
};

// Geometry / region / boundary arrays (device, 1-D) + field arrays (device, ghost-inclusive rows).
struct HpDev {
    int nr, nz;
    // radial 1-D, [nr] unless noted
    const double *r_v;       // [nr+1]
    const double *r_c, *dr;
    const double *wc;        // [nr]   radial weight of cell volumes / axial face areas: r_c (cylindrical) or 1
    const double *wf;        // [nr+1] radial weight of radial face areas: r_v (cylindrical) or 1
    const double *ar;        // arithmetic-face weight of radial face i+1/2: (r_v[i+1]-r_c[i])/(r_c[i+1]-r_c[i])
    const double *dcr;       // r_c[i+1]-r_c[i]
    const int8_t *cell_band, *rface_band, *zface_band;
    // axial 1-D, ghost-inclusive [nz+2]
    const double *z_c, *dz;
    const double *az;        // weight of the face between row g and g+1
    const double *dcz;       // z_c[g+1]-z_c[g]
    const uint8_t *zc_flags, *zv_flags, *axial_open;
    const double *q_line, *h_line, *Tamb_line, *eps_line;
    const double *k0_out;    // snapshot of the outer-surface k (main.py:145)
    double A_out_r;          // wf[nr]
    double d_out;            // r_v[nr]-r_c[nr-1]
    // fields, (nz+2)*nr each, row g at g*nr
    double *T, *Told, *cR, *cZ, *diag, *dinv, *r, *q, *z, *p0, *p1;
    double *d1, *d2;      // increments of the two steps before the last one (higher-order extrapolated start); may be null
    double *c2;           // first-sweep solution of the running step / correction the later sweeps added to it; may be null
    HpState* st;
    hp_sweep_stat* stats;    // ring of HP_STATS_CAP
    double* partials;        // [HP_NRED * HP_MAX_PARTIALS]
    // axial slabs (nranks > 1): per-rank sums and their all-gathered copies
    int nranks, rank;
    int ncolA;               // column tiles of k_cg_A
    // peer-memory (NVLink) exchange fused into the kernels (p2p != 0): no NCCL call inside a sweep
    int p2p;
    HpComm* comm_peer[HP_MAX_RANKS];   // every rank's mailbox (own one included) as seen from this GPU
    unsigned long long* seq;           // [HP_COMM_KINDS] how many times this rank has produced each kind
    int mg_local;                      // 1: the V-cycle ignores the couplings to the ghost rows (slabs without peer memory)
    int* fault_host;                   // pinned host word (mapped): set when a wait for a peer timed out
    double *z_lo, *z_hi;               // neighbour's z ghost row that mirrors my first / last owned row (or null)
    double *T_lo, *T_hi;               // same for the field T
    double* red_local;       // [HP_NRED]
    double* red_all;         // [nranks * HP_NRED]
};

// One level of the axially semi-coarsened hierarchy of the 'mgline' preconditioner (rows are aggregated in pairs,
// radial lines stay whole; every level keeps the 5-point structure and the ghost-row layout of the fine fields).
#define HP_MAX_LEVELS 12
struct HpLevel {
    int nz;                              // owned rows of this level
    double *cR, *cZ, *diag, *dinv;       // operator (Galerkin, piecewise-constant aggregation) + line pivots
    double *rhs, *x, *x2, *tmp;          // right-hand side, correction, prolonged correction, residual scratch
};

struct HpLevelSet { int n; HpLevel lev[2 * HP_MAX_LEVELS]; };   // entry 0 is skipped by k_factor_levels

// Hand-over of the two boundary rows a V-cycle stage produces to the neighbouring slabs (peer-memory stores fused into
// the producing kernel) and the wait of the consuming stage, fused into its first / last row team.
struct HpXchg {
    double* push_lo;     // lower neighbour's ghost row that mirrors the first row of this stage's output (or null)
    double* push_hi;     // upper neighbour's ghost row that mirrors the last row (or null)
    int sig;             // >= 0: flag slot HP_COMM_STAGE0 + k raised at the neighbours once the rows are stored
    int wait;            // >= 0: flag slot of the stage whose rows this kernel reads from its ghost rows
};

// arguments of k_mg_res / k_mg_line (hp_kernels.cu)
struct HpMgResArgs {
    const double* xin; const double* ec; double* xout; double* resout; double* crhs;
    double scale; int nzc;
    const double* c_dinv; const double* c_cR; double* c_x;     // k_mg_fused MODE 2: pivots / couplings / correction of level+1
    int xout_ghost;      // also write the two ghost rows of xout (the scaled iterate of the level-0 down stage)
    HpXchg xc;
};
struct HpMgLineArgs {
    const double* rhs; const double* base; double* out; int final_;
    HpXchg xc;
};

// Per-launch scalar arguments shared by all kernels: signature (HpDev, hp_phys, HpKArgs [, extra]).
struct HpKArgs {
    double dt;
    double tol;
    int32_t has_old;
    int32_t max_iter;
    int32_t criterion;
    int32_t use_cond;
    int32_t _pad1;
    int32_t mg;                // 1: 'mgline' preconditioner: k_cg_B only decides convergence, the V-cycle's last kernel
                               //    produces z, r.z and advances the recurrences
    int32_t init_phase;        // (unused since the V-cycle opens the loop body: HpState.mg_init tells the first cycle of a sweep)
    double omega;              // damping of the line smoother in the V-cycle
    double ext_guard;          // k_delta: the cubic term of the predicted start is applied where it exceeds this [K]
    int32_t ext_order;         // k_delta: order of the extrapolated increment (1 linear, 2 quadratic, 3 cubic in the step index)
    int32_t c2_valid;          // k_delta / k_stash: d.c2 holds usable history
    int32_t noop_enable;       // k_diag / INIT kernel may take the bookkeeping-only path (single GPU, no debug outputs, no shift)
    int32_t shift;             // 1: this (A, B) pair moves the start of the solve to x0 + z (z = extrapolated increment):
                               //    alpha := 1, and k_cg_B finalises like the INIT kernel
    unsigned long long timeout_ns;   // bound of every wait for a peer flag
    unsigned long long cond;   // cudaGraphConditionalHandle of the enclosing WHILE node (use_cond != 0)
    double* rhs_out;           // optional (debug): full rhs, ghost-inclusive layout
};

struct HpLaunchCfg {
    // k_cg_B / k_cg_init (whole rows per team: the line solve scans along the row)
    int tpl, ept, vec;       // threads per line, elements per thread, vector loads ok
    int cta_threads, teams_per_cta;              // marching kernels; their grid = one resident wave (team_grid)
    // k_cg_A (no row-wide dependency: rows wider than tplA*eptA are cut into ncolA column tiles)
    int tplA, eptA, vecA, ctaA, teams_per_ctaA, ncolA;
    int clF;                 // fused V-cycle stages: CTAs per row (thread-block cluster), 1 = one CTA team, 0 = unsupported
    int num_sms, nz;
    int grid_cell, cta_cell;                     // grid for thread-per-cell kernels
};

// launch descriptors (so that hp_api.cu can either launch on a stream or add CUDA-graph kernel nodes)
struct HpKernelLaunch {
    const void* func;
    dim3 grid, block;
    size_t smem;
};

namespace hpk {
bool pick_config(int nr, int nz, int num_sms, HpLaunchCfg* cfg, const char** err);

// Each returns the launch descriptor; kernel arguments are passed by the caller as void*[] pointing to
// variables with exactly these types/order (documented next to each kernel in hp_kernels.cu).
HpKernelLaunch desc_snapshot_kout(const HpLaunchCfg&, const HpDev&);
HpKernelLaunch desc_coef(const HpLaunchCfg&, const HpDev&);
HpKernelLaunch desc_stash(const HpLaunchCfg&, const HpDev&);   // second sweep of a step: z <- c2 (predicted correction), c2 <- T
HpKernelLaunch desc_delta(const HpLaunchCfg&, const HpDev&);   // z = extrapolated increment, Told <- T, on every row (ghosts included)
HpKernelLaunch desc_diag(const HpLaunchCfg&, const HpDev&);     // args: (HpDev, hp_phys, HpKArgs, int write_dinv, int tpr_log2)
int diag_tpr_log2(int nr);
HpKernelLaunch desc_factor_rows(const HpLaunchCfg&, long long rows);   // args: (HpLevelSet, int first, int nr, const int* skip)
HpKernelLaunch desc_cg_init(const HpLaunchCfg&, const HpDev&, int precond, int criterion);
HpKernelLaunch desc_cg_A(const HpLaunchCfg&, const HpDev&);
// 'mgline' V-cycle pieces (argument lists next to the kernels in hp_kernels.cu)
HpKernelLaunch desc_mg_coarsen(const HpLaunchCfg&, int nzc, int nr);   // args: (HpLevel f, HpLevel c, int nr, double zscale, int keep_ghost, const int* skip)
HpKernelLaunch desc_mg_res(const HpLaunchCfg&, int nz_level);
HpKernelLaunch desc_mg_line(const HpLaunchCfg&, int nz_level);
HpKernelLaunch desc_mg_down(const HpLaunchCfg&, int nz_level);   // same args: residual + restriction + pre-smoothing of level+1
HpKernelLaunch desc_mg_final(const HpLaunchCfg&, int nz_level);  // same args: level-0 prolong + residual + smoothing = z, r.z (end of the cycle)
HpKernelLaunch desc_mg_up(const HpLaunchCfg&, int nz_level);     // args: (HpDev, HpKArgs, HpLevel, HpMgResArgs): fused prolong+residual+smooth
bool mg_fused_supported(const HpLaunchCfg&);
HpKernelLaunch desc_wait();       // args: (HpDev, hp_phys, HpKArgs, int kind): peer-memory variant of finalize
HpKernelLaunch desc_finalize();   // args: (HpDev, hp_phys, HpKArgs, int kind) kind: 0 diag, 1 A, 2 B, 3 init
HpKernelLaunch desc_cg_B(const HpLaunchCfg&, const HpDev&, int precond, int criterion, int tma);
bool cg_B_tma_supported(const HpLaunchCfg&, int precond, int criterion);
// persistent one-pipe-per-CTA step: args (HpDev, hp_phys, HpKArgs, int nzm, int sweeps, int update_old, int predict, int track, double* part)
bool pipe_kernel_supported(int nr, int nzm);
HpKernelLaunch desc_pipe_step(int nr, int nzm, int members);
HpKernelLaunch desc_pipe_finalize();                 // args (HpDev, HpKArgs, int M, int sweeps, const double* part)

cudaError_t launch_eval_property(const hp_phys& ph, int id, const double* T, double* out, long long n, cudaStream_t s);
cudaError_t launch_cell_fields(const HpDev& d, const hp_phys& ph, double* rho, double* cp, double* kavg, cudaStream_t s);
cudaError_t launch_copy_owned(const double* src_ghost, double* dst, int nr, long long n, cudaStream_t s);
cudaError_t launch_fill_rows(const HpDev& d, double* field, cudaStream_t s);
}  // namespace hpk
