// CUDA kernels (sm_100a, fp64) of the transient r-z conduction hot path.
//
//   k_snapshot_kout  outer-surface k at T_amb                      (main.py:145  `b = FaceVariable(value=k)`)
//   k_coef           face conductivities -> couplings cR, cZ       (main.py:103-106,133; FiPy DiffusionTerm)
//   k_diag           rho*cp, diagonal, boundary terms, r0 = b - L T, ||L T - b||^2   (main.py:108-117,123-150,202)
//   k_factor_team    LDL^T pivots of the radial-line tridiagonal blocks (2x2 scan on row teams)
//   k_cg_B<INIT>     z = M^-1 r0, r.z                                (PCG start)
//   k_cg_A           p = z + beta p ; q = L p ; p.q                  (PCG, fused direction update + stencil + dot)
//   k_cg_B           x += alpha p ; r -= alpha q ; z = M^-1 r ; r.z  (PCG, fused updates + line solve + dots)
//   k_cg_B_tma       the same with cp.async.bulk / mbarrier staged operand rows (opt-in)
//   k_mg_coarsen, k_mg_res, k_mg_line, k_mg_fused<1|2|3>, k_mg_fused_cl2/cl4   the 'mgline' V-cycle (axial semi-coarsening,
//                    line smoother; fused stages, cluster teams for wide rows, stage hand-over between slabs)
//   k_delta, k_stash predicted PCG starts from the step history
//   k_pipe_step      one whole time step of one small pipe per CTA, on chip (native 500 x 9 mesh, ensembles)
//   k_wait, k_finalize   scalar recurrences under axial slabs (peer-memory flags / NCCL all-gather)
//
// Layout: every field has nz+2 rows of nr doubles (row 0 / nz+1 = ghost rows, zero-coupled at physical
// ends, a neighbour's row under axial slabs).  r is the fast index, so flat neighbours are idx+-1 (radial;
// the last coupling of every row is 0, which also terminates the row at idx-1 of the next one) and
// idx+-nr (axial).  Reductions are deterministic: per-CTA partials + fixed-order final sum by the last
// CTA (ticket), which also advances the PCG scalar recurrences on the device -- no host round trip.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "hp_internal.h"
#include "hp_props.cuh"


namespace {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld256(const double* p, double* v) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st256(double* p, const double* v) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

template <int EPT, bool VEC>
__device__ __forceinline__ void load_row(const double* __restrict__ row, int i0, int nr, double (&v)[EPT]) {
    if constexpr (VEC) {
        if (i0 < nr) {
#pragma unroll
            for (int k = 0; k < EPT; k += 4) ld256(row + i0 + k, &v[k]);
        } else {
#pragma unroll
            for (int k = 0; k < EPT; ++k) v[k] = 0.0;
        }
    } else {
#pragma unroll
        for (int k = 0; k < EPT; ++k) v[k] = (i0 + k < nr) ? row[i0 + k] : 0.0;
    }
}
template <int EPT, bool VEC>
__device__ __forceinline__ void store_row(double* __restrict__ row, int i0, int nr, const double (&v)[EPT]) {
    if constexpr (VEC) {
        if (i0 < nr) {
#pragma unroll
            for (int k = 0; k < EPT; k += 4) st256(row + i0 + k, &v[k]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < EPT; ++k)
            if (i0 + k < nr) row[i0 + k] = v[k];
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-reduce NV values (bit k of max_mask: value k is max-reduced, else summed), publish the CTA partial,
// elect the last CTA by ticket; in the last CTA reduce all partials in a fixed order.  Returns true for every
// thread of the last CTA; tot[] is valid on thread 0 of it.  All threads of the CTA must call.
template <int NV>
__device__ bool cta_publish(double (&v)[NV], unsigned max_mask, double* __restrict__ partials, unsigned int* ticket,
                            double (&tot)[NV], bool sys_scope = false) {
    __shared__ double s_red[NV][32];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double w = ((max_mask >> k) & 1u) ? warp_max(v[k]) : warp_sum(v[k]);
        if (lane == 0) s_red[k][warp] = w;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const bool mx = (max_mask >> k) & 1u;
            double w = (lane < nwarp) ? s_red[k][lane] : 0.0;   // all reduced quantities are >= 0 where max is used
            w = mx ? warp_max(w) : warp_sum(w);
            if (lane == 0) partials[k * HP_MAX_PARTIALS + blockIdx.x] = w;
        }
        if (lane == 0) {
            if (sys_scope) __threadfence_system();      // remote (peer) stores of this CTA before the ticket
            else __threadfence();
            unsigned t = atomicAdd(ticket, 1u);
            s_last = (t == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const bool mx = (max_mask >> k) & 1u;
        double w = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            double pv = __ldcg(&partials[k * HP_MAX_PARTIALS + b]);
            w = mx ? fmax(w, pv) : (w + pv);
        }
        w = mx ? warp_max(w) : warp_sum(w);
        __syncthreads();
        if (lane == 0) s_red[k][warp] = w;
        __syncthreads();
        if (warp == 0) {
            double u = (lane < nwarp) ? s_red[k][lane] : 0.0;
            u = mx ? warp_max(u) : warp_sum(u);
            tot[k] = u;
        }
    }
    if (threadIdx.x == 0) *ticket = 0u;
    return true;
}

__device__ __forceinline__ void set_cond(const HpKArgs& a, int keep_going) {
    if (a.use_cond) cudaGraphSetConditional((cudaGraphConditionalHandle)a.cond, keep_going ? 1u : 0u);
}

__device__ __forceinline__ bool converged(const HpKArgs& a, double crit, double rr, double b2) {
    if (a.criterion == 2) return rr <= a.tol * a.tol * b2;
    return crit <= a.tol;
}

// ------------------------------------------------------------------------------------------------
// scalar recurrences of the sweep / PCG, executed by ONE thread per kernel: the last CTA of the producing
// kernel (single GPU) or k_finalize after the all-gather of the per-rank sums (axial slabs).
// ------------------------------------------------------------------------------------------------
enum { FIN_DIAG = 0, FIN_A = 1, FIN_B = 2, FIN_INIT = 3, FIN_MGFINAL = 4 };

__device__ __forceinline__ void finalize_diag(const HpDev& d, const HpKArgs& a, double res2, double b2) {
    HpState* s = d.st;
    s->res2 = res2;
    s->b2 = b2;
    s->iter = 0; s->done = 0; s->mg_skip = 0; s->first = 1; s->flags &= 8;      // bit3 (a peer never arrived) is sticky
    hp_sweep_stat* stt = &d.stats[s->sweep_count % HP_STATS_CAP];
    stt->residual_l2 = sqrt(res2);
    stt->rhs_l2 = sqrt(b2);
    stt->crit = 0.0; stt->iters = 0; stt->flags = 0;
    s->sweep_count += 1;
}

__device__ __forceinline__ void finalize_A(const HpDev& d, const HpKArgs& a, double pq) {
    HpState* s = d.st;
    s->pq = pq;
    if (a.shift) {
        s->alpha = 1.0;         // not a CG step: the pair only moves the start of the solve
    } else if (pq > 0.0 && isfinite(pq)) {
        s->alpha = s->rz / pq;
    } else {
        s->alpha = 0.0;
        s->flags |= 4;          // breakdown
    }
    s->parity ^= 1;
    s->first = 0;
}

// true when this sweep repeats the previous one exactly (see HpState.noop_ok): k_diag and the INIT kernel then only redo
// the bookkeeping (sweep count, statistics entry, WHILE condition) from the values kept in the state
__device__ __forceinline__ bool sweep_is_noop(const HpDev& d, const HpKArgs& a) {
    const HpState* s = d.st;
    return a.noop_enable && !a.shift && !a.rhs_out && s->noop_ok && s->noop_dt == a.dt && s->noop_has_old == a.has_old;
}

// stage 0: plain PCG (k_cg_B produced z = M^-1 r and r.z).  'mgline': stage 1 after k_cg_B (x, r updated, the
// line-preconditioned residual measures convergence; the V-cycle only runs if the solve goes on), stage 2 after the
// last kernel of the V-cycle (z and r.z are final: beta / rz recurrences).
__device__ __forceinline__ void finalize_B(const HpDev& d, const HpKArgs& a, bool init, double rz_new, double rr, double crit,
                                           int stage = 0) {
    HpState* s = d.st;
    int flags = s->flags;
    int done = 0;
    hp_sweep_stat* stt = &d.stats[(s->sweep_count - 1) % HP_STATS_CAP];
    // mgline: the V-cycle opens the loop body, so the same kernel nodes serve the first cycle of a sweep (after the INIT
    // kernel / the shift pair) and the later ones: the INIT-position decision leaves the mark, the cycle's end consumes it
    if (stage == 1 && init) s->mg_init = 1;
    if (stage == 2) { init = s->mg_init != 0; s->mg_init = 0; }
    if (stage != 1) {
        if (init) {
            s->rz = rz_new;
            s->parity = 0;
            s->first = 1;
        } else {
            s->beta = (s->rz != 0.0) ? rz_new / s->rz : 0.0;
            s->rz = rz_new;
        }
    }
    if (stage != 2) {
        if (!init) { s->iter += 1; s->total_iters += 1; }
        s->rr = rr; s->crit = crit;
    }
    if (stage == 2) {
        if (!isfinite(rz_new)) { flags |= 4; s->done = 1; s->flags = flags; stt->flags = flags; set_cond(a, 0); }
        return;
    }
    const bool nan_bad = !(isfinite(crit) && (stage == 1 || isfinite(rz_new)));
    if (converged(a, crit, rr, s->b2)) { done = 1; flags |= 1; }
    else if (nan_bad || (flags & 4)) { done = 1; flags |= 4; }
    else if (s->iter >= a.max_iter) { done = 1; flags |= 2; }
    s->done = done; s->flags = flags;
    s->mg_skip = done;
    if (stage == 1 && !done) s->cycle += 1;      // a V-cycle follows: sequence number of its stage flags (same on every rank)
    // converged before any iteration, from the un-shifted start: nothing of the sweep's inputs or outputs moved
    if (init && !a.shift && flags == 1 && s->iter == 0) { s->noop_ok = 1; s->noop_dt = a.dt; s->noop_has_old = a.has_old; }
    else s->noop_ok = 0;
    stt->crit = (a.criterion == 2) ? sqrt(rr) : crit;
    stt->iters = s->iter;
    stt->flags = flags;
    set_cond(a, !done);
}

// slabs: this rank's sums go to red_local; k_finalize combines all ranks' after the all-gather
template <int NV>
__device__ __forceinline__ void stash_local(const HpDev& d, const double (&tot)[NV], int kind) {
    if (!d.p2p) {
#pragma unroll
        for (int k = 0; k < NV; ++k) d.red_local[k] = tot[k];
        return;
    }
    // peer-memory path: my sums into slot [kind][me] of EVERY rank's mailbox, then the sequence flag (release, system
    // scope).  The boundary rows other CTAs of this kernel stored into the neighbours' ghost rows are ordered before
    // the flag through their system fence + ticket (cta_publish) and this fence.
    const unsigned long long sq = d.seq[kind] + 1ull;
    d.seq[kind] = sq;
    for (int r = 0; r < d.nranks; ++r) {
        HpComm* c = d.comm_peer[r];
#pragma unroll
        for (int k = 0; k < NV; ++k) c->red[kind][d.rank][k] = tot[k];
    }
    // ONE system fence orders the sums (and, through the tickets, every CTA's boundary rows) before all the flags; the flag
    // stores themselves are relaxed and pipeline over NVLink (a release per flag serialises a round trip per rank: 8 ranks
    // cost ~25 us per reducing kernel)
    __threadfence_system();
    for (int r = 0; r < d.nranks; ++r) {
        unsigned long long* f = &d.comm_peer[r]->flag[kind][d.rank];
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(f), "l"(sq) : "memory");
    }
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// a wait for a peer gave up: sticky bit3 of the sweep flags + the host-visible fault word (every later API call on the handle
// fails until hp_comm_clear_fault)
__device__ __forceinline__ void raise_fault(const HpDev& d) {
    atomicOr(&d.st->flags, 8);
    if (d.fault_host) { *(volatile int*)d.fault_host = 1; __threadfence_system(); }
}
// bounded wait until the flag `slot` written by rank `src` into this rank's mailbox reaches `want`
__device__ __forceinline__ bool wait_peer_flag(const HpDev& d, const HpKArgs& a, int slot, int src, unsigned long long want) {
    const unsigned long long* f = &d.comm_peer[d.rank]->flag[slot][src];
    if (ld_acquire_sys(f) >= want) return true;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(f) < want) {
        if (global_ns() - t0 > a.timeout_ns) return false;
        __nanosleep(32);
    }
    return true;
}

// k_wait: one warp; lane r waits until rank r's flag of `kind` reaches this rank's own production count (all ranks run
// the same kernel sequence), then lane 0 combines the sums in rank order and advances the scalar recurrences.
// The wait is bounded (hp_solve_opts.peer_timeout_ms): on timeout the sweep is aborted and the handle's fault is raised.
__global__ void k_wait(HpDev d, hp_phys ph, HpKArgs a, int fin_kind) {
    const int kind = (fin_kind == FIN_DIAG) ? HP_COMM_DIAG : (fin_kind == FIN_A ? HP_COMM_A : (fin_kind == FIN_MGFINAL ? HP_COMM_MG : HP_COMM_B));
    HpState* s = d.st;
    if ((fin_kind == FIN_A || fin_kind == FIN_B) && s->done) return;
    if (fin_kind == FIN_MGFINAL && s->mg_skip) return;
    const int lane = threadIdx.x;
    const HpComm* mine = d.comm_peer[d.rank];
    const unsigned long long want = d.seq[kind];
    bool ok = true;
    if (lane < d.nranks && !(s->flags & 8)) ok = wait_peer_flag(d, a, kind, lane, want);
    const unsigned bad = __ballot_sync(0xffffffffu, !ok);
    if (lane != 0) return;
    (void)mine;
    double v[HP_NRED] = {0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < d.nranks; ++r) {
        const volatile double* p = mine->red[kind][r];
        if (fin_kind == FIN_DIAG) { v[0] += p[0]; v[1] += p[1]; }
        else if (fin_kind == FIN_A || fin_kind == FIN_MGFINAL) { v[0] += p[0]; }
        else { v[0] += p[0]; v[1] += p[1]; v[2] = fmax(v[2], p[2]); }
    }
    if (fin_kind == FIN_DIAG) finalize_diag(d, a, v[0], v[1]);
    else if (fin_kind == FIN_A) finalize_A(d, a, v[0]);
    else if (fin_kind == FIN_MGFINAL) finalize_B(d, a, a.init_phase != 0, v[0], 0.0, 0.0, 2);
    else finalize_B(d, a, fin_kind == FIN_INIT || a.shift, v[0], v[1], v[2], a.mg ? 1 : 0);
    if (bad || (s->flags & 8)) {          // a peer never arrived: stop this sweep, make the failure sticky
        raise_fault(d);
        s->done = 1; s->mg_skip = 1;
        d.stats[(s->sweep_count - 1) % HP_STATS_CAP].flags |= 8;
        set_cond(a, 0);
    }
}

__global__ void k_finalize(HpDev d, hp_phys ph, HpKArgs a, int kind) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if ((kind == FIN_A || kind == FIN_B) && d.st->done) return;
    if (kind == FIN_MGFINAL && d.st->mg_skip) return;
    double v[HP_NRED];
    for (int k = 0; k < HP_NRED; ++k) v[k] = 0.0;
    for (int r = 0; r < d.nranks; ++r) {           // fixed rank order: every rank computes identical scalars
        const double* p = d.red_all + r * HP_NRED;
        if (kind == FIN_DIAG) { v[0] += p[0]; v[1] += p[1]; }
        else if (kind == FIN_A || kind == FIN_MGFINAL) { v[0] += p[0]; }
        else { v[0] += p[0]; v[1] += p[1]; v[2] = fmax(v[2], p[2]); }
    }
    if (kind == FIN_DIAG) finalize_diag(d, a, v[0], v[1]);
    else if (kind == FIN_A) finalize_A(d, a, v[0]);
    else if (kind == FIN_MGFINAL) finalize_B(d, a, a.init_phase != 0, v[0], 0.0, 0.0, 2);
    else finalize_B(d, a, kind == FIN_INIT || a.shift, v[0], v[1], v[2], a.mg ? 1 : 0);
}

// ------------------------------------------------------------------------------------------------
// face conductivity helpers
// ------------------------------------------------------------------------------------------------
// k on the radial face between (g,i) and (g,i+1)
__device__ __forceinline__ double k_rface(const HpDev& d, const hp_phys& ph, int g, int i, double Tc, double Te) {
    const unsigned zf = d.zc_flags[g];
    if (ph.face_k == 0) {
        double Tf = Tc + (Te - Tc) * d.ar[i];
        return hp::k_region(ph, d.rface_band[i], zf, Tf);
    }
    double k1 = hp::k_region(ph, d.cell_band[i], zf, Tc);
    double k2 = hp::k_region(ph, d.cell_band[i + 1], zf, Te);
    double d1 = d.r_v[i + 1] - d.r_c[i], d2 = d.r_c[i + 1] - d.r_v[i + 1];
    double kk = (d1 + d2) / (d1 / k1 + d2 / k2);
    return isfinite(kk) ? kk : 0.0;
}
// k on the axial face between (g,i) and (g+1,i)
__device__ __forceinline__ double k_zface(const HpDev& d, const hp_phys& ph, int g, int i, double Tc, double Tn) {
    if (ph.face_k == 0) {
        double Tf = Tc + (Tn - Tc) * d.az[g];
        return hp::k_region(ph, d.zface_band[i], d.zv_flags[g], Tf);
    }
    double k1 = hp::k_region(ph, d.cell_band[i], d.zc_flags[g], Tc);
    double k2 = hp::k_region(ph, d.cell_band[i], d.zc_flags[g + 1], Tn);
    double zv = d.z_c[g] + 0.5 * d.dz[g];
    double e1 = zv - d.z_c[g], e2 = d.z_c[g + 1] - zv;
    double kk = (e1 + e2) / (e1 / k1 + e2 / k2);
    return isfinite(kk) ? kk : 0.0;
}
// k on the outer surface of row g (exterior face: T_f = T_P)
__device__ __forceinline__ double k_outface(const HpDev& d, const hp_phys& ph, int g, double Tc) {
    const int i = d.nr - 1;
    if (ph.face_k == 0) return hp::k_region(ph, d.rface_band[i], d.zc_flags[g], Tc);
    return hp::k_region(ph, d.cell_band[i], d.zc_flags[g], Tc);
}

// ------------------------------------------------------------------------------------------------
// k_snapshot_kout : k0_out[g] = k(outer face, T_amb[g])                       grid: rows
// ------------------------------------------------------------------------------------------------
__global__ void k_snapshot_kout(HpDev d, hp_phys ph, HpKArgs a, double* k0_out) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < d.nz + 2) k0_out[g] = k_outface(d, ph, g, d.Tamb_line[g]);
}

// ------------------------------------------------------------------------------------------------
// k_coef : cR[g,i] (owned rows), cZ[g,i] (rows 0..nz) from the field T           thread per cell
//   cR = Gamma_r * (wf[i+1] * dz[g]) / (r_c[i+1]-r_c[i])       FiPy: coeff * faceArea / cellDistance
//   cZ = Gamma_z * (wc[i] * dr[i])   / (z_c[g+1]-z_c[g])       (interior faces only)
//   wf = r_v, wc = r_c on CylindricalGrid2D (areas / volumes per radian); 1 on a Cartesian Grid2D
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_coef(HpDev d, hp_phys ph, HpKArgs a) {
    const long long n = (long long)(d.nz + 1) * d.nr;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(idx / d.nr), i = (int)(idx - (long long)g * d.nr);
        const double Tc = d.T[idx];
        double cz = 0.0;
        if (d.axial_open[g]) {
            const double Tn = d.T[idx + d.nr];
            cz = k_zface(d, ph, g, i, Tc, Tn) * ((d.wc[i] * d.dr[i]) / d.dcz[g]);
        }
        d.cZ[idx] = cz;
        if (g >= 1) {
            double cr = 0.0;
            if (i < d.nr - 1) {
                const double Te = d.T[idx + 1];
                cr = k_rface(d, ph, g, i, Tc, Te) * ((d.wf[i + 1] * d.dz[g]) / d.dcr[i]);
            }
            d.cR[idx] = cr;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_diag : per owned cell                                                       thread per cell
//   cap  = cp*rho*V/dt                         TransientTerm(coeff=cp*rho)        main.py:108-117,148
//   diag = cap + cE + cW + cN + cS (+ Robin)   DiffusionTerm + ImplicitSourceTerm main.py:148-149
//   rhs  = cap*T_old (+ Robin*g + evaporator)                                     main.py:146-150
//   r    = rhs - L T          (PCG start residual; FiPy's `res` = ||L T - rhs||_2) main.py:202
// ------------------------------------------------------------------------------------------------
// one cell of the assembly: diagonal, right-hand side and start residual from the loaded neighbourhood
struct DiagCell { double dg, rhs, r0; };
__device__ __forceinline__ DiagCell diag_cell(const HpDev& d, const hp_phys& ph, const HpKArgs& a, int g, int i, unsigned zf, double dzg,
                                              double Tc, double To, double cE, double cW, double cN, double cS, double Tnb) {
    double rho, cp;
    hp::rho_cp_region(ph, d.cell_band[i], zf, Tc, rho, cp);
    const double vol = d.wc[i] * d.dr[i] * dzg;
    const double cap = cp * rho * vol / a.dt;
    double dg = cap;
    dg += cE; dg += cW; dg += cN; dg += cS;
    double rhs = cap * To;
    if (i == d.nr - 1) {
        const double kout = k_outface(d, ph, g, Tc);
        const double k0 = (ph.gamma == 0) ? d.k0_out[g] : kout;
        const double A_out = d.A_out_r * dzg;
        const double h = d.h_line[g], Ta = d.Tamb_line[g];
        if (zf & HP_ZF_BC_COND) {
            const double robin = kout * A_out / (h * d.d_out + k0);
            dg += robin * h;
            double gs = h * Ta;
            if (ph.radiation) gs += ph.sigma * d.eps_line[g] * (Ta * Ta * Ta * Ta - Tc * Tc * Tc * Tc);
            rhs += robin * gs;
        }
        if (zf & HP_ZF_BC_EVAP) {
            const double q = d.q_line[g];
            rhs += ((ph.source == 0) ? (q / kout) : q) * A_out;
        }
    }
    const double LT = dg * Tc - Tnb;
    return DiagCell{dg, rhs, rhs - LT};
}

constexpr int DIAG_U = 4;      // cells per thread whose loads are issued before any of their (long, fp64) property chains
template <bool ROWS_OVER_LANES>
__global__ void __launch_bounds__(256, 2) k_diag(HpDev d, hp_phys ph, HpKArgs a, int write_dinv_jacobi, int tpr_log2) {
    // rows over CTAs (and over row groups inside a CTA when a row is shorter than the CTA), columns over threads:
    // no integer division per cell, per-row quantities read once
    const int nr = d.nr;
    if (sweep_is_noop(d, a)) {          // same T, T_old, dt as the sweep before: diag, r and both sums are what they were
        if (blockIdx.x == 0 && threadIdx.x == 0) finalize_diag(d, a, d.st->res2, d.st->b2);
        return;
    }
    const int tpr = 1 << tpr_log2, rpb = blockDim.x >> tpr_log2;
    const int col0 = threadIdx.x & (tpr - 1), rsub = threadIdx.x >> tpr_log2;
    double acc[2] = {0.0, 0.0};
    if constexpr (ROWS_OVER_LANES) {
        // Very short rows (the native 9-cell rows of the ensembles): with columns over lanes a warp holds cells of all three
        // bands and executes the vapor-core, wick AND wall laws one after the other.  Here lanes run over ROWS and a warp
        // handles one column at a time (one band, one law); the strided loads of a tile of 32 rows hit the same L1 lines
        // column after column.
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
        for (long long j0 = (long long)blockIdx.x * 32; j0 < d.nz; j0 += (long long)gridDim.x * 32) {
            const long long j = j0 + lane;
            if (j < d.nz) {
                const int g = (int)j + 1;
                const long long rb = (long long)g * nr;
                const unsigned zf = d.zc_flags[g];
                const double dzg = d.dz[g];
                for (int i = warp; i < nr; i += nwarp) {
                    const long long idx = rb + i;
                    const double Tc = d.T[idx], To = a.has_old ? d.Told[idx] : Tc;
                    const double cE = d.cR[idx], cW = d.cR[idx - 1], cN = d.cZ[idx], cS = d.cZ[idx - nr];
                    const double Tnb = cE * d.T[idx + 1] + cW * d.T[idx - 1] + cN * d.T[idx + nr] + cS * d.T[idx - nr];
                    const DiagCell c = diag_cell(d, ph, a, g, i, zf, dzg, Tc, To, cE, cW, cN, cS, Tnb);
                    d.diag[idx] = c.dg;
                    d.r[idx] = c.r0;
                    if (write_dinv_jacobi) d.dinv[idx] = 1.0 / c.dg;
                    if (a.rhs_out) a.rhs_out[idx] = c.rhs;
                    acc[0] += c.r0 * c.r0;
                    acc[1] += c.rhs * c.rhs;
                }
            }
        }
    } else {
    for (long long j = (long long)blockIdx.x * rpb + rsub; j < d.nz; j += (long long)gridDim.x * rpb) {
        const int g = (int)j + 1;
        const long long rb = (long long)g * nr;
        const unsigned zf = d.zc_flags[g];
        const double dzg = d.dz[g];
        // The vapor-core columns cost several times the wall columns (pow / exp / log chains) and sit in the first warps'
        // column groups: rotate the warp -> column-group assignment with the row so that every warp of the CTA gets its
        // share of them (no barrier in this loop: the warps run through their rows independently)
        const int colr = (tpr == 256) ? ((col0 + 32 * (int)(j & 7)) & 255) : col0;
        for (int ib = colr; ib < nr; ib += tpr * DIAG_U) {
            double Tc[DIAG_U], To[DIAG_U], cE[DIAG_U], cW[DIAG_U], cN[DIAG_U], cS[DIAG_U], Tnb[DIAG_U];
#pragma unroll
            for (int u = 0; u < DIAG_U; ++u) {
                const int i = ib + u * tpr;
                const long long idx = rb + (i < nr ? i : ib);          // out-of-row slots re-read a valid cell, never stored
                Tc[u] = d.T[idx];
                To[u] = a.has_old ? d.Told[idx] : Tc[u];
                cE[u] = d.cR[idx]; cW[u] = d.cR[idx - 1]; cN[u] = d.cZ[idx]; cS[u] = d.cZ[idx - nr];
                // off-diagonal part of L T (the diagonal term joins once diag is known)
                Tnb[u] = cE[u] * d.T[idx + 1] + cW[u] * d.T[idx - 1] + cN[u] * d.T[idx + nr] + cS[u] * d.T[idx - nr];
            }
#pragma unroll
            for (int u = 0; u < DIAG_U; ++u) {
                const int i = ib + u * tpr;
                if (i >= nr) break;
                const long long idx = rb + i;
                const DiagCell c = diag_cell(d, ph, a, g, i, zf, dzg, Tc[u], To[u], cE[u], cW[u], cN[u], cS[u], Tnb[u]);
                d.diag[idx] = c.dg;
                d.r[idx] = c.r0;
                if (write_dinv_jacobi) d.dinv[idx] = 1.0 / c.dg;
                if (a.rhs_out) a.rhs_out[idx] = c.rhs;
                acc[0] += c.r0 * c.r0;
                acc[1] += c.rhs * c.rhs;
            }
        }
    }
    }
    double tot[2];
    if (cta_publish<2>(acc, 0u, d.partials, &d.st->ticket, tot, d.p2p != 0) && threadIdx.x == 0) {
        if (d.nranks > 1) stash_local<2>(d, tot, HP_COMM_DIAG);
        else finalize_diag(d, a, tot[0], tot[1]);
    }
}

// ------------------------------------------------------------------------------------------------
// marching "row team" kernels.  A team of TPL threads owns whole rows (EPT consecutive cells per thread,
// TPL*EPT >= nr) and walks a contiguous block of rows.
// ------------------------------------------------------------------------------------------------
template <int TPL>
struct TeamGeom {
    static constexpr int CTA = (TPL < 256) ? 256 : TPL;
    // resident CTAs the register budget is tuned for: kernel A keeps three rows of the direction in registers
    // (2 CTAs of 256 threads, <= 128 regs); EPT = 8 doubles the per-thread state of both kernels
    static constexpr int minb_A(int ept) { return TPL <= 256 ? 2 : 1; }
    static constexpr int minb_B(int ept) { return TPL <= 256 ? (ept >= 8 ? 2 : 3) : 1; }
    static constexpr int LPC = CTA / TPL;
    static constexpr int NW = (TPL >= 32) ? TPL / 32 : 1;   // warps per team
    static constexpr int W = (TPL < 32) ? TPL : 32;         // lanes of a team inside one warp (sub-warp teams for nr <= 16)
};

// lanes of this thread's team inside its warp: shuffles of different teams in one warp must not wait for each other
template <int TPL>
__device__ __forceinline__ unsigned team_mask() {
    if constexpr (TPL >= 32) return 0xffffffffu;
    else return ((1u << TPL) - 1u) << ((threadIdx.x & 31) & ~(TPL - 1));
}

// ---- thread-block clusters: a row wider than one CTA's team (nr > 1024) is owned by CL CTAs of 256 threads; the scan
// carries of the line solve travel through distributed shared memory, one cluster barrier per direction.
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store v into the same shared-memory slot of CTA `cta` of this cluster
__device__ __forceinline__ void st_cluster_f64(double* local_slot, unsigned cta, double v) {
    unsigned la = (unsigned)__cvta_generic_to_shared(local_slot), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(cta));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}

template <int TPL, int CL = 1>
__device__ __forceinline__ void team_barrier(int team_in_cta) {
    if constexpr (CL > 1) {
        cluster_barrier();
    } else if constexpr (TPL > 32) {
        asm volatile("bar.sync %0, %1;" ::"r"(team_in_cta + 1), "r"(TPL) : "memory");
    } else {
        __syncwarp(team_mask<TPL>());
    }
}

// ------------------------------------------------------------------------------------------------
// z = M^-1 r on one row held by a team (EPT cells per thread): two affine first-order scans over the precomputed
// pivots,  y_i = r_i + (cR_{i-1} dinv_{i-1}) y_{i-1} ,  z_i = dinv_i y_i + (cR_i dinv_i) z_{i+1}.
// Thread-local composition, warp shuffle scan, <= 32 warp totals through shared memory (one team barrier per
// direction).  m0_edge = cR_{i0-1} dinv_{i0-1} for lane 0 of every warp but the first (fetched with the row loads).
// ------------------------------------------------------------------------------------------------
template <int TPL, int EPT, int CL = 1>
__device__ __forceinline__ void line_solve(const double (&rv)[EPT], const double (&dv)[EPT], const double (&cr)[EPT],
                                           double m0_edge, int lane, int wt, int team_in_cta, double* s_fA, double* s_fB,
                                           double* s_bA, double* s_bB, double (&zv)[EPT]) {
    constexpr int NW = TeamGeom<TPL>::NW * CL, W = TeamGeom<TPL>::W;       // wt = warp index inside the whole team
    const unsigned tm = team_mask<TPL>();
    double nb[EPT];     // cR_i * dinv_i : backward multiplier of cell i, forward multiplier of cell i+1
#pragma unroll
    for (int e = 0; e < EPT; ++e) nb[e] = cr[e] * dv[e];
    double m0 = __shfl_up_sync(tm, nb[EPT - 1], 1, W);
    if (lane == 0) m0 = m0_edge;
    double Pe[EPT], ye[EPT];
    {
        double P = 1.0, y = 0.0;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const double m = (e == 0) ? m0 : nb[e - 1];
            y = fma(m, y, rv[e]);
            P *= m;
            Pe[e] = P; ye[e] = y;
        }
    }
    double A = Pe[EPT - 1], B = ye[EPT - 1];
#pragma unroll
    for (int o = 1; o < W; o <<= 1) {
        const double Ap = __shfl_up_sync(tm, A, o, W);
        const double Bp = __shfl_up_sync(tm, B, o, W);
        if (lane >= o) { B = fma(A, Bp, B); A *= Ap; }
    }
    double Ax = __shfl_up_sync(tm, A, 1, W), Bx = __shfl_up_sync(tm, B, 1, W);
    if (lane == 0) { Ax = 1.0; Bx = 0.0; }
    double carry = 0.0;
    if constexpr (NW > 1) {
        if constexpr (CL > 1) {
            if (lane == W - 1) {
#pragma unroll
                for (unsigned c = 0; c < (unsigned)CL; ++c) { st_cluster_f64(&s_fA[wt], c, A); st_cluster_f64(&s_fB[wt], c, B); }
            }
        } else {
            if (lane == W - 1) { s_fA[wt] = A; s_fB[wt] = B; }
        }
        team_barrier<TPL, CL>(team_in_cta);
        if constexpr (NW >= 16) {
            // many warps per row (wide rows, cluster teams): a serial walk over up to NW-1 totals costs ~1000 cycles for the
            // last warp, twice per row; every warp scans the NW maps y -> A y + B with shuffles instead (5 steps)
            double ca = (lane < NW) ? s_fA[lane] : 1.0, cb = (lane < NW) ? s_fB[lane] : 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double ap = __shfl_up_sync(0xffffffffu, ca, o), bp = __shfl_up_sync(0xffffffffu, cb, o);
                if (lane >= o) { cb = fma(ca, bp, cb); ca *= ap; }
            }
            const double through = __shfl_sync(0xffffffffu, cb, wt > 0 ? wt - 1 : 0);
            carry = wt > 0 ? through : 0.0;
        } else {
            for (int w = 0; w < wt; ++w) carry = fma(s_fA[w], carry, s_fB[w]);
        }
    }
    const double ybefore = fma(Ax, carry, Bx);
#pragma unroll
    for (int e = 0; e < EPT; ++e) ye[e] = fma(Pe[e], ybefore, ye[e]);
    {
        double Q = 1.0, zz = 0.0;
#pragma unroll
        for (int e = EPT - 1; e >= 0; --e) {
            zz = fma(nb[e], zz, dv[e] * ye[e]);
            Q *= nb[e];
            Pe[e] = Q; zv[e] = zz;
        }
    }
    A = Pe[0]; B = zv[0];
#pragma unroll
    for (int o = 1; o < W; o <<= 1) {
        const double Ap = __shfl_down_sync(tm, A, o, W);
        const double Bp = __shfl_down_sync(tm, B, o, W);
        if (lane + o < W) { B = fma(A, Bp, B); A *= Ap; }
    }
    Ax = __shfl_down_sync(tm, A, 1, W); Bx = __shfl_down_sync(tm, B, 1, W);
    if (lane == W - 1) { Ax = 1.0; Bx = 0.0; }
    carry = 0.0;
    if constexpr (NW > 1) {
        if constexpr (CL > 1) {
            if (lane == 0) {
#pragma unroll
                for (unsigned c = 0; c < (unsigned)CL; ++c) { st_cluster_f64(&s_bA[wt], c, A); st_cluster_f64(&s_bB[wt], c, B); }
            }
        } else {
            if (lane == 0) { s_bA[wt] = A; s_bB[wt] = B; }
        }
        team_barrier<TPL, CL>(team_in_cta);
        if constexpr (NW >= 16) {
            // lane j holds the map of warp NW-1-j: the walk from the last warp down becomes an upward scan
            double ca = (lane < NW) ? s_bA[NW - 1 - lane] : 1.0, cb = (lane < NW) ? s_bB[NW - 1 - lane] : 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double ap = __shfl_up_sync(0xffffffffu, ca, o), bp = __shfl_up_sync(0xffffffffu, cb, o);
                if (lane >= o) { cb = fma(ca, bp, cb); ca *= ap; }
            }
            const double through = __shfl_sync(0xffffffffu, cb, wt < NW - 1 ? NW - 2 - wt : 0);
            carry = wt < NW - 1 ? through : 0.0;
        } else {
            for (int w = NW - 1; w > wt; --w) carry = fma(s_bA[w], carry, s_bB[w]);
        }
    }
    const double zafter = fma(Ax, carry, Bx);
#pragma unroll
    for (int e = 0; e < EPT; ++e) zv[e] = fma(Pe[e], zafter, zv[e]);
}

// ---- k_cg_A -------------------------------------------------------------------------------------
//   p_new = first ? z : z + beta*p_in          (rows ja .. jb+1 incl. halo rows: recomputed, never re-read)
//   q     = L p_new                             5-point, couplings cR / cZ, diagonal diag
//   pq    = sum p_new*q
// last CTA: alpha = rz/pq, parity ^= 1, first = 0
template <int TPL, int EPT, bool VEC, int MINB = TeamGeom<TPL>::minb_A(EPT)>
__global__ void __launch_bounds__(TeamGeom<TPL>::CTA, MINB) k_cg_A(HpDev d, hp_phys ph, HpKArgs a) {
    using G = TeamGeom<TPL>;
    const HpState s0 = *d.st;
    if (s0.done) return;
    const int nr = d.nr, nz = d.nz;
    const int team_in_cta = threadIdx.x / TPL, tl = threadIdx.x % TPL, lane = threadIdx.x & (TeamGeom<TPL>::W - 1);
    // teams tile the mesh as (row chunk) x (column tile); neighbours in another tile / warp come from global memory
    const long long team = (long long)blockIdx.x * G::LPC + team_in_cta, nteams = (long long)gridDim.x * G::LPC;
    const int ncol = d.ncolA;
    const long long nchunks = nteams / ncol;
    const long long chunk = team / ncol;
    const int i0 = ((int)(team % ncol) * TPL + tl) * EPT;
    int ja = 0, jb = 0;
    if (chunk < nchunks) { ja = (int)(chunk * nz / nchunks); jb = (int)((chunk + 1) * nz / nchunks); }
    const double beta = s0.beta;
    const bool first = s0.first != 0;
    const double* __restrict__ pin = s0.parity ? d.p1 : d.p0;
    double* __restrict__ pout = s0.parity ? d.p0 : d.p1;
    const double* __restrict__ zf = d.z;

    auto pnew_row = [&](int g, double (&out)[EPT]) {
        double zz[EPT];
        load_row<EPT, VEC>(zf + (long long)g * nr, i0, nr, zz);
        if (first) {
#pragma unroll
            for (int e = 0; e < EPT; ++e) out[e] = zz[e];
        } else {
            double pp[EPT];
            load_row<EPT, VEC>(pin + (long long)g * nr, i0, nr, pp);
#pragma unroll
            for (int e = 0; e < EPT; ++e) out[e] = fma(beta, pp[e], zz[e]);
        }
    };
    auto pnew_at = [&](long long idx) -> double {
        const double zz = zf[idx];
        return first ? zz : fma(beta, pin[idx], zz);
    };
    // radial neighbours that live in another warp: only lane 0 / lane 31 fetch them, one row AHEAD, together with
    // the row's vector loads (a dependent second round trip to DRAM per row would halve the streaming rate)
    const bool has_l = (lane == 0) && (i0 != 0) && (i0 < nr);
    const bool has_r = (lane == TeamGeom<TPL>::W - 1) && (i0 + EPT < nr);

    double acc[1] = {0.0};
    if (jb > ja) {
        double pm[EPT], pc[EPT], pn[EPT], czm[EPT];
        double pc_l = 0.0, pc_r = 0.0;
        pnew_row(ja, pm);          // row g = ja  (the row below the first owned row of this team)
        pnew_row(ja + 1, pc);
        if (has_l) pc_l = pnew_at((long long)(ja + 1) * nr + i0 - 1);
        if (has_r) pc_r = pnew_at((long long)(ja + 1) * nr + i0 + EPT);
        load_row<EPT, VEC>(d.cZ + (long long)ja * nr, i0, nr, czm);
        if (ja == 0) store_row<EPT, VEC>(pout, i0, nr, pm);                      // lower ghost row
        for (int g = ja + 1; g <= jb; ++g) {
            const long long rb = (long long)g * nr;
            // ---- every load of this row is issued before the first use ----
            double zz[EPT], pp[EPT], cr[EPT], cz[EPT], dg[EPT];
            load_row<EPT, VEC>(zf + rb + nr, i0, nr, zz);
            if (!first) load_row<EPT, VEC>(pin + rb + nr, i0, nr, pp);
            load_row<EPT, VEC>(d.cR + rb, i0, nr, cr);
            load_row<EPT, VEC>(d.cZ + rb, i0, nr, cz);
            load_row<EPT, VEC>(d.diag + rb, i0, nr, dg);
            double zl = 0.0, plo = 0.0, zr = 0.0, pro = 0.0, crl_e = 0.0;
            if (has_l) { zl = zf[rb + nr + i0 - 1]; if (!first) plo = pin[rb + nr + i0 - 1]; crl_e = d.cR[rb + i0 - 1]; }
            if (has_r) { zr = zf[rb + nr + i0 + EPT]; if (!first) pro = pin[rb + nr + i0 + EPT]; }
#pragma unroll
            for (int e = 0; e < EPT; ++e) pn[e] = first ? zz[e] : fma(beta, pp[e], zz[e]);
            const double pn_l = first ? zl : fma(beta, plo, zl);
            const double pn_r = first ? zr : fma(beta, pro, zr);
            // radial neighbours across thread boundaries
            const unsigned tm = team_mask<TPL>();
            double pl = __shfl_up_sync(tm, pc[EPT - 1], 1, G::W);
            double crl = __shfl_up_sync(tm, cr[EPT - 1], 1, G::W);
            double pr = __shfl_down_sync(tm, pc[0], 1, G::W);
            if (lane == 0) { pl = pc_l; crl = crl_e; }      // zero when there is no left neighbour
            if (lane == G::W - 1) pr = pc_r;
            double qv[EPT];
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const double left = (e == 0) ? pl : pc[e - 1];
                const double cleft = (e == 0) ? crl : cr[e - 1];
                const double right = (e == EPT - 1) ? pr : pc[(e + 1) % EPT];
                double t = dg[e] * pc[e];
                t = fma(-cr[e], right, t);
                t = fma(-cleft, left, t);
                t = fma(-cz[e], pn[e], t);
                t = fma(-czm[e], pm[e], t);
                qv[e] = t;
                acc[0] = fma(pc[e], t, acc[0]);
            }
            store_row<EPT, VEC>(d.q + rb, i0, nr, qv);
            store_row<EPT, VEC>(pout + rb, i0, nr, pc);
#pragma unroll
            for (int e = 0; e < EPT; ++e) { pm[e] = pc[e]; pc[e] = pn[e]; czm[e] = cz[e]; }
            pc_l = pn_l; pc_r = pn_r;
        }
        if (jb == nz) store_row<EPT, VEC>(pout + (long long)(nz + 1) * nr, i0, nr, pc);   // upper ghost row
    }
    double tot[1];
    if (cta_publish<1>(acc, 0u, d.partials, &d.st->ticket, tot, d.p2p != 0) && threadIdx.x == 0) {
        if (d.nranks > 1) stash_local<1>(d, tot, HP_COMM_A);
        else finalize_A(d, a, tot[0]);
    }
}

// ---- k_cg_B -------------------------------------------------------------------------------------
//   INIT = false:  x += alpha p ; r -= alpha q          INIT = true: r as assembled by k_diag
//   RLINE: z = M^-1 r by two affine first-order scans per row (forward, backward) over precomputed pivots:
//        y_i = r_i + (cR_{i-1} dinv_{i-1}) y_{i-1} ;  z_i = dinv_i y_i + (cR_i dinv_i) z_{i+1}
//   JACOBI: z = r * dinv
//   rz = sum r z ; rr = sum r r ; crit = max|z| (criterion 0) or max|r/diag| (criterion 1)
// last CTA: beta = rz_new/rz, iter++, convergence flag, sweep statistics, WHILE condition
template <int TPL, int EPT, bool VEC, bool RLINE, bool INIT>
__global__ void __launch_bounds__(TeamGeom<TPL>::CTA, TeamGeom<TPL>::minb_B(EPT)) k_cg_B(HpDev d, hp_phys ph, HpKArgs a) {
    using G = TeamGeom<TPL>;
    const HpState s0 = *d.st;
    if (!INIT && s0.done) return;
    if constexpr (INIT) {
        if (sweep_is_noop(d, a)) {      // z, r.z, r.r and the criterion are those of the sweep before: same decision
            if (blockIdx.x == 0 && threadIdx.x == 0) finalize_B(d, a, true, s0.rz, s0.rr, s0.crit, a.mg ? 1 : 0);
            return;
        }
    }
    __shared__ double s_fA[G::LPC][G::NW], s_fB[G::LPC][G::NW], s_bA[G::LPC][G::NW], s_bB[G::LPC][G::NW];
    const int nr = d.nr, nz = d.nz;
    const int team_in_cta = threadIdx.x / TPL, tl = threadIdx.x % TPL, lane = threadIdx.x & (TeamGeom<TPL>::W - 1);
    const int wt = tl >> 5;     // warp within team
    const int i0 = tl * EPT;
    const long long team = (long long)blockIdx.x * G::LPC + team_in_cta, nteams = (long long)gridDim.x * G::LPC;
    const int ja = (int)(team * nz / nteams), jb = (int)((team + 1) * nz / nteams);
    const double alpha = a.shift ? 1.0 : s0.alpha;
    const double* __restrict__ pcur = s0.parity ? d.p1 : d.p0;
    const bool need_diag = (a.criterion == 1);

    double acc[3] = {0.0, 0.0, 0.0};   // rz, rr, crit(max)
    for (int g = ja + 1; g <= jb; ++g) {
        const long long rb = (long long)g * nr;
        // ---- every load of this row is issued before the first store / use (the vector loads and stores are
        //      volatile asm and keep program order: a load placed after a store would wait for a full extra
        //      DRAM round trip) ----
        double rv[EPT], dv[EPT], zv[EPT];
        [[maybe_unused]] double xv[EPT], pv[EPT], qv[EPT], cr[EPT], dgv[EPT];
        load_row<EPT, VEC>(d.r + rb, i0, nr, rv);
        if constexpr (!INIT) {
            load_row<EPT, VEC>(d.T + rb, i0, nr, xv);
            load_row<EPT, VEC>(pcur + rb, i0, nr, pv);
            load_row<EPT, VEC>(d.q + rb, i0, nr, qv);
        }
        load_row<EPT, VEC>(d.dinv + rb, i0, nr, dv);
        if constexpr (RLINE) load_row<EPT, VEC>(d.cR + rb, i0, nr, cr);
        if (need_diag) load_row<EPT, VEC>(d.diag + rb, i0, nr, dgv);
        [[maybe_unused]] double m0_edge = 0.0;      // forward multiplier entering this warp from the previous one
        if constexpr (RLINE) {
            if (lane == 0 && tl != 0 && i0 < nr) m0_edge = d.cR[rb + i0 - 1] * d.dinv[rb + i0 - 1];
        }
        if constexpr (!INIT) {
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                xv[e] = fma(alpha, pv[e], xv[e]);
                rv[e] = fma(-alpha, qv[e], rv[e]);
            }
            store_row<EPT, VEC>(d.T + rb, i0, nr, xv);
            store_row<EPT, VEC>(d.r + rb, i0, nr, rv);
            if (d.p2p) {        // my boundary rows of the iterate -> the neighbours' ghost rows (NVLink stores)
                if (g == 1 && d.T_lo) store_row<EPT, VEC>(d.T_lo, i0, nr, xv);
                if (g == nz && d.T_hi) store_row<EPT, VEC>(d.T_hi, i0, nr, xv);
            }
        }
        if constexpr (!RLINE) {
#pragma unroll
            for (int e = 0; e < EPT; ++e) zv[e] = rv[e] * dv[e];
        } else {
            line_solve<TPL, EPT>(rv, dv, cr, m0_edge, lane, wt, team_in_cta, s_fA[team_in_cta], s_fB[team_in_cta],
                                 s_bA[team_in_cta], s_bB[team_in_cta], zv);
        }
        store_row<EPT, VEC>(d.z + rb, i0, nr, zv);
        if (d.p2p) {            // halo of the (line-)preconditioned residual, fused into its producer
            if (g == 1 && d.z_lo) store_row<EPT, VEC>(d.z_lo, i0, nr, zv);
            if (g == nz && d.z_hi) store_row<EPT, VEC>(d.z_hi, i0, nr, zv);
        }
        if (need_diag) {
#pragma unroll
            for (int e = 0; e < EPT; ++e)
                if (i0 + e < nr) acc[2] = fmax(acc[2], fabs(rv[e] / dgv[e]));
        } else {
#pragma unroll
            for (int e = 0; e < EPT; ++e) acc[2] = fmax(acc[2], fabs(zv[e]));
        }
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            acc[0] = fma(rv[e], zv[e], acc[0]);
            acc[1] = fma(rv[e], rv[e], acc[1]);
        }
    }
    double tot[3];
    if (cta_publish<3>(acc, 4u, d.partials, &d.st->ticket, tot, d.p2p != 0) && threadIdx.x == 0) {
        if (d.nranks > 1) stash_local<3>(d, tot, HP_COMM_B);
        else finalize_B(d, a, INIT || a.shift, tot[0], tot[1], tot[2], a.mg ? 1 : 0);
    }
}


// ---- k_cg_B_tma : k_cg_B with a TMA-staged row pipeline ----------------------------------------------------------------
// Same arithmetic as k_cg_B<TPL 256, EPT 4, RLINE, !INIT>.  The six operand rows of the NEXT row(s) (r, x, p, q, dinv, cR:
// 6 x nr x 8 bytes) are brought into shared memory by bulk asynchronous copies (cp.async.bulk global -> shared, completion
// on an mbarrier) issued by one thread, so they are in flight while the team runs the two scans and the two named barriers
// of the current row; in k_cg_B the loads of a row are issued at the top of its iteration and nothing is in flight during
// its line solve (ncu: long_scoreboard + barrier stalls, 0.82 of the copy peak).  Radial neighbours (the multiplier
// entering a warp) come from the staged row instead of an extra global load.  Results go out through registers as before.
constexpr int BT_STAGES = 2, BT_ARR = 6;
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void lds4(const double* p, double (&v)[4]) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(smem_u32(p)));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[2]), "=d"(v[3]) : "r"(smem_u32(p + 2)));
}

__global__ void __launch_bounds__(256, 2) k_cg_B_tma(HpDev d, hp_phys ph, HpKArgs a) {
    constexpr int TPL = 256, EPT = 4;
    using G = TeamGeom<TPL>;
    const HpState s0 = *d.st;
    if (s0.done) return;
    extern __shared__ __align__(128) unsigned char bt_smem[];
    const int nr = d.nr, nz = d.nz;
    const int rowp = (nr + 3) & ~3;                                  // staged row length (doubles)
    double* stage0 = reinterpret_cast<double*>(bt_smem);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(stage0 + (size_t)BT_STAGES * BT_ARR * rowp);
    __shared__ double s_fA[G::NW], s_fB[G::NW], s_bA[G::NW], s_bB[G::NW];
    const int tl = threadIdx.x, lane = tl & 31, wt = tl >> 5;
    const int i0 = tl * EPT;
    const long long team = blockIdx.x, nteams = gridDim.x;
    const int ja = (int)(team * nz / nteams), jb = (int)((team + 1) * nz / nteams);
    const double alpha = a.shift ? 1.0 : s0.alpha;
    const double* __restrict__ pcur = s0.parity ? d.p1 : d.p0;
    const unsigned row_bytes = (unsigned)nr * 8u;
    if (tl == 0) {
        for (int s = 0; s < BT_STAGES; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int g, int s) {             // one thread: the six operand rows of row g -> stage s
        const long long rb = (long long)g * nr;
        double* st = stage0 + (size_t)s * BT_ARR * rowp;
        mbar_expect_tx(&bars[s], BT_ARR * row_bytes);
        bulk_g2s(st + 0 * rowp, d.r + rb, row_bytes, &bars[s]);
        bulk_g2s(st + 1 * rowp, d.T + rb, row_bytes, &bars[s]);
        bulk_g2s(st + 2 * rowp, pcur + rb, row_bytes, &bars[s]);
        bulk_g2s(st + 3 * rowp, d.q + rb, row_bytes, &bars[s]);
        bulk_g2s(st + 4 * rowp, d.dinv + rb, row_bytes, &bars[s]);
        bulk_g2s(st + 5 * rowp, d.cR + rb, row_bytes, &bars[s]);
    };
    if (tl == 0)
        for (int k = 0; k < BT_STAGES && ja + 1 + k <= jb; ++k) issue(ja + 1 + k, k);

    double acc[3] = {0.0, 0.0, 0.0};   // rz, rr, crit(max)
    for (int g = ja + 1; g <= jb; ++g) {
        const int k = g - ja - 1, s = k % BT_STAGES;
        const long long rb = (long long)g * nr;
        const double* st = stage0 + (size_t)s * BT_ARR * rowp;
        mbar_wait(&bars[s], (unsigned)(k / BT_STAGES) & 1u);
        double rv[EPT], xv[EPT], pv[EPT], qv[EPT], dv[EPT], cr[EPT], zv[EPT];
        double m0_edge = 0.0;
        if (i0 < nr) {
            lds4(st + 0 * rowp + i0, rv); lds4(st + 1 * rowp + i0, xv); lds4(st + 2 * rowp + i0, pv);
            lds4(st + 3 * rowp + i0, qv); lds4(st + 4 * rowp + i0, dv); lds4(st + 5 * rowp + i0, cr);
            if (lane == 0 && tl != 0) m0_edge = st[5 * rowp + i0 - 1] * st[4 * rowp + i0 - 1];
        } else {
#pragma unroll
            for (int e = 0; e < EPT; ++e) { rv[e] = xv[e] = pv[e] = qv[e] = dv[e] = cr[e] = 0.0; }
        }
        // every thread has taken its operands out of the stage: refill it with the row BT_STAGES ahead
        asm volatile("bar.sync 15, 256;" ::: "memory");
        if (tl == 0 && g + BT_STAGES <= jb) issue(g + BT_STAGES, s);
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            xv[e] = fma(alpha, pv[e], xv[e]);
            rv[e] = fma(-alpha, qv[e], rv[e]);
        }
        store_row<EPT, true>(d.T + rb, i0, nr, xv);
        store_row<EPT, true>(d.r + rb, i0, nr, rv);
        if (d.p2p) {
            if (g == 1 && d.T_lo) store_row<EPT, true>(d.T_lo, i0, nr, xv);
            if (g == nz && d.T_hi) store_row<EPT, true>(d.T_hi, i0, nr, xv);
        }
        line_solve<TPL, EPT>(rv, dv, cr, m0_edge, lane, wt, 0, s_fA, s_fB, s_bA, s_bB, zv);
        store_row<EPT, true>(d.z + rb, i0, nr, zv);
        if (d.p2p) {
            if (g == 1 && d.z_lo) store_row<EPT, true>(d.z_lo, i0, nr, zv);
            if (g == nz && d.z_hi) store_row<EPT, true>(d.z_hi, i0, nr, zv);
        }
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            acc[2] = fmax(acc[2], fabs(zv[e]));
            acc[0] = fma(rv[e], zv[e], acc[0]);
            acc[1] = fma(rv[e], rv[e], acc[1]);
        }
    }
    double tot[3];
    if (cta_publish<3>(acc, 4u, d.partials, &d.st->ticket, tot, d.p2p != 0) && threadIdx.x == 0) {
        if (d.nranks > 1) stash_local<3>(d, tot, HP_COMM_B);
        else finalize_B(d, a, a.shift, tot[0], tot[1], tot[2], a.mg ? 1 : 0);
    }
}


// ================================================================================================
// k_pipe_step : one whole time step ([T.updateOld()] + `sweeps` sweeps, each = assembly + PCG with the radial-line
// preconditioner) of ONE heat pipe per CTA, everything on chip.          (main.py:200-202 / 267-271)
// The native Faghri mesh (500 x 9 = 4500 cells) and the members of a parameter ensemble are far too small to stream:
// through the row kernels a step is ~150 dependent launches of microseconds of work each (latency-bound), and an
// ensemble streams every vector through HBM once per PCG iteration.  Here thread j owns radial line j of its pipe:
//   registers      : x (= T), r, p, q / z of the line's NR cells
//   shared memory  : cR, cZ, T_old, diag, pivots of the whole pipe (column-major: [i][j], conflict-free) + one exchange
//                    buffer for the axial neighbours of p / T   (6 x NR x nz x 8 B = 211 KB for 500 x 9)
// so the line solve is a 9-step recurrence inside a thread (no shuffles), the stencil needs one shared-memory exchange,
// the dot products are block reductions, and nothing but the field is read or written in global memory per step.  Every
// pipe runs its own PCG to the tolerance (the stacked formulation shares the CG scalars over the ensemble and iterates
// until the slowest member is done).  Per-sweep statistics go to `part` [sweep][member][4] = res2, b2, crit, iters+flags;
// k_pipe_finalize folds them in member order (deterministic) into the statistics ring.
// ================================================================================================
template <int NS, int NM>
__device__ __forceinline__ void pipe_reduce(double (&v)[NS + NM], double (*s_red)[5][16], int& par) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NS + NM; ++k) {
        const double w = k < NS ? warp_sum(v[k]) : warp_max(v[k]);
        if (lane == 0) s_red[par][k][warp] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NS + NM; ++k) {
        double acc = 0.0;                                   // every max-reduced quantity is >= 0
        for (int w = 0; w < nwarp; ++w) acc = k < NS ? acc + s_red[par][k][w] : fmax(acc, s_red[par][k][w]);
        v[k] = acc;
    }
    par ^= 1;       // the next reduction writes the other buffer; the one after that is separated by a barrier
}

template <int NR>
__global__ void __launch_bounds__(512, 1) k_pipe_step(HpDev d, hp_phys ph, HpKArgs a, int nzm, int sweeps, int update_old,
                                                     int predict, int track, double* __restrict__ part) {
    extern __shared__ __align__(16) double pipe_sm[];
    __shared__ double s_red[2][5][16];
    const int NZP = nzm;
    double* s_cR = pipe_sm;
    double* s_cZ = s_cR + NR * NZP;
    double* s_To = s_cZ + NR * NZP;
    double* s_dg = s_To + NR * NZP;
    double* s_dv = s_dg + NR * NZP;
    double* s_x = s_dv + NR * NZP;
    const int m = blockIdx.x, M = gridDim.x, t = threadIdx.x;
    const bool act = t < nzm;
    const int g = m * nzm + (act ? t : 0) + 1;              // ghost-inclusive row of this thread's line
    const long long rb = (long long)g * NR;
    const unsigned zf = d.zc_flags[g];
    const double dzg = d.dz[g];
    int par = 0;
    double x[NR], r[NR], p[NR], q[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        x[i] = act ? d.T[rb + i] : 0.0;
        r[i] = p[i] = q[i] = 0.0;
        if (act) {
            s_cR[i * NZP + t] = d.cR[rb + i];
            s_cZ[i * NZP + t] = d.cZ[rb + i];
            const double told = d.Told[rb + i];
            // predict >= 1: Told still holds the field at the start of the PREVIOUS step, so x - Told is that step's
            // increment D: the first solve of this step starts from x + D (the system is still assembled at x);
            // predict == 2: d1 holds the increment of the step before, start from x + 2 D - d1 (quadratic in the step index)
            if (predict) {
                const double inc = x[i] - told;
                p[i] = (predict >= 2) ? 2.0 * inc - d.d1[rb + i] : inc;
                if (d.d1) d.d1[rb + i] = inc;
            }
            const double to = update_old ? x[i] : told;
            s_To[i * NZP + t] = to;
            if (update_old) d.Told[rb + i] = to;            // T.updateOld() (main.py:269), kept visible to the other entry points
        }
    }
    // z = M^-1 v on this thread's line (pivots in s_dv): forward / backward substitution of the LDL^T factors
    auto line_solve_local = [&](const double (&v)[NR], double (&z)[NR]) {
        double y[NR];
        y[0] = v[0];
#pragma unroll
        for (int i = 1; i < NR; ++i) y[i] = fma(s_cR[(i - 1) * NZP + t] * s_dv[(i - 1) * NZP + t], y[i - 1], v[i]);
        z[NR - 1] = s_dv[(NR - 1) * NZP + t] * y[NR - 1];
#pragma unroll
        for (int i = NR - 2; i >= 0; --i) {
            const double dv = s_dv[i * NZP + t];
            z[i] = fma(s_cR[i * NZP + t] * dv, z[i + 1], dv * y[i]);
        }
    };
    for (int s = 0; s < sweeps; ++s) {
        // ---- assembly at the current iterate (k_diag's arithmetic: diag_cell) ----
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NR; ++i) if (act) s_x[i * NZP + t] = x[i];
        __syncthreads();
        double red[5] = {0.0, 0.0, 0.0, 0.0, 0.0};          // res2, b2, rz, rr | crit
        if (act) {
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const double cE = s_cR[i * NZP + t], cW = i > 0 ? s_cR[(i - 1) * NZP + t] : 0.0;
                const double cN = s_cZ[i * NZP + t], cS = t > 0 ? s_cZ[i * NZP + t - 1] : 0.0;
                const double Tn = t + 1 < nzm ? s_x[i * NZP + t + 1] : 0.0, Ts = t > 0 ? s_x[i * NZP + t - 1] : 0.0;
                const double Te = i + 1 < NR ? x[i + 1 < NR ? i + 1 : i] : 0.0, Tw = i > 0 ? x[i > 0 ? i - 1 : 0] : 0.0;
                const double Tnb = cE * Te + cW * Tw + cN * Tn + cS * Ts;
                const double To = a.has_old ? s_To[i * NZP + t] : x[i];
                const DiagCell c = diag_cell(d, ph, a, g, i, zf, dzg, x[i], To, cE, cW, cN, cS, Tnb);
                s_dg[i * NZP + t] = c.dg;
                r[i] = c.r0;
                red[0] += c.r0 * c.r0;
                red[1] += c.rhs * c.rhs;
            }
            // pivots of the line (w_i = d_i - c_{i-1}^2 / w_{i-1})
            double prev = 0.0;
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const double inv = 1.0 / (s_dg[i * NZP + t] - prev);
                s_dv[i * NZP + t] = inv;
                const double c = s_cR[i * NZP + t];
                prev = c * c * inv;
            }
        }
        // predicted start (first sweep of a step): p already holds delta = last step's increment; the first pass of the loop
        // below then runs with alpha := 1 (x <- x + delta, r <- r0 - L delta) and ends like this INIT decision
        // Second sweep of a step (track bit1): the first solve of a step stops short of the step's final field by the
        // correction the re-sweeps add (property update, smooth in time); start from x + C of the PREVIOUS step (c2).
        if (s == 1 && (track & 2) && act) {
#pragma unroll
            for (int i = 0; i < NR; ++i) p[i] = d.c2[rb + i];
        }
        bool shift = (predict && s == 0) || (s == 1 && (track & 2));
        double rz = 0.0, rr = 0.0, crit = 0.0;
        int iters = 0, flags = 0;
        bool done = false;
        if (!shift) {
            if (act) {
                line_solve_local(r, q);
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    red[2] = fma(r[i], q[i], red[2]);
                    red[3] = fma(r[i], r[i], red[3]);
                    red[4] = fmax(red[4], a.criterion == 1 ? fabs(r[i] / s_dg[i * NZP + t]) : fabs(q[i]));
                }
            }
            pipe_reduce<4, 1>(red, s_red, par);
            rz = red[2]; rr = red[3]; crit = red[4];
            if (converged(a, crit, rr, red[1])) { done = true; flags |= 1; }
            else if (!isfinite(crit) || !isfinite(rz)) { done = true; flags |= 4; }
#pragma unroll
            for (int i = 0; i < NR; ++i) p[i] = q[i];
        } else {
            pipe_reduce<4, 1>(red, s_red, par);
        }
        const double res2 = red[0], b2 = red[1];
        while (!done) {
            // q = L p  (axial neighbours through shared memory), p.q
            __syncthreads();
#pragma unroll
            for (int i = 0; i < NR; ++i) if (act) s_x[i * NZP + t] = p[i];
            __syncthreads();
            double pq[1] = {0.0};
            if (act) {
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const double pn = t + 1 < nzm ? s_x[i * NZP + t + 1] : 0.0, ps = t > 0 ? s_x[i * NZP + t - 1] : 0.0;
                    const double cS = t > 0 ? s_cZ[i * NZP + t - 1] : 0.0;
                    double v = s_dg[i * NZP + t] * p[i];
                    if (i + 1 < NR) v = fma(-s_cR[i * NZP + t], p[i + 1 < NR ? i + 1 : i], v);
                    if (i > 0) v = fma(-s_cR[(i - 1) * NZP + t], p[i > 0 ? i - 1 : 0], v);
                    v = fma(-s_cZ[i * NZP + t], pn, v);
                    v = fma(-cS, ps, v);
                    q[i] = v;
                    pq[0] = fma(p[i], v, pq[0]);
                }
            }
            double alpha = 1.0;
            if (!shift) {
                pipe_reduce<1, 0>(pq, s_red, par);
                if (!(pq[0] > 0.0) || !isfinite(pq[0])) { flags |= 4; break; }      // breakdown
                alpha = rz / pq[0];
            }
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                x[i] = fma(alpha, p[i], x[i]);
                r[i] = fma(-alpha, q[i], r[i]);
            }
            double rd[3] = {0.0, 0.0, 0.0};                 // rz_new, rr | crit
            if (act) {
                line_solve_local(r, q);                      // q <- z
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    rd[0] = fma(r[i], q[i], rd[0]);
                    rd[1] = fma(r[i], r[i], rd[1]);
                    rd[2] = fmax(rd[2], a.criterion == 1 ? fabs(r[i] / s_dg[i * NZP + t]) : fabs(q[i]));
                }
            }
            pipe_reduce<2, 1>(rd, s_red, par);
            rr = rd[1]; crit = rd[2];
            if (!shift) ++iters;
            if (converged(a, crit, rr, b2)) { flags |= 1; break; }
            if (!isfinite(crit) || !isfinite(rd[0])) { flags |= 4; break; }
            if (iters >= a.max_iter) { flags |= 2; break; }
            const double beta = (!shift && rz != 0.0) ? rd[0] / rz : 0.0;     // after the shift pass the direction restarts: p = z
            rz = rd[0];
            shift = false;
#pragma unroll
            for (int i = 0; i < NR; ++i) p[i] = fma(beta, p[i], q[i]);
        }
        if (s == 0 && (track & 1) && act) {                 // first-sweep solution, kept for the correction formed at the end
#pragma unroll
            for (int i = 0; i < NR; ++i) d.d2[rb + i] = x[i];
        }
        if (t == 0) {
            double* o = part + ((size_t)s * M + m) * 4;
            o[0] = res2; o[1] = b2; o[2] = (a.criterion == 2) ? sqrt(rr) : crit; o[3] = (double)(iters + (flags << 24));
        }
    }
    if (act) {
#pragma unroll
        for (int i = 0; i < NR; ++i) d.T[rb + i] = x[i];
        if ((track & 1) && sweeps >= 2) {
#pragma unroll
            for (int i = 0; i < NR; ++i) d.c2[rb + i] = x[i] - d.d2[rb + i];
        }
    }
}

// member partials -> statistics ring and solver state (one CTA; fixed summation order)
__global__ void __launch_bounds__(256) k_pipe_finalize(HpDev d, HpKArgs a, int M, int sweeps, const double* __restrict__ part) {
    __shared__ double s_v[4][256];
    HpState* st = d.st;
    for (int s = 0; s < sweeps; ++s) {
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        int fl_and = 1, fl_or = 0, it = 0;
        for (int m = threadIdx.x; m < M; m += 256) {
            const double* o = part + ((size_t)s * M + m) * 4;
            v[0] += o[0]; v[1] += o[1]; v[2] = fmax(v[2], o[2]);
            const int packed = (int)o[3];
            const int f = packed >> 24, i = packed & 0xffffff;
            it = i > it ? i : it;
            fl_and &= (f & 1); fl_or |= (f & 6);
        }
        v[3] = (double)(it + ((fl_and | fl_or | 8) << 24));      // bit3 marks "this thread saw members" for the AND below
        if (threadIdx.x >= M) v[3] = -1.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) s_v[k][threadIdx.x] = v[k];
        __syncthreads();
        if (threadIdx.x == 0) {
            double res2 = 0.0, b2 = 0.0, crit = 0.0;
            int iters = 0, all_conv = 1, bad = 0;
            const int n = M < 256 ? M : 256;
            for (int k = 0; k < n; ++k) {
                res2 += s_v[0][k]; b2 += s_v[1][k]; crit = fmax(crit, s_v[2][k]);
                const int packed = (int)s_v[3][k];
                const int f = packed >> 24, i = packed & 0xffffff;
                iters = i > iters ? i : iters;
                all_conv &= (f & 1); bad |= (f & 6);
            }
            hp_sweep_stat* stt = &d.stats[st->sweep_count % HP_STATS_CAP];
            stt->residual_l2 = sqrt(res2);
            stt->rhs_l2 = sqrt(b2);
            stt->crit = crit;
            stt->iters = iters;
            stt->flags = (all_conv && !bad ? 1 : 0) | bad;
            st->sweep_count += 1;
            st->total_iters += iters;
            st->pipe_iters += iters;
            st->res2 = res2; st->b2 = b2; st->crit = crit; st->iter = iters; st->flags = (st->flags & 8) | stt->flags;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { st->done = 1; st->mg_skip = 1; st->noop_ok = 0; }
}

// ------------------------------------------------------------------------------------------------
// auxiliary kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_eval_property(hp_phys ph, int id, const double* __restrict__ T, double* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = hp::eval_property(ph, id, T[i]);
}

// rho, cp per owned cell; kavg = mean of k over the cell's 4 faces (main.py:243: k.value[cellFaceIDs] mean)
__global__ void k_cell_fields(HpDev d, hp_phys ph, double* __restrict__ rho_o, double* __restrict__ cp_o,
                              double* __restrict__ kavg_o) {
    const long long n = (long long)d.nz * d.nr;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(c / d.nr), i = (int)(c - (long long)j * d.nr), g = j + 1;
        const long long idx = c + d.nr;
        const double Tc = d.T[idx];
        double rho, cp;
        hp::rho_cp_region(ph, d.cell_band[i], d.zc_flags[g], Tc, rho, cp);
        if (rho_o) rho_o[c] = rho;
        if (cp_o) cp_o[c] = cp;
        if (kavg_o) {
            // bottom / top axial faces (exterior ones: T_f = T_P, masked law at the face position)
            double kb = d.axial_open[g - 1] ? k_zface(d, ph, g - 1, i, d.T[idx - d.nr], Tc)
                                            : hp::k_region(ph, d.zface_band[i], d.zv_flags[g - 1], Tc);
            double kt = d.axial_open[g] ? k_zface(d, ph, g, i, Tc, d.T[idx + d.nr])
                                        : hp::k_region(ph, d.zface_band[i], d.zv_flags[g], Tc);
            double kr = (i < d.nr - 1) ? k_rface(d, ph, g, i, Tc, d.T[idx + 1]) : k_outface(d, ph, g, Tc);
            double kl = (i > 0) ? k_rface(d, ph, g, i - 1, d.T[idx - 1], Tc)
                                : hp::k_region(ph, HP_BAND_VC, d.zc_flags[g], Tc);   // axis face r = 0 <= R_vc
            if (ph.face_k != 0) {
                // harmonic mode: exterior faces take the cell value
                if (!d.axial_open[g - 1]) kb = hp::k_region(ph, d.cell_band[i], d.zc_flags[g], Tc);
                if (!d.axial_open[g]) kt = hp::k_region(ph, d.cell_band[i], d.zc_flags[g], Tc);
                if (i == 0) kl = hp::k_region(ph, d.cell_band[i], d.zc_flags[g], Tc);
            }
            kavg_o[c] = (kb + kr + kt + kl) / 4.0;
        }
    }
}

// rhs (owned rows, compact nz*nr) from a ghost-inclusive scratch array
__global__ void k_copy_owned(const double* __restrict__ src_ghost, double* __restrict__ dst, int nr, long long n) {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x)
        dst[c] = src_ghost[c + nr];
}

// field[g,:] = Tamb_line[g] for all rows incl. ghosts  (main.py:34,152)
__global__ void k_fill_rows(HpDev d, double* __restrict__ field) {
    const long long n = (long long)(d.nz + 2) * d.nr;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x)
        field[c] = d.Tamb_line[(int)(c / d.nr)];
}

// Start of a time step with an extrapolated first solve: with D_n = T - Told the increment of the step just finished and
// d1 = D_{n-1}, d2 = D_{n-2} those of the steps before,
//   z = D_n                       (order 1: x0 = 2 T_n - T_{n-1})
//   z = 2 D_n - d1                (order 2: x0 = 3 T_n - 3 T_{n-1} + T_{n-2})
//   z = 3 D_n - 3 d1 + d2         (order 3: x0 = 4 T_n - 6 T_{n-1} + 4 T_{n-2} - T_{n-3})
// is the predicted increment of the next step (the system itself is still assembled at T, like FiPy's; the prediction only
// seeds the PCG).  The history shifts (d2 <- d1 <- D_n) and Told <- T (`T.updateOld()`, main.py:269) in the same pass.
// With re-sweeps the first solve of a step does not produce the step's final field F_n but S1_n, short of it by the
// correction C_n = F_n - S1_n the later sweeps add (property update; ~1e-6 K, smooth in time).  k_stash keeps S1_n in c2;
// here C_n is formed, the history is built from the first-sweep increments D_n = S1_n - F_{n-1} (so the prediction aims
// at S1_{n+1}, what the first solve actually converges to), and c2 <- C_n is the predicted start of the second sweep.
__global__ void __launch_bounds__(256) k_delta(HpDev d, hp_phys ph, HpKArgs a) {
    const long long n = (long long)(d.nz + 2) * d.nr;
    const int ord = a.ext_order;
    if (blockIdx.x == 0 && threadIdx.x == 0) d.st->noop_ok = 0;         // T_old moves
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const double Tc = d.T[c];
        double dn = Tc - d.Told[c];
        if (d.c2 && a.c2_valid) {
            const double cn = Tc - d.c2[c];
            d.c2[c] = cn;
            dn -= cn;
        }
        double z = dn;
        if (d.d1) {
            const double p1 = d.d1[c];
            if (ord >= 2) z = 2.0 * dn - p1;
            if (ord == 3) {
                // the cubic term = third difference of the history.  Every stored field carries solver noise of the order of
                // the tolerance and the third difference amplifies it 8x: once the transient is smooth enough for the term to
                // sink to that floor it only injects noise (erratic iteration counts late in a run), so it is applied
                // cell by cell with a weight that rises from 0 at the floor (ext_guard) to 1 at twice the floor
                // (a taper, not a switch: a hard threshold puts steps of the size of the guard into the start field
                // and makes the start depend discontinuously on round-off differences between decompositions)
                const double d3 = dn - 2.0 * p1 + d.d2[c];
                const double w = a.ext_guard > 0.0 ? fmin(1.0, fmax(0.0, (fabs(d3) - a.ext_guard) / a.ext_guard)) : 1.0;
                z = fma(w, d3, z);
            }
            d.d2[c] = p1;
            d.d1[c] = dn;
        }
        d.z[c] = z;
        d.Told[c] = Tc;
    }
}

// head of the second sweep of a step (T = S1_n, every row incl. ghosts): z <- C_{n-1} as the predicted increment of this
// sweep (applied by the shift pair), c2 <- S1_n for k_delta of the next step
__global__ void __launch_bounds__(256) k_stash(HpDev d, hp_phys ph, HpKArgs a) {
    const long long n = (long long)(d.nz + 2) * d.nr;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        if (a.c2_valid) d.z[c] = d.c2[c];
        d.c2[c] = d.T[c];
    }
}

// ================================================================================================
// 'mgline' preconditioner: V(1,1) cycle, axial semi-coarsening (row pairs), damped radial-line Jacobi smoother.
// The radial stiffness is removed exactly by the line solves on every level; coarsening the axial index removes the
// axial stiffness that makes plain line-PCG need O(sqrt(D)) iterations (D = axial diffusion number of a cell).
// Galerkin operators with piecewise-constant aggregation stay 5-point, so every level reuses the same row kernels.
// Under peer-memory slabs the hierarchy is the single-domain one, distributed: every level keeps its couplings to the
// neighbouring slabs, every stage hands its two boundary rows to the neighbours (HpXchg: peer stores + a sequence flag, fused
// into the producing row team) and the first / last row team of the consuming stage waits for them.  Without peer memory
// (NCCL slabs, d.mg_local) the cycle ignores the ghost couplings: no communication in the preconditioner.
// ================================================================================================

// the team that owns the first (lo) / last (hi) rows of a stage waits for the rows its neighbours pushed in stage x.wait
template <int TPL, int CL = 1>
__device__ __forceinline__ void stage_wait(const HpDev& d, const HpKArgs& a, const HpXchg& x, bool lo, bool hi, int tl,
                                           int team_in_cta, unsigned long long cyc) {
    if (x.wait < 0 || !d.p2p) return;
    lo = lo && d.rank > 0;
    hi = hi && d.rank < d.nranks - 1;
    if (!(lo || hi)) return;                      // uniform over the team
    if (tl == 0 && !(d.st->flags & 8)) {
        bool ok = true;
        if (lo) ok = wait_peer_flag(d, a, x.wait, d.rank - 1, cyc);
        if (hi && ok) ok = wait_peer_flag(d, a, x.wait, d.rank + 1, cyc);
        if (!ok) raise_fault(d);
    }
    team_barrier<TPL, CL>(team_in_cta);
    __threadfence_system();                       // every reader of the ghost rows orders its loads after the acquire
}

// boundary row of a stage's output -> the neighbour's ghost row, then the stage flag (release, system scope)
template <int TPL, int EPT, bool VEC, int CL = 1>
__device__ __forceinline__ void stage_push(const HpDev& d, const HpXchg& x, bool is_lo, double* dst, int i0, int nr,
                                           const double (&v)[EPT], int tl, int team_in_cta, unsigned long long cyc) {
    store_row<EPT, VEC>(dst, i0, nr, v);
    if (x.sig < 0) return;
    __threadfence_system();
    team_barrier<TPL, CL>(team_in_cta);
    if (tl == 0) {
        unsigned long long* f = &d.comm_peer[is_lo ? d.rank - 1 : d.rank + 1]->flag[x.sig][d.rank];
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(cyc) : "memory");
    }
}

// coarse level from a fine one (rows aggregated in pairs, P = row-pair injection):
//   cR_c = cR[2J] + cR[2J+1]                                   radial couplings: Galerkin (P^T A P)
//   cZ_c = zs * (coupling of the pair's top row to the next pair)
//   diag_c = diag[2J] + diag[2J+1] - 2 cZ[2J] - (1 - zs) (cZ_up + cZ_down)
// zs = 1 is the Galerkin operator.  With piecewise-constant aggregation its axial coupling is twice what a discretisation
// on the coarse rows (twice the centre distance) gives, which halves the coarse correction of smooth axial modes; zs = 1/2
// restores it (row sums = capacity + boundary terms are kept, the operator stays a symmetric M-matrix and keeps the
// 5-point shape): 20 -> 7 PCG iterations on the 4096-row mesh.  Row 0 of cZ_c is the coupling to the slab below
// (keep_ghost: distributed hierarchy under peer-memory slabs; zero at a physical end anyway).     thread per coarse cell
__global__ void __launch_bounds__(256) k_mg_coarsen(HpLevel f, HpLevel c, int nr, double zs, int keep_ghost, const int* skip) {
    if (*skip) return;
    const long long n = (long long)(c.nz + 1) * nr;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int J = (int)(t / nr) - 1, i = (int)(t - (long long)(J + 1) * nr);
        if (J < 0) {                                   // ghost coupling below the first pair
            c.cZ[i] = keep_ghost ? zs * f.cZ[i] : 0.0;
            continue;
        }
        const long long g0 = (long long)(2 * J + 1) * nr + i, g1 = g0 + nr;
        const bool two = (2 * J + 1) < f.nz;
        const long long gc = (long long)(J + 1) * nr + i;
        const double inner = two ? f.cZ[g0] : 0.0;
        const double up = two ? f.cZ[g1] : f.cZ[g0], dn = f.cZ[g0 - nr];
        c.cR[gc] = f.cR[g0] + (two ? f.cR[g1] : 0.0);
        c.cZ[gc] = (J == c.nz - 1 && !keep_ghost) ? 0.0 : zs * up;
        c.diag[gc] = f.diag[g0] + (two ? f.diag[g1] - 2.0 * inner : 0.0) - (1.0 - zs) * (up + dn);
    }
}

// ---- k_mg_res : residual of one level, optionally with the coarse correction prolonged on load and / or the
// residual restricted to the next level on store.  Same marching, column-tiled, barrier-free structure as k_cg_A.
//   x(g)   = scale * xin(g) [+ ec(pair of g)]                          (PROLONG)
//   res(g) = rhs(g) - (diag x - cR x_e - cR_w x_w - cZ x_n - cZ_s x_s) (couplings to ghost rows ignored)
//   xout(g) = x(g) if xout;  resout(g) = res(g) if resout;  crhs(J) = res(2J) + res(2J+1) if crhs (RESTRICT)
// MODE 1 fuses the post-smoothing step into the same pass:  xout(g) = x(g) + omega * M_L^-1 res(g)  (the line solve of
// row g only needs the residual of row g).  MODE 2 fuses the pre-smoothing of the NEXT level into the restriction:
// c_x(J) = omega * M_{L+1}^-1 crhs(J) as soon as the coarse row J = pair-sum of residual rows is complete.  Both need the
// team to own whole rows (no column tiles).
// MODE 3 = MODE 1 on the fine level at the END of the cycle: the output is the preconditioned residual z of the PCG; the
// pass also accumulates r.z (returned; the kernel publishes it and advances beta / rz) -- one launch instead of the residual
// pass + the final line solve (48 B/cell less per iteration: the residual and the prolonged iterate never go to memory).
template <int TPL, int EPT, bool VEC, int MODE = 0, int CL = 1>
__device__ __forceinline__ double mg_res_body(const HpDev& d, const HpKArgs& a, const HpLevel& L, const HpMgResArgs& m, long long team,
                                            long long nteams, double omega = 0.0, double* s_fA = nullptr,
                                            double* s_fB = nullptr, double* s_bA = nullptr, double* s_bB = nullptr) {
    const int nr = d.nr, nz = L.nz;
    // CL > 1: the team is a cluster of CL CTAs (TPL threads each); tl runs over the whole team
    const int tl = (CL > 1 ? (int)cluster_ctarank() * TPL : 0) + threadIdx.x % TPL, lane = threadIdx.x & (TeamGeom<TPL>::W - 1);
    const int team_in_cta = threadIdx.x / TPL;
    const unsigned long long cyc = d.st->cycle;
    const int ncol = (MODE == 0) ? d.ncolA : 1;            // the fused stages own whole rows
    const long long nchunks = nteams / ncol;
    const long long chunk = team / ncol;
    const int i0 = ((int)(team % ncol) * TPL * CL + tl) * EPT;
    // chunks are made of whole row pairs so that a restricted row is produced by one team
    const long long npairs = (nz + 1) / 2;
    int ja = 0, jb = 0;
    if (chunk < nchunks) {
        // Under peer-memory slabs the first / last row pair gets a chunk of its own when the level is tall enough: that
        // team consumes the neighbour's boundary row and produces this slab's one within two rows of work, so the
        // hand-over chain between the slabs runs ahead of the interior teams instead of adding its NVLink latency to
        // every stage (an edge team that marches a full chunk produces its far boundary row at the END of the kernel and
        // the neighbour's next stage needs it at its START).
        const long long e0 = (d.p2p && d.rank > 0 && (m.xc.wait >= 0 || m.xc.push_lo) && nchunks >= 4 && npairs >= 4 * nchunks) ? 1 : 0;
        const long long e1 = (d.p2p && d.rank < d.nranks - 1 && (m.xc.wait >= 0 || m.xc.push_hi) && nchunks >= 4 && npairs >= 4 * nchunks) ? 1 : 0;
        long long pa, pb;
        if (e0 && chunk == 0) { pa = 0; pb = 1; }
        else if (e1 && chunk == nchunks - 1) { pa = npairs - 1; pb = npairs; }
        else {
            const long long ci = chunk - e0, Ci = nchunks - e0 - e1, Pi = npairs - e0 - e1;
            pa = e0 + ci * Pi / Ci;
            pb = e0 + (ci + 1) * Pi / Ci;
        }
        ja = (int)(2 * pa);
        jb = (int)(2 * pb);
        if (jb > nz) jb = nz;
    }
    const bool has_l = (lane == 0) && (i0 != 0) && (i0 < nr);
    const bool has_r = (lane == TeamGeom<TPL>::W - 1) && (i0 + EPT < nr);
    const bool prolong = m.ec != nullptr;
    const double scale = m.scale;
    constexpr bool UP = (MODE == 1 || MODE == 3);
    double rz_acc = 0.0;
    // coarse ghost-inclusive row of fine row g: the ghost rows 0 / nz+1 map to the coarse ghost rows 0 / nzc+1 (the
    // neighbouring slab's boundary pair; never-written zeros at a physical end, where the coupling is zero)
    auto crow = [&](int g) { return (long long)(((g - 1) >> 1) + 1) * nr; };
    auto xrow = [&](int g, double (&out)[EPT]) {
        load_row<EPT, VEC>(m.xin + (long long)g * nr, i0, nr, out);
        if (prolong) {
            double e[EPT];
            load_row<EPT, VEC>(m.ec + crow(g), i0, nr, e);
#pragma unroll
            for (int k = 0; k < EPT; ++k) out[k] = fma(scale, out[k], e[k]);
        } else {
#pragma unroll
            for (int k = 0; k < EPT; ++k) out[k] *= scale;
        }
    };
    auto xat = [&](int g, int i) -> double {
        double v = scale * m.xin[(long long)g * nr + i];
        if (prolong) v += m.ec[crow(g) + i];
        return v;
    };
    if (jb > ja) {
        stage_wait<TPL, CL>(d, a, m.xc, ja == 0, jb == nz, tl, team_in_cta, cyc);
        double xm[EPT], xc[EPT], xn[EPT], czm[EPT], rprev[EPT];
        double xc_l = 0.0, xc_r = 0.0;
        xrow(ja, xm);
        xrow(ja + 1, xc);
        if (has_l) xc_l = xat(ja + 1, i0 - 1);
        if (has_r) xc_r = xat(ja + 1, i0 + EPT);
        load_row<EPT, VEC>(L.cZ + (long long)ja * nr, i0, nr, czm);
        if (ja == 0) {
            if (d.mg_local) {
#pragma unroll
                for (int k = 0; k < EPT; ++k) czm[k] = 0.0;    // rank-local cycle: no coupling to the lower ghost row
            }
            if (m.xout_ghost) store_row<EPT, VEC>(m.xout, i0, nr, xm);                  // ghost row of the scaled iterate
        }
#pragma unroll
        for (int k = 0; k < EPT; ++k) rprev[k] = 0.0;
        for (int g = ja + 1; g <= jb; ++g) {
            const long long rb = (long long)g * nr;
            double cr[EPT], cz[EPT], dg[EPT], rh[EPT];
            [[maybe_unused]] double dv[EPT];
            xrow(g + 1, xn);
            load_row<EPT, VEC>(L.cR + rb, i0, nr, cr);
            load_row<EPT, VEC>(L.cZ + rb, i0, nr, cz);
            load_row<EPT, VEC>(L.diag + rb, i0, nr, dg);
            load_row<EPT, VEC>(L.rhs + rb, i0, nr, rh);
            [[maybe_unused]] double m0_edge = 0.0;
            if constexpr (UP) {
                load_row<EPT, VEC>(L.dinv + rb, i0, nr, dv);
                if (lane == 0 && tl != 0 && i0 < nr) m0_edge = L.cR[rb + i0 - 1] * L.dinv[rb + i0 - 1];
            }
            [[maybe_unused]] double cdv[EPT], ccr[EPT];
            const bool second = ((g - 1) & 1) != 0;
            const bool crow_done = second || g == nz;                 // coarse row J = (g-1)/2 + 1 completes with this row
            const long long cb = (long long)(((g - 1) >> 1) + 1) * nr;
            if constexpr (MODE == 2) {
                if (crow_done) {
                    load_row<EPT, VEC>(m.c_dinv + cb, i0, nr, cdv);
                    load_row<EPT, VEC>(m.c_cR + cb, i0, nr, ccr);
                    if (lane == 0 && tl != 0 && i0 < nr) m0_edge = m.c_cR[cb + i0 - 1] * m.c_dinv[cb + i0 - 1];
                }
            }
            double xn_l = 0.0, xn_r = 0.0, crl_e = 0.0;
            if (has_l) { xn_l = xat(g + 1, i0 - 1); crl_e = L.cR[rb + i0 - 1]; }
            if (has_r) xn_r = xat(g + 1, i0 + EPT);
            if (g == nz) {
                if (d.mg_local) {
#pragma unroll
                    for (int k = 0; k < EPT; ++k) cz[k] = 0.0; // rank-local cycle: no coupling to the upper ghost row
                }
                if (m.xout_ghost) store_row<EPT, VEC>(m.xout + rb + nr, i0, nr, xn);
            }
            const unsigned tm = team_mask<TPL>();
            constexpr int W = TeamGeom<TPL>::W;
            double xl = __shfl_up_sync(tm, xc[EPT - 1], 1, W);
            double crl = __shfl_up_sync(tm, cr[EPT - 1], 1, W);
            double xr = __shfl_down_sync(tm, xc[0], 1, W);
            if (lane == 0) { xl = xc_l; crl = crl_e; }
            if (lane == W - 1) xr = xc_r;
            double res[EPT];
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                const double left = (k == 0) ? xl : xc[k - 1];
                const double cleft = (k == 0) ? crl : cr[k - 1];
                const double right = (k == EPT - 1) ? xr : xc[(k + 1) % EPT];
                double t = dg[k] * xc[k];
                t = fma(-cr[k], right, t);
                t = fma(-cleft, left, t);
                t = fma(-cz[k], xn[k], t);
                t = fma(-czm[k], xm[k], t);
                res[k] = rh[k] - t;
            }
            if constexpr (UP) {
                double zv[EPT];
                line_solve<TPL, EPT, CL>(res, dv, cr, m0_edge, lane, tl >> 5, team_in_cta, s_fA, s_fB, s_bA, s_bB, zv);
#pragma unroll
                for (int k = 0; k < EPT; ++k) zv[k] = fma(omega, zv[k], xc[k]);
                store_row<EPT, VEC>(m.xout + rb, i0, nr, zv);
                if constexpr (MODE == 3) {
#pragma unroll
                    for (int k = 0; k < EPT; ++k) rz_acc = fma(rh[k], zv[k], rz_acc);      // rhs of level 0 is the PCG residual r
                }
                if (g == 1 && m.xc.push_lo) stage_push<TPL, EPT, VEC, CL>(d, m.xc, true, m.xc.push_lo, i0, nr, zv, tl, team_in_cta, cyc);
                if (g == nz && m.xc.push_hi) stage_push<TPL, EPT, VEC, CL>(d, m.xc, false, m.xc.push_hi, i0, nr, zv, tl, team_in_cta, cyc);
            } else {
                if (m.xout) store_row<EPT, VEC>(m.xout + rb, i0, nr, xc);
                if (m.resout) store_row<EPT, VEC>(m.resout + rb, i0, nr, res);
            }
            if (m.crhs) {
                if (crow_done) {
                    double sum[EPT];
#pragma unroll
                    for (int k = 0; k < EPT; ++k) sum[k] = second ? rprev[k] + res[k] : res[k];
                    store_row<EPT, VEC>(m.crhs + cb, i0, nr, sum);
                    if constexpr (MODE == 2) {
                        double zc[EPT];
                        line_solve<TPL, EPT, CL>(sum, cdv, ccr, m0_edge, lane, tl >> 5, team_in_cta, s_fA, s_fB, s_bA, s_bB, zc);
#pragma unroll
                        for (int k = 0; k < EPT; ++k) zc[k] *= omega;
                        store_row<EPT, VEC>(m.c_x + cb, i0, nr, zc);
                        const int Jg = ((g - 1) >> 1) + 1;
                        if (Jg == 1 && m.xc.push_lo) stage_push<TPL, EPT, VEC, CL>(d, m.xc, true, m.xc.push_lo, i0, nr, zc, tl, team_in_cta, cyc);
                        if (Jg == m.nzc && m.xc.push_hi) stage_push<TPL, EPT, VEC, CL>(d, m.xc, false, m.xc.push_hi, i0, nr, zc, tl, team_in_cta, cyc);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < EPT; ++k) rprev[k] = res[k];
                }
            }
#pragma unroll
            for (int k = 0; k < EPT; ++k) { xm[k] = xc[k]; xc[k] = xn[k]; czm[k] = cz[k]; }
            xc_l = xn_l; xc_r = xn_r;
        }
    }
    return rz_acc;
}

// end of the V-cycle: r.z of this CTA -> grid sum -> beta / rz recurrences (finalize_B stage 2), or this rank's mailbox slots
__device__ __forceinline__ void mg_final_publish(const HpDev& d, const HpKArgs& a, double (&acc)[1]) {
    double tot[1];
    if (cta_publish<1>(acc, 0u, d.partials, &d.st->ticket, tot, d.p2p != 0) && threadIdx.x == 0) {
        if (d.nranks > 1) stash_local<1>(d, tot, HP_COMM_MG);
        else finalize_B(d, a, a.init_phase != 0, tot[0], 0.0, 0.0, 2);
    }
}

template <int TPL, int EPT, bool VEC>
__global__ void __launch_bounds__(TeamGeom<TPL>::CTA, TeamGeom<TPL>::minb_A(EPT))
k_mg_res(HpDev d, HpKArgs a, HpLevel L, HpMgResArgs m) {
    using G = TeamGeom<TPL>;
    if (d.st->mg_skip) return;
    (void)mg_res_body<TPL, EPT, VEC>(d, a, L, m, (long long)blockIdx.x * G::LPC + threadIdx.x / TPL, (long long)gridDim.x * G::LPC);
}

// ---- k_mg_up : prolong + residual + post-smoothing of one coarse level in ONE pass (one stage instead of two in the
// latency-bound coarse part of the cycle):  xout = xt + omega M_L^-1 (rhs - L xt),  xt = xin + P ec
template <int TPL, int EPT, bool VEC, int MODE>
__global__ void __launch_bounds__(TeamGeom<TPL>::CTA, TeamGeom<TPL>::minb_A(EPT))
k_mg_fused(HpDev d, HpKArgs a, HpLevel L, HpMgResArgs m) {
    using G = TeamGeom<TPL>;
    if (d.st->mg_skip) return;
    __shared__ double s_fA[G::LPC][G::NW], s_fB[G::LPC][G::NW], s_bA[G::LPC][G::NW], s_bB[G::LPC][G::NW];
    const int tic = threadIdx.x / TPL;
    double acc[1];
    acc[0] = mg_res_body<TPL, EPT, VEC, MODE>(d, a, L, m, (long long)blockIdx.x * G::LPC + tic, (long long)gridDim.x * G::LPC,
                                              a.omega, s_fA[tic], s_fB[tic], s_bA[tic], s_bB[tic]);
    if constexpr (MODE == 3) mg_final_publish(d, a, acc);
}
// the same stages for rows of 1024 < nr <= 4096 cells: a cluster of CL CTAs (256 threads x 4 cells each) owns a row.  The
// cluster shape is a compile-time attribute, so plain launches and graph kernel nodes need no launch attribute.
#define HP_CLUSTER_FUSED(CLN)                                                                                          \
    template <bool VEC, int MODE>                                                                                      \
    __global__ void __cluster_dims__(CLN, 1, 1) __launch_bounds__(256, 2)                                               \
    k_mg_fused_cl##CLN(HpDev d, HpKArgs a, HpLevel L, HpMgResArgs m) {                                                 \
        if (d.st->mg_skip) return;                                                                                     \
        __shared__ double s_fA[8 * CLN], s_fB[8 * CLN], s_bA[8 * CLN], s_bB[8 * CLN];                                  \
        double acc[1];                                                                                                 \
        acc[0] = mg_res_body<256, 4, VEC, MODE, CLN>(d, a, L, m, (long long)(blockIdx.x / CLN),                        \
                                                     (long long)(gridDim.x / CLN), a.omega, s_fA, s_fB, s_bA, s_bB);   \
        if constexpr (MODE == 3) mg_final_publish(d, a, acc);                                                          \
    }
HP_CLUSTER_FUSED(2)
HP_CLUSTER_FUSED(4)
#undef HP_CLUSTER_FUSED
static const void* fn_mg_fused_cluster(int cl, bool vec, int mode) {
    if (cl == 2) {
        if (mode == 3) return vec ? (const void*)k_mg_fused_cl2<true, 3> : (const void*)k_mg_fused_cl2<false, 3>;
        if (mode == 1) return vec ? (const void*)k_mg_fused_cl2<true, 1> : (const void*)k_mg_fused_cl2<false, 1>;
        return vec ? (const void*)k_mg_fused_cl2<true, 2> : (const void*)k_mg_fused_cl2<false, 2>;
    }
    if (mode == 3) return vec ? (const void*)k_mg_fused_cl4<true, 3> : (const void*)k_mg_fused_cl4<false, 3>;
    if (mode == 1) return vec ? (const void*)k_mg_fused_cl4<true, 1> : (const void*)k_mg_fused_cl4<false, 1>;
    return vec ? (const void*)k_mg_fused_cl4<true, 2> : (const void*)k_mg_fused_cl4<false, 2>;
}

template <int TPL, int EPT, bool VEC>
const void* fn_mg_up() { return (const void*)k_mg_fused<TPL, EPT, VEC, 1>; }
template <int TPL, int EPT, bool VEC>
const void* fn_mg_final() { return (const void*)k_mg_fused<TPL, EPT, VEC, 3>; }
template <int TPL, int EPT, bool VEC>
const void* fn_mg_down() { return (const void*)k_mg_fused<TPL, EPT, VEC, 2>; }

// ---- k_mg_line : out(g) = [base(g) +] omega * M_L^-1 rhs(g)       (row teams, same line solve as k_cg_B)
// FINAL (level 0, end of the V-cycle): out is the preconditioned residual z of the PCG; accumulates r.z, pushes the
// boundary rows under peer-memory slabs, and the last CTA advances beta / rz (finalize_B stage 2).
template <int TPL, int EPT, bool VEC>
__device__ __forceinline__ double mg_line_body(const HpDev& d, double omega, const HpLevel& L, const HpMgLineArgs& m,
                                               long long team, long long nteams, double* s_fA, double* s_fB, double* s_bA,
                                               double* s_bB) {
    const int nr = d.nr, nz = L.nz;
    const int team_in_cta = threadIdx.x / TPL, tl = threadIdx.x % TPL, lane = threadIdx.x & (TeamGeom<TPL>::W - 1);
    const int wt = tl >> 5;
    const int i0 = tl * EPT;
    int ja = 0, jb = 0;
    if (team < nteams) { ja = (int)(team * nz / nteams); jb = (int)((team + 1) * nz / nteams); }
    double acc = 0.0;
    const unsigned long long cyc = d.st->cycle;
    for (int g = ja + 1; g <= jb; ++g) {
        const long long rb = (long long)g * nr;
        double rv[EPT], dv[EPT], cr[EPT], zv[EPT];
        [[maybe_unused]] double bs[EPT], r0[EPT];
        load_row<EPT, VEC>(m.rhs + rb, i0, nr, rv);
        load_row<EPT, VEC>(L.dinv + rb, i0, nr, dv);
        load_row<EPT, VEC>(L.cR + rb, i0, nr, cr);
        if (m.base) load_row<EPT, VEC>(m.base + rb, i0, nr, bs);
        if (m.final_) load_row<EPT, VEC>(d.r + rb, i0, nr, r0);
        double m0_edge = 0.0;
        if (lane == 0 && tl != 0 && i0 < nr) m0_edge = L.cR[rb + i0 - 1] * L.dinv[rb + i0 - 1];
        line_solve<TPL, EPT>(rv, dv, cr, m0_edge, lane, wt, team_in_cta, s_fA, s_fB, s_bA, s_bB, zv);
#pragma unroll
        for (int k = 0; k < EPT; ++k) zv[k] = m.base ? fma(omega, zv[k], bs[k]) : omega * zv[k];
        store_row<EPT, VEC>(m.out + rb, i0, nr, zv);
        if (m.final_) {
#pragma unroll
            for (int k = 0; k < EPT; ++k) acc = fma(r0[k], zv[k], acc);
        }
        // boundary rows -> the neighbours' ghost rows (final stage: ordered before the flag of the r.z reduction)
        if (g == 1 && m.xc.push_lo) stage_push<TPL, EPT, VEC>(d, m.xc, true, m.xc.push_lo, i0, nr, zv, tl, team_in_cta, cyc);
        if (g == nz && m.xc.push_hi) stage_push<TPL, EPT, VEC>(d, m.xc, false, m.xc.push_hi, i0, nr, zv, tl, team_in_cta, cyc);
    }
    return acc;
}

template <int TPL, int EPT, bool VEC>
__global__ void __launch_bounds__(TeamGeom<TPL>::CTA, TeamGeom<TPL>::minb_B(EPT))
k_mg_line(HpDev d, HpKArgs a, HpLevel L, HpMgLineArgs m) {
    using G = TeamGeom<TPL>;
    if (d.st->mg_skip) return;
    __shared__ double s_fA[G::LPC][G::NW], s_fB[G::LPC][G::NW], s_bA[G::LPC][G::NW], s_bB[G::LPC][G::NW];
    const int tic = threadIdx.x / TPL;
    double acc[1];
    acc[0] = mg_line_body<TPL, EPT, VEC>(d, a.omega, L, m, (long long)blockIdx.x * G::LPC + tic, (long long)gridDim.x * G::LPC,
                                         s_fA[tic], s_fB[tic], s_bA[tic], s_bB[tic]);
    if (m.final_) {
        double tot[1];
        if (cta_publish<1>(acc, 0u, d.partials, &d.st->ticket, tot, d.p2p != 0) && threadIdx.x == 0) {
            if (d.nranks > 1) stash_local<1>(d, tot, HP_COMM_MG);
            else finalize_B(d, a, a.init_phase != 0, tot[0], 0.0, 0.0, 2);
        }
    }
}

template <int TPL, int EPT, bool VEC>
const void* fn_mg_res() { return (const void*)k_mg_res<TPL, EPT, VEC>; }
template <int TPL, int EPT, bool VEC>
const void* fn_mg_line() { return (const void*)k_mg_line<TPL, EPT, VEC>; }

// ------------------------------------------------------------------------------------------------
// k_factor_team : pivots of the radial-line tridiagonal blocks  M_j = tridiag(-cR[i-1], diag[i], -cR[i])
//   w_0 = diag_0 ;  w_i = diag_i - cR_{i-1}^2 / w_{i-1} ;  dinv_i = 1 / w_i
// of the rows of levels first..n-1 of a level set (the fine level alone, or all coarse levels in one launch).
// The recurrence is a Moebius map, i.e. linear on the pair (p, q) with w = p / q:
//   (p_i, q_i)^T = [[diag_i, -cR_{i-1}^2], [1, 0]] (p_{i-1}, q_{i-1})^T ,  (p_-1, q_-1) = (1, 0)
// so the prefix products of those 2x2 matrices are an associative scan: a row team (the shape of the line solve: EPT
// cells per thread) composes its cells, the warp scans the products with shuffles, the warps of the team exchange their
// totals through shared memory, and every thread restarts the true recurrence from w at the cell before its first one.
// The partial products are continuants and grow geometrically: every combination is rescaled by a power of two (exact).
// The scanned start value carries a relative error of ~1e-10 (nr = 1024) .. 3e-9 (nr = 4096) against the sequential
// recurrence (tools/proto_pivot_scan.py); the pivots only define the preconditioner M = L D L^T (symmetric positive
// definite for any positive w), never the answer.  Replaces a thread-per-row chain of nr dependent divisions
// (0.13 ms at 4096 x 1024, 0.52 ms at nr = 4096 whatever the row count).
// ------------------------------------------------------------------------------------------------
struct M2 { double a, b, c, d; };       // [[a, b], [c, d]]
__device__ __forceinline__ void m2_rescale(M2& m) {
    const double s = fmax(fmax(fabs(m.a), fabs(m.b)), fmax(fabs(m.c), fabs(m.d)));
    const int ex = (__double2hiint(s) >> 20) & 0x7ff;                       // biased exponent of the largest entry
    const double sc = __hiloint2double((2046 - ex) << 20, 0);               // 2^-(exponent): exact scaling
    m.a *= sc; m.b *= sc; m.c *= sc; m.d *= sc;
}
// the map `later` applied after the map `earlier`
__device__ __forceinline__ M2 m2_after(const M2& later, const M2& earlier) {
    M2 r;
    r.a = fma(later.a, earlier.a, later.b * earlier.c);
    r.b = fma(later.a, earlier.b, later.b * earlier.d);
    r.c = fma(later.c, earlier.a, later.d * earlier.c);
    r.d = fma(later.c, earlier.b, later.d * earlier.d);
    return r;
}

template <int TPL, int EPT, bool VEC>
__global__ void __launch_bounds__(TeamGeom<TPL>::CTA, TeamGeom<TPL>::minb_B(EPT)) k_factor_team(HpLevelSet ls, int first, int nr,
                                                                                            const int* skip) {
    using G = TeamGeom<TPL>;
    if (*skip) return;
    constexpr int NW = G::NW, W = G::W;
    __shared__ double s_tot[G::LPC][NW][4];
    const int team_in_cta = threadIdx.x / TPL, tl = threadIdx.x % TPL, lane = threadIdx.x & (W - 1);
    const int wt = tl >> 5;
    const int i0 = tl * EPT;
    const unsigned tm = team_mask<TPL>();
    long long total = 0;
    for (int l = first; l < ls.n; ++l) total += ls.lev[l].nz;
    const long long team = (long long)blockIdx.x * G::LPC + team_in_cta, nteams = (long long)gridDim.x * G::LPC;
    const long long Ra = team * total / nteams, Rb = (team + 1) * total / nteams;
    int l = first;
    long long base = 0;                          // rows of the levels before level l
    for (long long R = Ra; R < Rb; ++R) {
        while (R - base >= ls.lev[l].nz) { base += ls.lev[l].nz; ++l; }
        const long long rb = (R - base + 1) * nr;
        const double* __restrict__ dgr = ls.lev[l].diag + rb;
        const double* __restrict__ crr = ls.lev[l].cR + rb;
        double dv[EPT], cv[EPT];
        load_row<EPT, VEC>(dgr, i0, nr, dv);
        load_row<EPT, VEC>(crr, i0, nr, cv);
        double c_edge = 0.0;                     // coupling to the cell before this warp's first one
        if (lane == 0 && tl != 0 && i0 < nr) c_edge = crr[i0 - 1];
        double cprev = __shfl_up_sync(tm, cv[EPT - 1], 1, W);
        if (lane == 0) cprev = c_edge;
        // this thread's cells composed: T = M_{i0+EPT-1} ... M_{i0}
        M2 T{dv[0], -cprev * cprev, 1.0, 0.0};
#pragma unroll
        for (int e = 1; e < EPT; ++e) {
            const double c2 = cv[e - 1] * cv[e - 1];
            M2 n{fma(dv[e], T.a, -c2 * T.c), fma(dv[e], T.b, -c2 * T.d), T.a, T.b};
            T = n;
        }
        m2_rescale(T);
        // inclusive scan over the lanes of the warp (sub-warp teams: over the team's lanes)
        M2 S = T;
#pragma unroll
        for (int o = 1; o < W; o <<= 1) {
            M2 P;
            P.a = __shfl_up_sync(tm, S.a, o, W); P.b = __shfl_up_sync(tm, S.b, o, W);
            P.c = __shfl_up_sync(tm, S.c, o, W); P.d = __shfl_up_sync(tm, S.d, o, W);
            if (lane >= o) { S = m2_after(S, P); m2_rescale(S); }
        }
        // exclusive prefix inside the warp
        M2 E;
        E.a = __shfl_up_sync(tm, S.a, 1, W); E.b = __shfl_up_sync(tm, S.b, 1, W);
        E.c = __shfl_up_sync(tm, S.c, 1, W); E.d = __shfl_up_sync(tm, S.d, 1, W);
        if (lane == 0) { E.a = 1.0; E.b = 0.0; E.c = 0.0; E.d = 1.0; }
        // (p, q) at the cell before this thread's first one = (E * carry) (1, 0)^T, carry = product of the warps before
        double p = E.a, q = E.c;
        if constexpr (NW > 1) {
            if (lane == W - 1) {
                double* t = s_tot[team_in_cta][wt];
                t[0] = S.a; t[1] = S.b; t[2] = S.c; t[3] = S.d;
            }
            team_barrier<TPL>(team_in_cta);
            double ca = 1.0, cc = 0.0;           // first column of the carry (the second one is never needed)
            for (int w = 0; w < wt; ++w) {
                const double* t = s_tot[team_in_cta][w];
                const double na = fma(t[0], ca, t[1] * cc), nc = fma(t[2], ca, t[3] * cc);
                const double s = fmax(fabs(na), fabs(nc));
                const int ex = (__double2hiint(s) >> 20) & 0x7ff;
                const double sc = __hiloint2double((2046 - ex) << 20, 0);
                ca = na * sc; cc = nc * sc;
            }
            p = fma(E.a, ca, E.b * cc);
            q = fma(E.c, ca, E.d * cc);
            team_barrier<TPL>(team_in_cta);      // s_tot is rewritten by the next row
        }
        // true recurrence from w_{i0-1} = p / q
        double prev = (tl == 0) ? 0.0 : cprev * cprev * (q / p);
        double out[EPT];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const double inv = 1.0 / (dv[e] - prev);
            out[e] = inv;
            prev = cv[e] * cv[e] * inv;
        }
        store_row<EPT, VEC>(ls.lev[l].dinv + rb, i0, nr, out);
    }
}
template <int TPL, int EPT, bool VEC>
const void* fn_factor_team() { return (const void*)k_factor_team<TPL, EPT, VEC>; }

// ---- instantiation table -----------------------------------------------------------------------
template <int TPL, int EPT, bool VEC>
const void* fn_A() { return (const void*)k_cg_A<TPL, EPT, VEC>; }
template <int TPL, int EPT, bool VEC>
const void* fn_B(int rline, int init) {
    if (rline) return init ? (const void*)k_cg_B<TPL, EPT, VEC, true, true> : (const void*)k_cg_B<TPL, EPT, VEC, true, false>;
    return init ? (const void*)k_cg_B<TPL, EPT, VEC, false, true> : (const void*)k_cg_B<TPL, EPT, VEC, false, false>;
}

#define HP_DISPATCH(cfg, EXPR_T)                                                          \
    do {                                                                                  \
        if (cfg.tpl == 16 && cfg.ept == 1) { EXPR_T(16, 1, false); }                      \
        else if (cfg.tpl == 32 && cfg.ept == 1) { EXPR_T(32, 1, false); }                 \
        else if (cfg.tpl == 32 && cfg.ept == 4) { if (cfg.vec) { EXPR_T(32, 4, true); } else { EXPR_T(32, 4, false); } }       \
        else if (cfg.tpl == 128 && cfg.ept == 4) { if (cfg.vec) { EXPR_T(128, 4, true); } else { EXPR_T(128, 4, false); } }    \
        else if (cfg.tpl == 256 && cfg.ept == 4) { if (cfg.vec) { EXPR_T(256, 4, true); } else { EXPR_T(256, 4, false); } }    \
        else if (cfg.tpl == 128 && cfg.ept == 8) { if (cfg.vec) { EXPR_T(128, 8, true); } else { EXPR_T(128, 8, false); } }    \
        else if (cfg.tpl == 256 && cfg.ept == 8) { if (cfg.vec) { EXPR_T(256, 8, true); } else { EXPR_T(256, 8, false); } }    \
        else if (cfg.tpl == 512 && cfg.ept == 8) { if (cfg.vec) { EXPR_T(512, 8, true); } else { EXPR_T(512, 8, false); } }    \
        else if (cfg.tpl == 1024 && cfg.ept == 8) { if (cfg.vec) { EXPR_T(1024, 8, true); } else { EXPR_T(1024, 8, false); } } \
    } while (0)

#define HP_DISPATCH_A(cfg, EXPR_T)                                                        \
    do {                                                                                  \
        if (cfg.tplA == 16 && cfg.eptA == 1) { EXPR_T(16, 1, false); }                    \
        else if (cfg.tplA == 32 && cfg.eptA == 1) { EXPR_T(32, 1, false); }               \
        else if (cfg.tplA == 32 && cfg.eptA == 4) { if (cfg.vecA) { EXPR_T(32, 4, true); } else { EXPR_T(32, 4, false); } }    \
        else if (cfg.tplA == 128 && cfg.eptA == 4) { if (cfg.vecA) { EXPR_T(128, 4, true); } else { EXPR_T(128, 4, false); } } \
        else if (cfg.tplA == 256 && cfg.eptA == 4) { if (cfg.vecA) { EXPR_T(256, 4, true); } else { EXPR_T(256, 4, false); } } \
    } while (0)

}  // namespace

// ================================================================================================
// host side: configuration + launch descriptors
// ================================================================================================
namespace hpk {

bool pick_config(int nr, int nz, int num_sms, HpLaunchCfg* cfg, const char** err) {
    if (nr < 1 || nz < 1) { *err = "nr and nz must be >= 1"; return false; }
    int tpl, ept;
    if (nr <= 16) { tpl = 16; ept = 1; }
    else if (nr <= 32) { tpl = 32; ept = 1; }
    else if (nr <= 128) { tpl = 32; ept = 4; }
    else if (nr <= 512) { tpl = 128; ept = 4; }
    else if (nr <= 1024) { tpl = 256; ept = 4; }
    else if (nr <= 4096) { tpl = 512; ept = 8; }
    else if (nr <= 8192) { tpl = 1024; ept = 8; }
    else { *err = "nr > 8192 radial cells per row is not supported"; return false; }
    cfg->tpl = tpl; cfg->ept = ept;
    cfg->vec = (ept >= 4) && (nr % ept == 0);
    cfg->cta_threads = tpl < 256 ? 256 : tpl;
    cfg->teams_per_cta = cfg->cta_threads / tpl;
    // k_cg_A: at most 256 threads x 4 cells per tile (three rows of the direction + five streams stay in <= 128 regs)
    if (nr <= 16) { cfg->tplA = 16; cfg->eptA = 1; }
    else if (nr <= 32) { cfg->tplA = 32; cfg->eptA = 1; }
    else if (nr <= 128) { cfg->tplA = 32; cfg->eptA = 4; }
    else if (nr <= 512) { cfg->tplA = 128; cfg->eptA = 4; }
    else { cfg->tplA = 256; cfg->eptA = 4; }
    cfg->vecA = (cfg->eptA >= 4) && (nr % cfg->eptA == 0);
    cfg->ctaA = 256;
    cfg->teams_per_ctaA = 256 / cfg->tplA;
    cfg->ncolA = (nr + cfg->tplA * cfg->eptA - 1) / (cfg->tplA * cfg->eptA);
    cfg->clF = nr <= 1024 ? 1 : (nr <= 2048 ? 2 : (nr <= 4096 ? 4 : 0));     // CTAs per row of the fused V-cycle stages
    cfg->num_sms = num_sms; cfg->nz = nz;
    cfg->cta_cell = 256;
    long long ncell = (long long)(nz + 1) * nr;
    long long gc = (ncell + 255) / 256;
    long long cap = (long long)num_sms * 8;
    cfg->grid_cell = (int)(gc < cap ? gc : cap);
    if (cfg->grid_cell > HP_MAX_PARTIALS) cfg->grid_cell = HP_MAX_PARTIALS;
    return true;
}

// Grid of a marching kernel: exactly one resident wave (occupancy of THIS instantiation x SM count) so that no
// partially filled second wave appears; fewer CTAs when the mesh has too few rows (>= 2 rows per team).
static int team_grid(const void* func, const HpLaunchCfg& c, int cta_threads, int teams_per_cta, int ncol, int min_units = 2) {
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, func, cta_threads, 0) != cudaSuccess || occ < 1) occ = 1;
    long long max_teams = (long long)c.num_sms * occ * teams_per_cta;
    long long chunks = max_teams / ncol;                 // row chunks; each is served by ncol column-tile teams
    if (chunks < 1) chunks = 1;
    if (chunks > c.nz) chunks = c.nz;
    if (chunks > 1 && c.nz / chunks < min_units) chunks = (c.nz + min_units - 1) / min_units;
    long long teams = chunks * ncol;
    long long grid = (teams + teams_per_cta - 1) / teams_per_cta;
    if (grid < 1) grid = 1;
    if (grid > HP_MAX_PARTIALS) grid = HP_MAX_PARTIALS;
    return (int)grid;
}

HpKernelLaunch desc_snapshot_kout(const HpLaunchCfg& c, const HpDev& d) {
    return {(const void*)k_snapshot_kout, dim3((d.nz + 2 + 127) / 128), dim3(128), 0};
}
HpKernelLaunch desc_coef(const HpLaunchCfg& c, const HpDev& d) {
    return {(const void*)k_coef, dim3(c.grid_cell), dim3(c.cta_cell), 0};
}
HpKernelLaunch desc_stash(const HpLaunchCfg& c, const HpDev& d) {
    return {(const void*)k_stash, dim3(c.grid_cell), dim3(c.cta_cell), 0};
}
HpKernelLaunch desc_delta(const HpLaunchCfg& c, const HpDev& d) {
    return {(const void*)k_delta, dim3(c.grid_cell), dim3(c.cta_cell), 0};
}
int diag_tpr_log2(int nr) {
    int l = 0;
    while ((1 << l) < nr && l < 8) ++l;
    return l;                                   // threads per row = min(256, next power of two >= nr)
}
HpKernelLaunch desc_diag(const HpLaunchCfg& c, const HpDev& d) {
    const int rpb = 256 >> diag_tpr_log2(d.nr);
    int occ = 1;
    const void* f = d.nr <= 16 ? (const void*)k_diag<true> : (const void*)k_diag<false>;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, f, 256, 0) != cudaSuccess || occ < 1) occ = 1;
    long long g = ((long long)d.nz + rpb - 1) / rpb, cap = (long long)c.num_sms * occ;
    if (g > cap) g = cap;
    if (g > HP_MAX_PARTIALS) g = HP_MAX_PARTIALS;
    return {f, dim3((unsigned)(g < 1 ? 1 : g)), dim3(256), 0};
}
// pivots of `rows` rows in total (the fine level, or every coarse level in one launch): args (HpLevelSet, int first, int nr, const int* skip)
HpKernelLaunch desc_factor_rows(const HpLaunchCfg& c, long long rows) {
    const void* f = nullptr;
#define HP_X(T, E, V) f = fn_factor_team<T, E, V>()
    HP_DISPATCH(c, HP_X);
#undef HP_X
    HpLaunchCfg cl = c;
    cl.nz = (int)(rows < 1 ? 1 : rows);
    return {f, dim3(team_grid(f, cl, c.cta_threads, c.teams_per_cta, 1)), dim3(c.cta_threads), 0};
}
HpKernelLaunch desc_mg_coarsen(const HpLaunchCfg& c, int nzc, int nr) {
    long long g = ((long long)(nzc + 1) * nr + 255) / 256;
    long long cap = (long long)c.num_sms * 8;
    return {(const void*)k_mg_coarsen, dim3((unsigned)(g < cap ? (g < 1 ? 1 : g) : cap)), dim3(256), 0};
}
HpKernelLaunch desc_mg_res(const HpLaunchCfg& c, int nz_level) {
    const void* f = nullptr;
#define HP_X(T, E, V) f = fn_mg_res<T, E, V>()
    HP_DISPATCH_A(c, HP_X);
#undef HP_X
    HpLaunchCfg cl = c;
    cl.nz = (nz_level + 1) / 2;                      // chunks are whole row pairs; small levels: one pair per team (they are
                                                     // chains of dependent row latencies, not bandwidth)
    return {f, dim3(team_grid(f, cl, c.ctaA, c.teams_per_ctaA, c.ncolA, 1)), dim3(c.ctaA), 0};
}
// fused stage on cluster teams: one resident wave of clusters, >= one row pair per cluster
static HpKernelLaunch desc_fused_cluster(const HpLaunchCfg& c, int nz_level, int mode) {
    const void* f = fn_mg_fused_cluster(c.clF, c.vecA, mode);
    int nclus = 0;
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3(c.clF * c.num_sms); lc.blockDim = dim3(256);
    if (cudaOccupancyMaxActiveClusters(&nclus, f, &lc) != cudaSuccess || nclus < 1) {
        cudaGetLastError();
        nclus = (c.num_sms * 2 / c.clF) * 7 / 8;
    }
    long long pairs = (nz_level + 1) / 2;
    long long chunks = nclus < pairs ? nclus : pairs;
    if (chunks < 1) chunks = 1;
    long long grid = chunks * c.clF;
    if (grid > HP_MAX_PARTIALS) grid = (HP_MAX_PARTIALS / c.clF) * c.clF;
    return {f, dim3((unsigned)grid), dim3(256), 0};
}
HpKernelLaunch desc_mg_up(const HpLaunchCfg& c, int nz_level) {       // only when mg_fused_supported(c)
    if (c.clF > 1) return desc_fused_cluster(c, nz_level, 1);
    const void* f = nullptr;
#define HP_X(T, E, V) f = fn_mg_up<T, E, V>()
    HP_DISPATCH_A(c, HP_X);
#undef HP_X
    HpLaunchCfg cl = c;
    cl.nz = (nz_level + 1) / 2;
    return {f, dim3(team_grid(f, cl, c.ctaA, c.teams_per_ctaA, 1, 1)), dim3(c.ctaA), 0};
}
HpKernelLaunch desc_mg_final(const HpLaunchCfg& c, int nz_level) {    // only when mg_fused_supported(c)
    if (c.clF > 1) return desc_fused_cluster(c, nz_level, 3);
    const void* f = nullptr;
#define HP_X(T, E, V) f = fn_mg_final<T, E, V>()
    HP_DISPATCH_A(c, HP_X);
#undef HP_X
    HpLaunchCfg cl = c;
    cl.nz = (nz_level + 1) / 2;
    return {f, dim3(team_grid(f, cl, c.ctaA, c.teams_per_ctaA, 1, 1)), dim3(c.ctaA), 0};
}
HpKernelLaunch desc_mg_down(const HpLaunchCfg& c, int nz_level) {     // only when mg_fused_supported(c)
    if (c.clF > 1) return desc_fused_cluster(c, nz_level, 2);
    const void* f = nullptr;
#define HP_X(T, E, V) f = fn_mg_down<T, E, V>()
    HP_DISPATCH_A(c, HP_X);
#undef HP_X
    HpLaunchCfg cl = c;
    cl.nz = (nz_level + 1) / 2;
    return {f, dim3(team_grid(f, cl, c.ctaA, c.teams_per_ctaA, 1, 1)), dim3(c.ctaA), 0};
}
HpKernelLaunch desc_mg_line(const HpLaunchCfg& c, int nz_level) {
    const void* f = nullptr;
#define HP_X(T, E, V) f = fn_mg_line<T, E, V>()
    HP_DISPATCH(c, HP_X);
#undef HP_X
    HpLaunchCfg cl = c;
    cl.nz = nz_level;
    return {f, dim3(team_grid(f, cl, c.cta_threads, c.teams_per_cta, 1)), dim3(c.cta_threads), 0};
}
// the fused stages need the row teams of the residual and of the line solve to have the same shape (nr <= 1024)
bool mg_fused_supported(const HpLaunchCfg& c) {
    if (c.clF > 1) return true;                       // cluster teams (1024 < nr <= 4096)
    return c.clF == 1 && c.ncolA == 1 && c.tplA == c.tpl && c.eptA == c.ept && c.vecA == c.vec && c.ctaA == c.cta_threads;
}
HpKernelLaunch desc_finalize() { return {(const void*)k_finalize, dim3(1), dim3(32), 0}; }
HpKernelLaunch desc_wait() { return {(const void*)k_wait, dim3(1), dim3(32), 0}; }
HpKernelLaunch desc_cg_A(const HpLaunchCfg& c, const HpDev& d) {
    const void* f = nullptr;
#define HP_X(T, E, V) f = fn_A<T, E, V>()
    HP_DISPATCH_A(c, HP_X);
#undef HP_X
    return {f, dim3(team_grid(f, c, c.ctaA, c.teams_per_ctaA, c.ncolA)), dim3(c.ctaA), 0};
}
static HpKernelLaunch desc_B(const HpLaunchCfg& c, int precond, int init) {
    const void* f = nullptr;
#define HP_X(T, E, V) f = fn_B<T, E, V>(precond, init)
    HP_DISPATCH(c, HP_X);
#undef HP_X
    return {f, dim3(team_grid(f, c, c.cta_threads, c.teams_per_cta, 1)), dim3(c.cta_threads), 0};
}
HpKernelLaunch desc_cg_init(const HpLaunchCfg& c, const HpDev& d, int precond, int criterion) { return desc_B(c, precond, 1); }
// the TMA-staged variant serves the 256-thread x 4-cell row shape (512 < nr <= 1024, nr % 4 == 0) with the line
// preconditioner and a stopping rule that does not need the diagonal
bool cg_B_tma_supported(const HpLaunchCfg& c, int precond, int criterion) {
    return precond != 0 && criterion != 1 && c.tpl == 256 && c.ept == 4 && c.vec;
}
HpKernelLaunch desc_cg_B(const HpLaunchCfg& c, const HpDev& d, int precond, int criterion, int tma) {
    if (!(tma && cg_B_tma_supported(c, precond, criterion))) return desc_B(c, precond, 0);
    const void* f = (const void*)k_cg_B_tma;
    const int rowp = (d.nr + 3) & ~3;
    const size_t smem = (size_t)BT_STAGES * BT_ARR * rowp * sizeof(double) + BT_STAGES * sizeof(unsigned long long);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 6 * 1024 * 8 + 64);
        attr_set = true;
    }
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, f, 256, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
    long long chunks = (long long)c.num_sms * occ;
    if (chunks > c.nz) chunks = c.nz;
    if (chunks > 1 && c.nz / chunks < 2) chunks = (c.nz + 1) / 2;
    if (chunks > HP_MAX_PARTIALS) chunks = HP_MAX_PARTIALS;
    return {f, dim3((unsigned)(chunks < 1 ? 1 : chunks)), dim3(256), smem};
}

// one pipe per CTA (nr == 9 rows of <= 512 lines); returns false when this row shape has no instantiation
bool pipe_kernel_supported(int nr, int nzm) { return nr == 9 && nzm >= 1 && nzm <= 512; }
HpKernelLaunch desc_pipe_step(int nr, int nzm, int members) {
    const void* f = (const void*)k_pipe_step<9>;
    const size_t smem = (size_t)6 * nr * nzm * sizeof(double);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 9 * 512 * 8);
        attr_set = true;
    }
    return {f, dim3((unsigned)members), dim3((unsigned)((nzm + 31) / 32 * 32)), smem};
}
HpKernelLaunch desc_pipe_finalize() { return {(const void*)k_pipe_finalize, dim3(1), dim3(256), 0}; }

cudaError_t launch_eval_property(const hp_phys& ph, int id, const double* T, double* out, long long n, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    long long g = (n + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    k_eval_property<<<(int)g, 256, 0, s>>>(ph, id, T, out, n);
    return cudaGetLastError();
}
cudaError_t launch_cell_fields(const HpDev& d, const hp_phys& ph, double* rho, double* cp, double* kavg, cudaStream_t s) {
    long long n = (long long)d.nz * d.nr;
    long long g = (n + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    k_cell_fields<<<(int)g, 256, 0, s>>>(d, ph, rho, cp, kavg);
    return cudaGetLastError();
}
cudaError_t launch_copy_owned(const double* src_ghost, double* dst, int nr, long long n, cudaStream_t s) {
    long long g = (n + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    k_copy_owned<<<(int)g, 256, 0, s>>>(src_ghost, dst, nr, n);
    return cudaGetLastError();
}
cudaError_t launch_fill_rows(const HpDev& d, double* field, cudaStream_t s) {
    long long n = (long long)(d.nz + 2) * d.nr;
    long long g = (n + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    k_fill_rows<<<(int)g, 256, 0, s>>>(d, field);
    return cudaGetLastError();
}

}  // namespace hpk
