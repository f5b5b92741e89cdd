// C ABI of libhpsolver: handle life cycle, sweep / step orchestration (plain stream launches or a CUDA
// graph with a device-side WHILE around the PCG iteration), statistics, optional NCCL axial slabs.
// See include/hp_solver.h for the contract and the reference lines each entry point replaces.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "hp_internal.h"


static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return -1; }
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(std::string(#call) + " -> " + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                        std::to_string(__LINE__) + ")");                                           \
    } while (0)

// ---- minimal NCCL binding, resolved at run time so the library loads without NCCL ------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclFloat64 = 8, ncclSum = 0, ncclMax = 2 };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static bool load_nccl() {
    if (g_nccl.lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return false;
#define LD(sym) *(void**)(&g_nccl.sym) = dlsym(lib, "nccl" #sym)
    LD(GetUniqueId); LD(CommInitRank); LD(CommDestroy); LD(AllReduce); LD(AllGather); LD(Send); LD(Recv);
    LD(GroupStart); LD(GroupEnd); LD(GetErrorString);
#undef LD
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.Send || !g_nccl.Recv) return false;
    g_nccl.lib = lib;
    return true;
}
#define NK(call)                                                                                        \
    do {                                                                                                \
        int r_ = (call);                                                                                \
        if (r_ != 0)                                                                                    \
            return fail(std::string(#call) + " -> " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error")); \
    } while (0)

// Coarse levels inside the level block: level l has ceil(nz_{l-1} / 2) rows, 8 fields each (cR, cZ, diag, dinv, rhs, x, x2,
// tmp), ghost-row layout like the fine fields.
enum { LV_CR = 0, LV_CZ, LV_DIAG, LV_DINV, LV_RHS, LV_X, LV_X2, LV_TMP, LV_FIELDS };
struct LevelLayout { int n; int nz[HP_MAX_LEVELS]; size_t off[HP_MAX_LEVELS][LV_FIELDS]; size_t elems; };
static LevelLayout level_layout(int nz, int nr) {
    LevelLayout L{};
    L.n = 1; L.nz[0] = nz;
    size_t e = 0;
    for (int l = 1; l < HP_MAX_LEVELS && L.nz[l - 1] > 1; ++l) {
        const int rows = (L.nz[l - 1] + 1) / 2;
        L.nz[l] = rows;
        for (int k = 0; k < LV_FIELDS; ++k) { L.off[l][k] = e; e += (size_t)(rows + 2) * nr + 64; }
        L.n = l + 1;
    }
    L.elems = e;
    return L;
}

struct GraphEntry {
    double dt; int sweeps; int has_old; int update_old; hp_solve_opts opts;
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    int static_nodes = 0;
    int body_nodes = 2;
};

struct hp_solver_s {
    int device = 0, num_sms = 148;
    cudaStream_t stream = nullptr;
    hp_phys phys{};
    hp_solve_opts opts{};
    int req_precond = 1;             // what the caller asked for (opts.precond falls back to 1 on a mesh too short to coarsen;
                                     // re-resolved when the slab topology changes)
    HpDev d{};
    HpLaunchCfg cfg{};
    std::vector<void*> allocs;
    size_t field_elems = 0;
    double* rhs_dbg = nullptr;
    // host mirrors
    HpState* h_state = nullptr;     // pinned
    long long launches_host = 0;    // kernels launched by the host + static graph nodes per graph launch
    long long stream_mode_iters = 0;
    long long sweeps_host = 0;
    int graph_body_nodes = 2;
    std::vector<GraphEntry> graphs;
    // slabs
    int rank = 0, nranks = 1;
    ncclComm_t nccl = nullptr;
    // 'mgline' hierarchy (level 0 aliases the fine fields; the coarse levels live in ONE allocation whose layout is a
    // function of (nz, nr) alone, so a rank can address the level arrays of its neighbouring slabs through peer memory)
    HpLevel lev[HP_MAX_LEVELS];
    int nlev = 0;
    double* lvblock = nullptr;
    LevelLayout lay{}, lay_lo{}, lay_hi{};
    double *lv_lo = nullptr, *lv_hi = nullptr;      // level blocks of the lower / upper neighbour (peer mappings)
    std::vector<int> nz_of_rank;
    int* h_fault = nullptr;          // pinned + mapped: raised by a kernel whose wait for a peer timed out
    // persistent one-pipe-per-CTA step (k_pipe_step): the mesh is pipe_M uncoupled members of pipe_nzm rows each
    int pipe_nzm = 0, pipe_M = 0;
    double* pipe_part = nullptr;     // [sweeps][M][4] member partials of a step
    int pipe_part_sweeps = 0;
    int pipe_hist = 0;               // completed steps of history (Told = start field of the last step, d1 = increment before)
    bool pipe_c2_valid = false;      // d.c2 holds the re-sweep correction of the last completed step
    double pipe_hist_dt = 0.0;
    // peer-memory exchange (hp_peer_export / hp_peer_import)
    HpComm* comm = nullptr;          // my mailbox
    double* qz_block = nullptr;      // base of the (q, z) allocation (IPC handle is per allocation)
    std::vector<void*> ipc_opened;
    bool mg_upfused = true;      // fused V-cycle stages (k_mg_fused<1|2>) usable: one row team owns a whole row
    int hist = 0;                    // completed steps since the field was last set: Told (and d1, d2) hold that much history
    double hist_dt = 0.0;            // time step the history was taken with
    bool c2_stashed = false;         // d.c2 holds the first-sweep solution of the step just finished
    bool ghost_dirty = true;         // ghost rows of T must be refreshed by an explicit exchange before the next sweep
    // L2 access-policy window over the contiguous (q, z) block
    bool l2_on = false;
    bool l2_limit_changed = false;      // cudaLimitPersistingL2CacheSize is device-wide: put back at hp_destroy
    size_t l2_limit_before = 0;
    void* l2_base = nullptr;
    size_t l2_bytes = 0;
    cudaAccessPolicyWindow l2_window{};
    // per-kernel event timing (hp_profile_begin / hp_profile_end; stream-launch mode only)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<int> prof_kind;
    size_t prof_used = 0;
};

static int check_fault(hp_solver_s* h) {
    if (h->h_fault && *(volatile int*)h->h_fault)
        return fail("a wait for a neighbouring slab timed out (hp_solve_opts.peer_timeout_ms): the simulation on this handle is "
                    "stopped; synchronise all ranks and call hp_comm_clear_fault, or destroy the handle");
    return 0;
}

static int fetch_state(hp_solver_s* h) {
    CK(cudaMemcpyAsync(h->h_state, h->d.st, sizeof(HpState), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

static int dev_alloc(hp_solver_s* h, void** p, size_t bytes) {
    CK(cudaMalloc(p, bytes ? bytes : 8));
    h->allocs.push_back(*p);
    return 0;
}
template <typename T>
static int upload(hp_solver_s* h, const T** dst, const std::vector<T>& v) {
    void* p = nullptr;
    if (dev_alloc(h, &p, v.size() * sizeof(T))) return -1;
    CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (const T*)p;
    return 0;
}

// L2 residency of the producer->consumer vectors of the PCG iteration (q: k_cg_A -> k_cg_B, z: k_cg_B -> k_cg_A).
// The window is attached to every PCG kernel launch (stream attribute in stream mode, kernel-node attribute in
// graphs).
static void setup_l2_window(hp_solver_s* h, const cudaDeviceProp& prop) {
    h->l2_on = false;
    if (prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) return;
    // Only when the whole window fits the persisting carve-out: a partially persisting window (hitRatio < 1) over
    // vectors larger than L2 halves the streaming rate of both PCG kernels (measured on 16384x4096, profiles/).
    size_t win = h->l2_bytes;                                   // q and z
    const size_t cap = (size_t)prop.persistingL2CacheMaxSize < (size_t)prop.accessPolicyMaxWindowSize
                           ? (size_t)prop.persistingL2CacheMaxSize : (size_t)prop.accessPolicyMaxWindowSize;
    if (win > cap) return;
    if (cudaDeviceGetLimit(&h->l2_limit_before, cudaLimitPersistingL2CacheSize) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, win) != cudaSuccess) { cudaGetLastError(); return; }
    h->l2_limit_changed = true;
    memset(&h->l2_window, 0, sizeof(h->l2_window));
    h->l2_window.base_ptr = h->l2_base;
    h->l2_window.num_bytes = win;
    h->l2_window.hitRatio = 1.0f;
    h->l2_window.hitProp = cudaAccessPropertyPersisting;
    h->l2_window.missProp = cudaAccessPropertyStreaming;
    h->l2_on = true;
}

static int apply_l2_to_stream(hp_solver_s* h) {
    if (!h->l2_on) return 0;
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));
    v.accessPolicyWindow = h->l2_window;
    CK(cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &v));
    return 0;
}

// ---- 'mgline' hierarchy: level l+1 has ceil(nz_l / 2) rows; a level is added while the one above has >= 32 rows ------
static int local_level_count(long long nz, int max_levels) {
    int n = 1;
    for (long long nzl = nz; n < max_levels && nzl >= 32; nzl = (nzl + 1) / 2) ++n;
    return n;
}

// Number of levels.  One GPU, or slabs without peer memory (rank-local cycle): from the local row count.  Peer-memory
// slabs: the hierarchy of the WHOLE mesh, distributed -- the same count on every rank, from the global row count, reduced
// until the row pairs of every level line up with the slab cuts (every slab but the last a multiple of 2^(levels-1) rows).
static int level_count_for(const hp_solver_s* h) {
    const int nr_ = h->nranks;
    if (!(nr_ > 1 && h->d.p2p && (int)h->nz_of_rank.size() == nr_)) return local_level_count(h->d.nz, h->opts.mg_levels);
    long long total = 0;
    for (int v : h->nz_of_rank) total += v;
    int n = local_level_count(total, h->opts.mg_levels);
    for (; n >= 2; --n) {
        const int A = 1 << (n - 1);
        bool ok = true;
        for (int r = 0; r + 1 < nr_; ++r) ok = ok && (h->nz_of_rank[r] % A == 0);
        if (ok) break;
    }
    return n < 1 ? 1 : n;
}

static int ensure_levels(hp_solver_s* h) {
    if (h->lvblock) return 0;
    HpDev& d = h->d;
    h->lay = level_layout(d.nz, d.nr);
    if (dev_alloc(h, (void**)&h->lvblock, (h->lay.elems + 64) * sizeof(double))) return -1;
    CK(cudaMemsetAsync(h->lvblock, 0, (h->lay.elems + 64) * sizeof(double), h->stream));
    HpLevel& L0 = h->lev[0];
    L0.nz = d.nz; L0.cR = d.cR; L0.cZ = d.cZ; L0.diag = d.diag; L0.dinv = d.dinv;
    L0.rhs = d.r; L0.x = d.z; L0.tmp = d.q;
    if (dev_alloc(h, (void**)&L0.x2, h->field_elems * sizeof(double))) return -1;
    CK(cudaMemsetAsync(L0.x2, 0, h->field_elems * sizeof(double), h->stream));
    for (int l = 1; l < h->lay.n; ++l) {
        HpLevel& L = h->lev[l];
        L.nz = h->lay.nz[l];
        double** f[LV_FIELDS] = {&L.cR, &L.cZ, &L.diag, &L.dinv, &L.rhs, &L.x, &L.x2, &L.tmp};
        for (int k = 0; k < LV_FIELDS; ++k) *f[k] = h->lvblock + h->lay.off[l][k];
    }
    h->mg_upfused = hpk::mg_fused_supported(h->cfg);
    return 0;
}

static int setup_mg(hp_solver_s* h) {
    if (ensure_levels(h)) return -1;
    int want = level_count_for(h);
    if (want > h->lay.n) want = h->lay.n;
    h->nlev = want;
    h->d.mg_local = (h->nranks > 1 && !h->d.p2p) ? 1 : 0;
    return 0;
}

// 'mgline' asked for: build the hierarchy for the current topology; a mesh too short to coarsen gets the plain line
// preconditioner.  Under peer-memory slabs the level count comes from the global mesh, so every rank takes the same
// decision; slabs without peer memory decide per rank -- the caller (distributed.make_solver_for_rank) resolves the
// preconditioner globally before it gets here.
static int resolve_mg(hp_solver_s* h) {
    if (h->req_precond != 2) return 0;
    h->opts.precond = 2;
    if (setup_mg(h)) return -1;
    if (h->nlev < 2) h->opts.precond = 1;
    return 0;
}

static HpKArgs make_args(const hp_solver_s* h, double dt, int has_old) {
    HpKArgs a{};
    a.dt = dt; a.tol = h->opts.tol; a.has_old = has_old; a.max_iter = h->opts.max_iter;
    a.criterion = h->opts.criterion; a.use_cond = 0; a.cond = 0; a.rhs_out = nullptr;
    a.mg = (h->opts.precond == 2) ? 1 : 0; a.init_phase = 0; a.omega = h->opts.mg_omega; a.shift = 0;
    // repeated converged sweeps as bookkeeping only (HpState.noop_ok, opt-in); under slabs the reducing kernels must run
    // on every rank for the flag protocol, so single GPU only
    a.noop_enable = (h->opts.reuse_converged && h->nranks == 1) ? 1 : 0;
    a.timeout_ns = (unsigned long long)h->opts.peer_timeout_ms * 1000000ull;
    a.ext_guard = h->opts.ext_guard_tols * h->opts.tol;
    return a;
}

enum { KIND_OTHER = 0, KIND_COEF, KIND_DIAG, KIND_FACTOR, KIND_INIT, KIND_A, KIND_B, KIND_MG_RES0, KIND_MG_LINE0, KIND_MG_COARSE,
       KIND_MG_SETUP, KIND_MG_L1, KIND_MG_L2, KIND_MG_L3, KIND_MG_L4, KIND_MG_L5, KIND_COUNT };   // L5 = level 5 and deeper

static int launch(hp_solver_s* h, const HpKernelLaunch& k, void** params, int kind = KIND_OTHER) {
    const bool prof = h->prof_on && h->prof_used < h->prof_ev.size() / 2;
    if (prof) CK(cudaEventRecord(h->prof_ev[2 * h->prof_used], h->stream));
    CK(cudaLaunchKernel(k.func, k.grid, k.block, params, k.smem, h->stream));
    if (prof) {
        CK(cudaEventRecord(h->prof_ev[2 * h->prof_used + 1], h->stream));
        h->prof_kind.push_back(kind);
        h->prof_used++;
    }
    h->launches_host++;
    return 0;
}

extern "C" {

const char* hp_last_error(void) { return g_err.c_str(); }
int hp_version(void) { return 100; }

int hp_create(hp_handle* out, const hp_mesh_desc* m, const hp_phys* phys, void* stream) {
    if (!out || !m || !phys) return fail("hp_create: null argument");
    if (!m->r_v || !m->z_v || !m->cell_band || !m->rface_band || !m->zface_band || !m->zc_flags || !m->zv_flags ||
        !m->axial_open || !m->q_line || !m->h_line || !m->Tamb_line || !m->eps_line)
        return fail("hp_create: null array in hp_mesh_desc");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("hp_create: no CUDA device available -- libhpsolver has no CPU fallback");
    hp_solver_s* h = new hp_solver_s();
    struct Guard { hp_solver_s* h; ~Guard() { if (h) hp_destroy(h); } } guard{h};    // every early return below frees the handle
    h->stream = (cudaStream_t)stream;
    h->phys = *phys;
    CK(cudaGetDevice(&h->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->device));
    h->num_sms = prop.multiProcessorCount;
    const char* err = nullptr;
    if (!hpk::pick_config(m->nr, m->nz, h->num_sms, &h->cfg, &err)) return fail(std::string("hp_create: ") + err);
    const int nr = m->nr, nz = m->nz;
    for (int i = 0; i < nr; ++i)
        if (!(m->r_v[i + 1] > m->r_v[i])) return fail("hp_create: r_v must be strictly increasing");
    for (int g = 0; g < nz + 2; ++g)
        if (!(m->z_v[g + 1] > m->z_v[g])) return fail("hp_create: z_v must be strictly increasing");

    HpDev& d = h->d;
    d.nr = nr; d.nz = nz; d.ncolA = h->cfg.ncolA; d.p2p = 0; d.seq = nullptr; d.z_lo = d.z_hi = d.T_lo = d.T_hi = nullptr;
    for (int r = 0; r < HP_MAX_RANKS; ++r) d.comm_peer[r] = nullptr;
    d.mg_local = 0; d.fault_host = nullptr;
    d.nranks = 1; d.rank = 0; d.red_local = nullptr; d.red_all = nullptr;
    std::vector<double> r_v(m->r_v, m->r_v + nr + 1), r_c(nr), dr(nr), ar(nr, 0.0), dcr(nr, 1.0);
    for (int i = 0; i < nr; ++i) { r_c[i] = 0.5 * (r_v[i] + r_v[i + 1]); dr[i] = r_v[i + 1] - r_v[i]; }
    for (int i = 0; i + 1 < nr; ++i) { dcr[i] = r_c[i + 1] - r_c[i]; ar[i] = (r_v[i + 1] - r_c[i]) / dcr[i]; }
    std::vector<double> z_c(nz + 2), dz(nz + 2), az(nz + 2, 0.0), dcz(nz + 2, 1.0);
    for (int g = 0; g < nz + 2; ++g) { z_c[g] = 0.5 * (m->z_v[g] + m->z_v[g + 1]); dz[g] = m->z_v[g + 1] - m->z_v[g]; }
    for (int g = 0; g + 1 < nz + 2; ++g) { dcz[g] = z_c[g + 1] - z_c[g]; az[g] = (m->z_v[g + 1] - z_c[g]) / dcz[g]; }
    std::vector<double> wc(nr, 1.0), wf(nr + 1, 1.0);
    if (!m->cartesian) { wc = r_c; wf = r_v; }
    d.A_out_r = wf[nr];
    d.d_out = r_v[nr] - r_c[nr - 1];
#define UP(field, vec) if (upload(h, &d.field, vec)) return -1;
    UP(r_v, r_v); UP(r_c, r_c); UP(dr, dr); UP(wc, wc); UP(wf, wf); UP(ar, ar); UP(dcr, dcr);
    UP(z_c, z_c); UP(dz, dz); UP(az, az); UP(dcz, dcz);
    UP(cell_band, std::vector<int8_t>(m->cell_band, m->cell_band + nr));
    UP(rface_band, std::vector<int8_t>(m->rface_band, m->rface_band + nr));
    UP(zface_band, std::vector<int8_t>(m->zface_band, m->zface_band + nr));
    UP(zc_flags, std::vector<uint8_t>(m->zc_flags, m->zc_flags + nz + 2));
    UP(zv_flags, std::vector<uint8_t>(m->zv_flags, m->zv_flags + nz + 2));
    UP(axial_open, std::vector<uint8_t>(m->axial_open, m->axial_open + nz + 2));
    UP(q_line, std::vector<double>(m->q_line, m->q_line + nz + 2));
    UP(h_line, std::vector<double>(m->h_line, m->h_line + nz + 2));
    UP(Tamb_line, std::vector<double>(m->Tamb_line, m->Tamb_line + nz + 2));
    UP(eps_line, std::vector<double>(m->eps_line, m->eps_line + nz + 2));
#undef UP
    {   // uncoupled members: maximal runs of rows between non-conducting faces (a single pipe is one member)
        int first = -1, len = -1, count = 0;
        bool uniform = m->axial_open[0] == 0 && m->axial_open[nz] == 0;
        for (int g = 0; g <= nz && uniform; ++g) {
            if (m->axial_open[g]) continue;
            if (first >= 0) {
                const int l = g - first;
                if (len < 0) len = l;
                uniform = uniform && (l == len);
                ++count;
            }
            first = g;
        }
        if (uniform && count >= 1 && (long long)count * len == nz && hpk::pipe_kernel_supported(nr, len)) { h->pipe_nzm = len; h->pipe_M = count; }
    }
    h->field_elems = (size_t)(nz + 2) * nr + 64;
    double** fields[] = {&d.T, &d.Told, &d.cR, &d.cZ, &d.diag, &d.dinv, &d.r, &d.p0, &d.p1};
    for (double** f : fields) {
        if (dev_alloc(h, (void**)f, h->field_elems * sizeof(double))) return -1;
        CK(cudaMemsetAsync(*f, 0, h->field_elems * sizeof(double), h->stream));
    }
    // q and z are produced by one PCG kernel and consumed by the next one: one contiguous block so that a single
    // L2 access-policy window can keep both resident across kernel boundaries (setup_l2_window)
    {
        double* qz = nullptr;
        if (dev_alloc(h, (void**)&qz, 2 * h->field_elems * sizeof(double))) return -1;
        CK(cudaMemsetAsync(qz, 0, 2 * h->field_elems * sizeof(double), h->stream));
        d.q = qz; d.z = qz + h->field_elems;
        h->qz_block = qz;
        h->l2_base = qz; h->l2_bytes = 2 * h->field_elems * sizeof(double);
    }
    setup_l2_window(h, prop);
    void* p = nullptr;
    if (dev_alloc(h, &p, sizeof(double) * (nz + 2))) return -1;
    d.k0_out = (const double*)p;
    if (dev_alloc(h, (void**)&d.st, sizeof(HpState))) return -1;
    CK(cudaMemsetAsync(d.st, 0, sizeof(HpState), h->stream));
    if (dev_alloc(h, (void**)&d.stats, sizeof(hp_sweep_stat) * HP_STATS_CAP)) return -1;
    CK(cudaMemsetAsync(d.stats, 0, sizeof(hp_sweep_stat) * HP_STATS_CAP, h->stream));
    if (dev_alloc(h, (void**)&d.partials, sizeof(double) * HP_NRED * HP_MAX_PARTIALS)) return -1;
    CK(cudaMallocHost((void**)&h->h_state, sizeof(HpState)));
    CK(cudaHostAlloc((void**)&h->h_fault, sizeof(int), cudaHostAllocMapped));
    *h->h_fault = 0;
    CK(cudaHostGetDevicePointer((void**)&d.fault_host, h->h_fault, 0));

    h->opts.precond = 1; h->opts.criterion = 0; h->opts.tol = 1e-9; h->opts.max_iter = 5000;
    h->opts.loop_mode = 1; h->opts.poll_chunk = 8; h->opts.refactor_every = 1; h->opts.extrapolate = 3;
    h->opts.mg_levels = 6; h->opts.mg_coarse_sweeps = 2; h->opts.mg_omega = 0.6; h->opts.reuse_converged = 0;
    h->opts.while_unroll = 1; h->opts.peer_timeout_ms = 60000; h->opts.mg_zscale = 0.5; h->opts.tma_rows = 0;
    h->opts.pipe_kernel = 1; h->opts.ext_guard_tols = 300.0;

    // initial field + snapshots (main.py:34,133,145,152)
    CK(hpk::launch_fill_rows(d, d.T, h->stream));
    CK(hpk::launch_fill_rows(d, d.Told, h->stream));
    HpKArgs a = make_args(h, 1.0, 0);
    {
        HpKernelLaunch k = hpk::desc_snapshot_kout(h->cfg, d);
        double* k0 = (double*)d.k0_out;
        void* params[] = {&d, &h->phys, &a, &k0};
        if (launch(h, k, params)) return -1;
        HpKernelLaunch kc = hpk::desc_coef(h->cfg, d);
        void* params2[] = {&d, &h->phys, &a};
        if (launch(h, kc, params2)) return -1;
    }
    CK(cudaStreamSynchronize(h->stream));
    guard.h = nullptr;
    *out = h;
    return 0;
}

static void destroy_graphs(hp_solver_s* h) {
    for (auto& g : h->graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        if (g.graph) cudaGraphDestroy(g.graph);
    }
    h->graphs.clear();
}

int hp_destroy(hp_handle h) {
    if (!h) return 0;
    cudaStreamSynchronize(h->stream);
    destroy_graphs(h);
    for (void* p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    if (h->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(h->nccl);
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    for (void* p : h->allocs) cudaFree(p);
    if (h->h_state) cudaFreeHost(h->h_state);
    if (h->h_fault) cudaFreeHost(h->h_fault);
    if (h->l2_limit_changed) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, h->l2_limit_before);
    delete h;
    return 0;
}

int hp_set_stream(hp_handle h, void* stream) {
    if (!h) return fail("null handle");
    CK(cudaStreamSynchronize(h->stream));
    h->stream = (cudaStream_t)stream;
    return 0;
}

// anything the host changes that a sweep depends on (T, T_old, solver options) ends the "next sweep repeats the last one"
// state kept on the device
static int invalidate_noop(hp_solver_s* h) {
    if (h->d.st) CK(cudaMemsetAsync(&h->d.st->noop_ok, 0, sizeof(int32_t), h->stream));
    return 0;
}

int hp_set_opts(hp_handle h, const hp_solve_opts* o) {
    if (!h || !o) return fail("hp_set_opts: null argument");
    if (o->precond < 0 || o->precond > 2) return fail("hp_set_opts: precond must be 0 (jacobi), 1 (rline) or 2 (mgline)");
    if (o->criterion < 0 || o->criterion > 2) return fail("hp_set_opts: criterion must be 0, 1 or 2");
    if (!(o->tol > 0.0)) return fail("hp_set_opts: tol must be > 0");
    if (o->extrapolate < 0 || o->extrapolate > 3) return fail("hp_set_opts: extrapolate must be 0 (off) .. 3 (cubic)");
    if (o->max_iter < 1) return fail("hp_set_opts: max_iter must be >= 1");
    if (o->loop_mode < 0 || o->loop_mode > 1) return fail("hp_set_opts: loop_mode must be 0 or 1");
    h->opts = *o;
    h->req_precond = o->precond;
    if (h->opts.poll_chunk < 1) h->opts.poll_chunk = 8;
    if (h->opts.refactor_every < 1) h->opts.refactor_every = 1;
    if (h->opts.mg_levels < 2) h->opts.mg_levels = 6;
    if (h->opts.mg_levels > HP_MAX_LEVELS) h->opts.mg_levels = HP_MAX_LEVELS;
    if (h->opts.mg_coarse_sweeps < 0) h->opts.mg_coarse_sweeps = 2;
    if (h->opts.mg_coarse_sweeps > 8) h->opts.mg_coarse_sweeps = 8;
    if (!(h->opts.mg_omega > 0.0 && h->opts.mg_omega <= 1.0)) h->opts.mg_omega = 0.6;
    if (!(h->opts.mg_zscale > 0.0 && h->opts.mg_zscale <= 1.0)) h->opts.mg_zscale = 0.5;
    if (h->opts.while_unroll < 1) h->opts.while_unroll = 1;
    if (h->opts.while_unroll > 8) h->opts.while_unroll = 8;
    if (h->opts.peer_timeout_ms < 1) h->opts.peer_timeout_ms = 60000;
    if (!(h->opts.ext_guard_tols >= 0.0)) h->opts.ext_guard_tols = 300.0;
    if (resolve_mg(h)) return -1;
    return invalidate_noop(h);
}

int hp_get_opts(hp_handle h, hp_solve_opts* out) {
    if (!h || !out) return fail("hp_get_opts: null argument");
    *out = h->opts;
    return 0;
}

static size_t owned_bytes(hp_handle h) { return (size_t)h->d.nz * h->d.nr * sizeof(double); }


int hp_set_T_dev(hp_handle h, const double* d_T) {
    if (!h || !d_T) return fail("hp_set_T_dev: null argument");
    CK(cudaMemcpyAsync(h->d.T + h->d.nr, d_T, owned_bytes(h), cudaMemcpyDeviceToDevice, h->stream));
    h->ghost_dirty = true;
    h->hist = 0; h->c2_stashed = false; h->pipe_hist = 0;
    return invalidate_noop(h);
}
int hp_get_T_dev(hp_handle h, double* d_T) {
    if (!h || !d_T) return fail("hp_get_T_dev: null argument");
    if (check_fault(h)) return -1;
    CK(cudaMemcpyAsync(d_T, h->d.T + h->d.nr, owned_bytes(h), cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}
int hp_set_T_host(hp_handle h, const double* h_T) {
    if (!h || !h_T) return fail("hp_set_T_host: null argument");
    CK(cudaMemcpyAsync(h->d.T + h->d.nr, h_T, owned_bytes(h), cudaMemcpyHostToDevice, h->stream));
    h->ghost_dirty = true;
    h->hist = 0; h->c2_stashed = false; h->pipe_hist = 0;
    return invalidate_noop(h);
}
int hp_get_T_host(hp_handle h, double* h_T) {
    if (!h || !h_T) return fail("hp_get_T_host: null argument");
    CK(cudaMemcpyAsync(h_T, h->d.T + h->d.nr, owned_bytes(h), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return check_fault(h);
}
const double* hp_T_ptr(hp_handle h) { return h ? h->d.T + h->d.nr : nullptr; }

int hp_update_old(hp_handle h) {
    if (!h) return fail("null handle");
    CK(cudaMemcpyAsync(h->d.Told, h->d.T, (size_t)(h->d.nz + 2) * h->d.nr * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return invalidate_noop(h);
}

// ---- slab halo exchange / scalar all-gather (NCCL) ----------------------------------------------------
static int exchange_rows(hp_solver_s* h, double* field) {
    if (h->nranks == 1) return 0;
    const int nr = h->d.nr, nz = h->d.nz;
    NK(g_nccl.GroupStart());
    if (h->rank > 0) {
        NK(g_nccl.Send(field + (size_t)1 * nr, nr, ncclFloat64, h->rank - 1, h->nccl, h->stream));
        NK(g_nccl.Recv(field, nr, ncclFloat64, h->rank - 1, h->nccl, h->stream));
    }
    if (h->rank < h->nranks - 1) {
        NK(g_nccl.Send(field + (size_t)nz * nr, nr, ncclFloat64, h->rank + 1, h->nccl, h->stream));
        NK(g_nccl.Recv(field + (size_t)(nz + 1) * nr, nr, ncclFloat64, h->rank + 1, h->nccl, h->stream));
    }
    NK(g_nccl.GroupEnd());
    return 0;
}

// ---- emission of the sweep either as plain stream launches or as CUDA-graph kernel nodes --------------------------------
struct NodeChain { cudaGraph_t g; cudaGraphNode_t last = nullptr; int count = 0; };
static int chain_kernel(NodeChain& c, const HpKernelLaunch& k, void** params, const hp_solver_s* h = nullptr) {
    cudaGraphNode_t n;
    cudaKernelNodeParams np{};
    np.func = (void*)k.func; np.gridDim = k.grid; np.blockDim = k.block; np.sharedMemBytes = (unsigned)k.smem;
    np.kernelParams = params; np.extra = nullptr;
    CK(cudaGraphAddKernelNode(&n, c.g, c.last ? &c.last : nullptr, c.last ? 1 : 0, &np));
    if (h && h->l2_on) {
        cudaKernelNodeAttrValue v;
        memset(&v, 0, sizeof(v));
        v.accessPolicyWindow = h->l2_window;
        CK(cudaGraphKernelNodeSetAttribute(n, cudaKernelNodeAttributeAccessPolicyWindow, &v));
    }
    c.last = n; c.count++;
    return 0;
}

struct Emitter {
    hp_solver_s* h;
    NodeChain* chain;        // nullptr: launch on the handle's stream
};
static int emit(Emitter& e, const HpKernelLaunch& k, void** params, int kind = KIND_OTHER) {
    if (e.chain) return chain_kernel(*e.chain, k, params, e.h);
    return launch(e.h, k, params, kind);
}

// all ranks' per-kernel sums -> every rank, then one thread advances the scalar recurrences identically everywhere.
// fin_kind: 0 diag, 1 A, 2 B, 3 init, 4 end of the mgline V-cycle (FIN_* of hp_kernels.cu)
static int emit_combine(Emitter& e, HpKArgs& a, int fin_kind) {
    hp_solver_s* h = e.h;
    if (h->nranks == 1) return 0;
    void* p4[] = {&h->d, &h->phys, &a, &fin_kind};
    if (h->d.p2p) return emit(e, hpk::desc_wait(), p4);          // sums already travelled by peer stores
    if (e.chain) return fail("NCCL slab exchange cannot be captured: use loop_mode 0 or the peer-memory exchange");
    NK(g_nccl.AllGather(h->d.red_local, h->d.red_all, HP_NRED, ncclFloat64, h->nccl, h->stream));
    return emit(e, hpk::desc_finalize(), p4);
}

// mgline: coarse operators and line pivots of the coarse levels (after the fine diagonal / pivots are current).  Needs no
// communication: with the slab cuts aligned to the row pairs every coarse coefficient is a function of this rank's rows.
static int emit_mg_setup(Emitter& e, bool always) {
    hp_solver_s* h = e.h;
    HpDev& d = h->d;
    int nr = d.nr;
    // skipped on the device when the INIT decision found the sweep converged -- unless later sweeps reuse this set-up
    // (refactor_every > 1): then it must exist, so it runs unconditionally (_pad0 is a constant zero)
    const int* skip = always ? &d.st->_pad0 : &d.st->mg_skip;
    HpLevelSet ls;
    ls.n = 1;
    ls.lev[0] = h->lev[0];
    int rows = 0;
    double zs = h->opts.mg_zscale;
    int keep_ghost = d.mg_local ? 0 : 1;
    for (int l = 1; l < h->nlev; ++l) {
        void* pc[] = {&h->lev[l - 1], &h->lev[l], &nr, &zs, &keep_ghost, &skip};
        if (emit(e, hpk::desc_mg_coarsen(h->cfg, h->lev[l].nz, nr), pc, KIND_MG_SETUP)) return -1;
        ls.lev[ls.n++] = h->lev[l];
        rows += h->lev[l].nz;
    }
    if (rows == 0) return 0;
    int first = 1;
    void* pf[] = {&ls, &first, &nr, &skip};
    return emit(e, hpk::desc_factor_rows(h->cfg, rows), pf, KIND_MG_SETUP);
}

// mgline V(1,1) cycle on r (level-0 rhs) starting from zline = M^-1 r in d.z (left there by k_cg_B); leaves the
// preconditioned residual in d.z and finalises (r.z, beta) in the last kernel.
//   down leg  l = 0..c-1 : residual of x_l (level 0: omega zline), restricted to rhs_{l+1}, x_{l+1} = omega M^-1 rhs_{l+1}
//   coarsest  level c    : `mg_coarse_sweeps` further smoothing steps, ping-pong between x_c and x2_c
//   up leg    l = c-1..1 : x2_l = xt + omega M^-1 (rhs_l - L xt),  xt = x_l + P e_{l+1}
//   level 0              : z = xt + omega M^-1 (r - L xt), r.z
// Fused stages (one row team per row, k_mg_fused<2|1>) when the launch shapes allow, else residual + line-solve launches.
// Under peer-memory slabs every stage that produces a correction hands its boundary rows to the neighbours (HpXchg) and
// the stage that reads them waits for the flag: the arrays a stage pushes into are never read by the stage running at the
// same time on the neighbour (the up leg and the coarsest sweeps write x2 / ping-pong, the down leg writes x).
static int emit_vcycle(Emitter& e, const HpKArgs& a_in, int init_phase) {
    hp_solver_s* h = e.h;
    HpDev& d = h->d;
    HpKArgs a = a_in;
    a.init_phase = init_phase;
    const size_t nr = d.nr;
    HpLevel* L = h->lev;
    const int nl = h->nlev, c = nl - 1;
    const bool dist = h->nranks > 1 && d.p2p;
    const bool upf = h->mg_upfused;
    if (c + h->opts.mg_coarse_sweeps + c > HP_MAX_STAGES) return fail("mgline: too many stages for the flag table");
    int next_slot = 0;
    auto ckind = [&](int l) -> int { return l < 1 ? KIND_MG_RES0 : KIND_MG_L1 + (l > 5 ? 5 : l) - 1; };
    // exchange descriptor of a stage that writes `field` (LV_X / LV_X2) of level l and reads ghost rows delivered by `wait`
    auto xout = [&](int l, int field, int wait) -> HpXchg {
        HpXchg x{nullptr, nullptr, -1, dist ? wait : -1};
        if (!dist) return x;
        x.sig = HP_COMM_STAGE0 + next_slot++;
        if (h->rank > 0) x.push_lo = h->lv_lo + h->lay_lo.off[l][field] + (size_t)(h->lay_lo.nz[l] + 1) * nr;
        if (h->rank < h->nranks - 1) x.push_hi = h->lv_hi + h->lay_hi.off[l][field];
        return x;
    };
    auto xwait = [&](int wait) -> HpXchg { return HpXchg{nullptr, nullptr, -1, dist ? wait : -1}; };
    auto res = [&](int l, const double* xin, const double* ec, int nzc, double* xo, double* resout, double* crhs, double scale,
                   int xo_ghost, HpXchg xc) -> int {
        HpMgResArgs m{xin, ec, xo, resout, crhs, scale, nzc, nullptr, nullptr, nullptr, xo_ghost, xc};
        void* p[] = {&d, &a, &L[l], &m};
        return emit(e, hpk::desc_mg_res(h->cfg, L[l].nz), p, ckind(l));
    };
    auto line = [&](int l, const double* rhs, const double* base, double* out, int fin, HpXchg xc) -> int {
        HpMgLineArgs m{rhs, base, out, fin, xc};
        void* p[] = {&d, &a, &L[l], &m};
        return emit(e, hpk::desc_mg_line(h->cfg, L[l].nz), p, l == 0 ? KIND_MG_LINE0 : ckind(l));
    };
    auto field_of = [&](int l, const double* p) -> int { return p == L[l].x ? LV_X : LV_X2; };
    int sig = -1;                       // slot of the stage that delivered the ghost rows of the array read next
    // ---- down leg
    for (int l = 0; l < c; ++l) {
        HpLevel& Ln = L[l + 1];
        const double* xin = l == 0 ? d.z : L[l].x;
        const double scale = l == 0 ? a.omega : 1.0;
        double* xo = l == 0 ? L[0].x2 : nullptr;
        const int wait = l == 0 ? -1 : sig;             // level 0: the z ghost rows came with k_cg_B and its k_wait
        if (upf) {
            HpXchg xc = xout(l + 1, LV_X, wait);
            HpMgResArgs m{xin, nullptr, xo, nullptr, Ln.rhs, scale, Ln.nz, Ln.dinv, Ln.cR, Ln.x, l == 0 ? 1 : 0, xc};
            void* p[] = {&d, &a, &L[l], &m};
            if (emit(e, hpk::desc_mg_down(h->cfg, L[l].nz), p, ckind(l))) return -1;
            sig = xc.sig;
        } else {
            if (res(l, xin, nullptr, Ln.nz, xo, nullptr, Ln.rhs, scale, l == 0 ? 1 : 0, xwait(wait))) return -1;
            HpXchg xc = xout(l + 1, LV_X, -1);
            if (line(l + 1, Ln.rhs, nullptr, Ln.x, 0, xc)) return -1;
            sig = xc.sig;
        }
    }
    // ---- coarsest level
    const double* cur = L[c].x;
    for (int k = 0; k < h->opts.mg_coarse_sweeps; ++k) {
        double* other = (cur == L[c].x) ? L[c].x2 : L[c].x;
        if (upf) {
            HpXchg xc = xout(c, field_of(c, other), sig);
            HpMgResArgs m{cur, nullptr, other, nullptr, nullptr, 1.0, 1, nullptr, nullptr, nullptr, 0, xc};
            void* p[] = {&d, &a, &L[c], &m};
            if (emit(e, hpk::desc_mg_up(h->cfg, L[c].nz), p, ckind(c))) return -1;
            sig = xc.sig;
        } else {
            if (res(c, cur, nullptr, 1, nullptr, L[c].tmp, nullptr, 1.0, 0, xwait(sig))) return -1;
            HpXchg xc = xout(c, field_of(c, other), -1);
            if (line(c, L[c].tmp, cur, other, 0, xc)) return -1;
            sig = xc.sig;
        }
        cur = other;
    }
    // ---- up leg
    for (int l = c - 1; l >= 1; --l) {
        if (upf) {
            HpXchg xc = xout(l, LV_X2, sig);
            HpMgResArgs m{L[l].x, cur, L[l].x2, nullptr, nullptr, 1.0, L[l + 1].nz, nullptr, nullptr, nullptr, 0, xc};
            void* p[] = {&d, &a, &L[l], &m};
            if (emit(e, hpk::desc_mg_up(h->cfg, L[l].nz), p, ckind(l))) return -1;
            sig = xc.sig;
        } else {
            if (res(l, L[l].x, cur, L[l + 1].nz, L[l].x2, L[l].tmp, nullptr, 1.0, 0, xwait(sig))) return -1;
            HpXchg xc = xout(l, LV_X2, -1);
            if (line(l, L[l].tmp, L[l].x2, L[l].x2, 0, xc)) return -1;
            sig = xc.sig;
        }
        cur = L[l].x2;
    }
    // ---- level 0: z = xt + omega M^-1 (r - L xt); its boundary rows travel with the r.z reduction
    if (upf && h->cfg.clF == 1) {
        // one pass: prolong + residual + line smoothing straight into z, r.z accumulated on the way (62 us instead of
        // 40 + 37 at 4096 x 1024).  Not on cluster teams (nr > 1024): there the column-tiled residual kernel and the
        // 512 x 8 line kernel stream at 5.3-5.5 TB/s and the fused pass at 3.1 (164 us against 88 + 78 at 2048 x 4096)
        HpXchg xf{dist ? d.z_lo : nullptr, dist ? d.z_hi : nullptr, -1, dist ? sig : -1};
        HpMgResArgs m{L[0].x2, cur, d.z, nullptr, nullptr, 1.0, L[1].nz, nullptr, nullptr, nullptr, 0, xf};
        void* p[] = {&d, &a, &L[0], &m};
        if (emit(e, hpk::desc_mg_final(h->cfg, L[0].nz), p, KIND_MG_LINE0)) return -1;
    } else {
        if (res(0, L[0].x2, cur, L[1].nz, d.z, L[0].tmp, nullptr, 1.0, 0, xwait(sig))) return -1;
        HpXchg xf{dist ? d.z_lo : nullptr, dist ? d.z_hi : nullptr, -1, -1};
        if (line(0, L[0].tmp, d.z, d.z, 1, xf)) return -1;
    }
    HpKArgs af = a;
    return emit_combine(e, af, 4);
}

// one PCG iteration: [ghost rows of z] A, B, [V-cycle]
static int emit_iteration(Emitter& e, HpKArgs& a) {
    hp_solver_s* h = e.h;
    HpDev& d = h->d;
    void* params[] = {&d, &h->phys, &a};
    // mgline: the cycle that turns z_line (left by the INIT kernel, the shift pair or the previous k_cg_B) into the
    // preconditioned residual OPENS the body: a sweep that is converged at its INIT decision, and the pass after the last
    // k_cg_B, launch no V-cycle kernels at all
    if (a.mg && emit_vcycle(e, a, 0)) return -1;
    if (h->nranks > 1 && !d.p2p) {                       // p2p: the producer stored the rows into the ghosts itself
        if (e.chain) return fail("NCCL slab exchange cannot be captured");
        if (exchange_rows(h, d.z)) return -1;
    }
    if (emit(e, hpk::desc_cg_A(h->cfg, d), params, KIND_A)) return -1;
    if (emit_combine(e, a, 1)) return -1;
    if (emit(e, hpk::desc_cg_B(h->cfg, d, h->opts.precond != 0, h->opts.criterion, h->opts.tma_rows), params, KIND_B)) return -1;
    if (emit_combine(e, a, 2)) return -1;
    return 0;
}

// everything of a sweep before the PCG loop: [coef] diag [factor] [mg setup] and the start of the solve (INIT kernel, or
// the shift pair of the extrapolated start); `a` carries the WHILE handle in graph mode
static int emit_sweep_head(Emitter& e, HpKArgs& a, bool refactor, bool shift) {
    hp_solver_s* h = e.h;
    HpDev& d = h->d;
    void* params[] = {&d, &h->phys, &a};
    if (h->phys.gamma == 1 && emit(e, hpk::desc_coef(h->cfg, d), params, KIND_COEF)) return -1;
    int write_dinv = (h->opts.precond == 0) ? 1 : 0;
    HpKArgs a0 = a;
    a0.use_cond = 0;
    int tprl = hpk::diag_tpr_log2(d.nr);
    void* p40[] = {&d, &h->phys, &a0, &write_dinv, &tprl};
    if (emit(e, hpk::desc_diag(h->cfg, d), p40, KIND_DIAG)) return -1;
    if (emit_combine(e, a0, 0)) return -1;
    if (h->opts.precond >= 1 && refactor) {
        HpLevelSet fine;
        fine.n = 1;
        fine.lev[0] = HpLevel{d.nz, d.cR, d.cZ, d.diag, d.dinv, nullptr, nullptr, nullptr, nullptr};
        int first = 0, nr_ = d.nr;
        const int* never = &d.st->_pad0;
        void* pf[] = {&fine, &first, &nr_, &never};
        if (emit(e, hpk::desc_factor_rows(h->cfg, d.nz), pf, KIND_FACTOR)) return -1;
    }
    if (shift) {
        // z holds T - T_prev (k_delta): one (A, B) pair with alpha := 1 moves the start of the solve to the extrapolated
        // field, r0 -> r0 - L z, and finalises like the INIT kernel
        HpKArgs as = a;
        as.shift = 1;
        void* ps[] = {&d, &h->phys, &as};
        if (h->nranks > 1 && !d.p2p) {
            if (e.chain) return fail("NCCL slab exchange cannot be captured");
            if (exchange_rows(h, d.z)) return -1;
        }
        if (emit(e, hpk::desc_cg_A(h->cfg, d), ps, KIND_A)) return -1;
        if (emit_combine(e, as, 1)) return -1;
        if (emit(e, hpk::desc_cg_B(h->cfg, d, h->opts.precond != 0, h->opts.criterion, h->opts.tma_rows), ps, KIND_B)) return -1;
        if (emit_combine(e, as, 3)) return -1;
        if (a.mg && refactor && emit_mg_setup(e, h->opts.refactor_every > 1)) return -1;
    } else {
        if (emit(e, hpk::desc_cg_init(h->cfg, d, h->opts.precond != 0, h->opts.criterion), params, KIND_INIT)) return -1;
        if (emit_combine(e, a, 3)) return -1;
        if (a.mg && refactor && emit_mg_setup(e, h->opts.refactor_every > 1)) return -1;
    }
    return 0;
}

// ---- one sweep with plain stream launches, host-polled PCG loop ------------------------------------------
// Under axial slabs (nranks > 1): ghost rows of T before assembly, ghost rows of z before every direction
// update, and the CG scalars combined over ranks after every reducing kernel.
// stash: 0 no, 1 second sweep of a step: c2 <- T, 2 the same and start from T + c2 (predicted correction)
static int sweep_stream(hp_solver_s* h, double dt, int has_old, bool refactor, bool shift = false, int stash = 0) {
    HpDev& d = h->d;
    HpKArgs a = make_args(h, dt, has_old);
    a.rhs_out = h->rhs_dbg;
    Emitter e{h, nullptr};
    if (apply_l2_to_stream(h)) return -1;
    if (h->nranks > 1 && (!d.p2p || h->ghost_dirty)) { if (exchange_rows(h, d.T)) return -1; h->ghost_dirty = false; }
    if (stash) {
        HpKArgs as = a;
        as.c2_valid = stash == 2;
        void* ps[] = {&d, &h->phys, &as};
        if (launch(h, hpk::desc_stash(h->cfg, d), ps, KIND_OTHER)) return -1;
        if (stash == 2) shift = true;
    }
    if (emit_sweep_head(e, a, refactor, shift)) return -1;
    int launched = 0;
    while (true) {
        CK(cudaMemcpyAsync(h->h_state, d.st, sizeof(HpState), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (h->h_state->done) break;
        if (launched >= h->opts.max_iter + h->opts.poll_chunk) return fail("sweep_stream: PCG loop did not terminate");
        for (int k = 0; k < h->opts.poll_chunk; ++k) {
            if (emit_iteration(e, a)) return -1;
            launched++;
        }
    }
    h->stream_mode_iters += h->h_state->iter;
    h->sweeps_host++;
    return 0;
}

// ---- CUDA graph: [delta] [Told <- T] + sweeps x ( head, WHILE{iteration} ) ---------------------------------------------
static bool same_opts(const hp_solve_opts& a, const hp_solve_opts& b) {
    return a.precond == b.precond && a.criterion == b.criterion && a.tol == b.tol && a.max_iter == b.max_iter &&
           a.loop_mode == b.loop_mode && a.refactor_every == b.refactor_every && a.extrapolate == b.extrapolate &&
           a.mg_levels == b.mg_levels && a.mg_coarse_sweeps == b.mg_coarse_sweeps && a.mg_omega == b.mg_omega &&
           a.reuse_converged == b.reuse_converged && a.while_unroll == b.while_unroll && a.mg_zscale == b.mg_zscale &&
           a.peer_timeout_ms == b.peer_timeout_ms && a.tma_rows == b.tma_rows && a.pipe_kernel == b.pipe_kernel && a.ext_guard_tols == b.ext_guard_tols;
}

static int get_graph(hp_solver_s* h, double dt, int sweeps, int has_old, int update_old, GraphEntry** out) {
    for (auto& g : h->graphs)
        if (g.dt == dt && g.sweeps == sweeps && g.has_old == has_old && g.update_old == update_old && same_opts(g.opts, h->opts)) {
            *out = &g;
            return 0;
        }
    if (h->graphs.size() >= 16) { CK(cudaStreamSynchronize(h->stream)); destroy_graphs(h); }
    GraphEntry ge;
    ge.dt = dt; ge.sweeps = sweeps; ge.has_old = has_old; ge.update_old = update_old; ge.opts = h->opts;
    CK(cudaGraphCreate(&ge.graph, 0));
    struct GraphGuard { cudaGraph_t g; ~GraphGuard() { if (g) cudaGraphDestroy(g); } } gguard{ge.graph};   // node-building failures
    NodeChain c{ge.graph};
    Emitter e{h, &c};
    HpDev& d = h->d;
    // update_old: 0 nothing, 1 Told <- T, 1 + k: z <- increment extrapolated at order k (start of the first solve) and
    // Told <- T in one kernel
    // bit 3: c2 holds the first-sweep solution of the previous step; bit 4: this step tracks it (k_stash at its second sweep)
    const int base = update_old & 7;
    const bool c2v = (update_old & 8) != 0, track = (update_old & 16) != 0;
    if (base >= 2) {
        HpKArgs a0 = make_args(h, dt, has_old);
        a0.ext_order = base - 1;
        a0.c2_valid = c2v ? 1 : 0;
        void* p0[] = {&d, &h->phys, &a0};
        if (chain_kernel(c, hpk::desc_delta(h->cfg, d), p0)) return -1;
    }
    if (base == 1) {
        cudaGraphNode_t n;
        CK(cudaGraphAddMemcpyNode1D(&n, ge.graph, c.last ? &c.last : nullptr, c.last ? 1 : 0, d.Told, d.T,
                                    (size_t)(d.nz + 2) * d.nr * sizeof(double), cudaMemcpyDeviceToDevice));
        c.last = n;
        // T_old moved: the first sweep of this step is not a repeat of the last sweep of the previous one
        cudaMemsetParams mp{};
        mp.dst = &d.st->noop_ok; mp.value = 0; mp.elementSize = 4; mp.width = 1; mp.height = 1; mp.pitch = 4;
        cudaGraphNode_t mn;
        CK(cudaGraphAddMemsetNode(&mn, ge.graph, &c.last, 1, &mp));
        c.last = mn;
    }
    for (int s = 0; s < sweeps; ++s) {
        HpKArgs a = make_args(h, dt, has_old);
        cudaGraphConditionalHandle cond;
        CK(cudaGraphConditionalHandleCreate(&cond, ge.graph, 0, cudaGraphCondAssignDefault));
        a.use_cond = 1; a.cond = (unsigned long long)cond;
        if (track && s == 1) {
            HpKArgs as = a;
            as.c2_valid = c2v ? 1 : 0;
            void* ps[] = {&d, &h->phys, &as};
            if (chain_kernel(c, hpk::desc_stash(h->cfg, d), ps)) return -1;
        }
        if (emit_sweep_head(e, a, (s % h->opts.refactor_every) == 0, (base >= 2 && s == 0) || (track && c2v && s == 1))) return -1;
        cudaGraphNodeParams cp{};
        cp.type = cudaGraphNodeTypeConditional;
        cp.conditional.handle = cond;
        cp.conditional.type = cudaGraphCondTypeWhile;
        cp.conditional.size = 1;
        cudaGraphNode_t wn;
        CK(cudaGraphAddNode(&wn, ge.graph, c.last ? &c.last : nullptr, c.last ? 1 : 0, &cp));
        c.last = wn;
        NodeChain body{cp.conditional.phGraph_out[0]};
        Emitter eb{h, &body};
        // while_unroll = U iterations per pass of the WHILE node: evaluating the node costs ~18 us, the kernels of iterations
        // past convergence return at once (the protocol of the host-polled chunks)
        const int unroll = h->opts.while_unroll;
        for (int u = 0; u < unroll; ++u)
            if (emit_iteration(eb, a)) return -1;
        ge.body_nodes = body.count / unroll;
    }
    ge.static_nodes = c.count;
    {
        CK(cudaGraphInstantiate(&ge.exec, ge.graph, 0));
    }
    gguard.g = nullptr;
    h->graphs.push_back(ge);
    *out = &h->graphs.back();
    return 0;
}

// is_step = 0: a bare sweep (hp_sweep; the caller manages T_old).  is_step = 1: whole time steps (hp_step / hp_run):
// Told <- T at the step start when has_old or extrapolate is set; with extrapolate and completed previous steps the
// first solve starts from the field extrapolated from them (k_delta).
static int run_steps(hp_solver_s* h, double dt, int n_steps, int sweeps, int has_old, int is_step) {
    if (!(dt > 0.0)) return fail("dt must be > 0");
    if (sweeps < 1) return fail("sweeps must be >= 1");
    if (check_fault(h)) return -1;
    // Small uncoupled pipes (the native 500 x 9 mesh, ensembles of it) with the line preconditioner: one CTA per pipe
    // runs the whole step on chip (k_pipe_step) instead of ~150 latency-bound launches / one HBM pass per PCG iteration
    if (h->opts.pipe_kernel && h->pipe_nzm > 0 && h->nranks == 1 && h->opts.precond == 1 && h->phys.gamma == 0 &&
        !h->rhs_dbg && !h->prof_on && sweeps <= 64) {
        HpDev& d = h->d;
        if (h->pipe_part_sweeps < sweeps) {
            if (dev_alloc(h, (void**)&h->pipe_part, (size_t)sweeps * h->pipe_M * 4 * sizeof(double))) return -1;
            h->pipe_part_sweeps = sweeps;
        }
        HpKArgs a = make_args(h, dt, has_old);
        // whole steps keep Told = field at the step start (like the row-kernel path with `extrapolate`), so the next step
        // can start its first solve from T + (T - Told): linear prediction, when the previous step used the same dt
        int nzm = h->pipe_nzm, M = h->pipe_M, sw = sweeps, upd = is_step ? 1 : 0;
        if (is_step && !has_old) a.has_old = 0;      // hasOld=False: the transient term uses the iterate, Told is history only
        double* part = h->pipe_part;
        const HpKernelLaunch ks = hpk::desc_pipe_step(d.nr, nzm, M), kf = hpk::desc_pipe_finalize();
        if (h->pipe_hist_dt != dt) { h->pipe_hist = 0; h->pipe_c2_valid = false; }
        if (h->pipe_hist == 0) h->pipe_c2_valid = false;
        if (is_step && sweeps >= 2 && h->opts.extrapolate >= 1 && !d.c2) {
            const size_t bytes = h->field_elems * sizeof(double);
            if (dev_alloc(h, (void**)&d.c2, bytes)) return -1;
            CK(cudaMemsetAsync(d.c2, 0, bytes, h->stream));
        }
        if (is_step && h->opts.extrapolate >= 1 && !d.d1) {
            const size_t bytes = h->field_elems * sizeof(double);
            if (dev_alloc(h, (void**)&d.d1, bytes) || dev_alloc(h, (void**)&d.d2, bytes)) return -1;
            CK(cudaMemsetAsync(d.d1, 0, bytes, h->stream));
            CK(cudaMemsetAsync(d.d2, 0, bytes, h->stream));
        }
        for (int s = 0; s < n_steps; ++s) {
            // order of the predicted start: linear once one step of history exists (Told), quadratic once the increment of
            // the step before has been stored too (d1, written by every predicting step)
            int predict = (is_step && h->opts.extrapolate >= 1 && h->pipe_hist >= 1) ? 1 : 0;
            if (predict && h->opts.extrapolate >= 2 && h->pipe_hist >= 3 && d.d1) predict = 2;
            // re-sweep correction (second-sweep start): bit0 keep S1 / form C this step, bit1 C of the previous step is valid
            int track = (is_step && sweeps >= 2 && h->opts.extrapolate >= 1 && d.c2 && d.d2) ? 1 : 0;
            if (track && h->pipe_c2_valid) track |= 2;
            void* p1[] = {&d, &h->phys, &a, &nzm, &sw, &upd, &predict, &track, &part};
            if (launch(h, ks, p1)) return -1;
            void* p2[] = {&d, &a, &M, &sw, &part};
            if (launch(h, kf, p2)) return -1;
            h->sweeps_host += sweeps;
            if (is_step) { if (h->pipe_hist < 3) h->pipe_hist++; h->pipe_hist_dt = dt; } else h->pipe_hist = 0;
            h->pipe_c2_valid = (track & 1) != 0;
        }
        h->hist = 0; h->c2_stashed = false;          // the higher-order history of the row-kernel path is not maintained here
        return 0;
    }
    const int max_order = is_step ? (h->opts.extrapolate > 3 ? 3 : h->opts.extrapolate) : 0;
    const bool ext = max_order > 0;
    if (ext && h->hist_dt != dt) { if (h->hist > 1) h->hist = 1; h->hist_dt = dt; }     // increments scale with dt
    if (max_order >= 2 && !h->d.d1) {
        const size_t bytes = h->field_elems * sizeof(double);
        if (dev_alloc(h, (void**)&h->d.d1, bytes) || dev_alloc(h, (void**)&h->d.d2, bytes)) return -1;
        CK(cudaMemsetAsync(h->d.d1, 0, bytes, h->stream));
        CK(cudaMemsetAsync(h->d.d2, 0, bytes, h->stream));
    }
    // The order ramps up with the history: right after an impulsive start the higher differences are large and a
    // high-order prediction is worse than a low-order one (measured with the oracle: quadratic pays from the 5th step,
    // cubic from the 8th).  A poor prediction only costs PCG iterations, never accuracy.
    auto order = [&]() {
        int o = h->hist >= 1 ? 1 : 0;
        if (max_order >= 2 && h->hist >= 4) o = 2;
        if (max_order >= 3 && h->hist >= 7) o = 3;
        return o;
    };
    // re-sweep correction tracking (see k_delta / k_stash): needs a second sweep to stash the first-sweep solution at
    const bool track = ext && sweeps >= 2;
    if (track && !h->d.c2) {
        const size_t bytes = h->field_elems * sizeof(double);
        if (dev_alloc(h, (void**)&h->d.c2, bytes)) return -1;
        CK(cudaMemsetAsync(h->d.c2, 0, bytes, h->stream));
    }
    if (!track) h->c2_stashed = false;
    // bits 0-2: 0 nothing, 1 Told <- T, 1 + k extrapolated start of order k; bit 3: c2 usable; bit 4: tracking on
    auto step_mode = [&]() {
        const int base = !is_step ? 0 : (ext ? 1 + order() : (has_old ? 1 : 0));
        return base | ((track && h->c2_stashed && base >= 2) ? 8 : 0) | (track ? 16 : 0);
    };
    auto step_done = [&]() { if (ext && h->hist < 1000) h->hist++; h->c2_stashed = track; };
    if (h->opts.loop_mode == 1 && (h->nranks == 1 || h->d.p2p) && !h->rhs_dbg && !h->prof_on) {
        if (h->nranks > 1 && h->ghost_dirty) { if (exchange_rows(h, h->d.T)) return -1; h->ghost_dirty = false; }
        for (int s = 0; s < n_steps; ++s) {
            GraphEntry* g = nullptr;
            if (get_graph(h, dt, sweeps, has_old, step_mode(), &g)) return -1;
            CK(cudaGraphLaunch(g->exec, h->stream));
            h->launches_host += g->static_nodes;
            h->graph_body_nodes = g->body_nodes;
            h->sweeps_host += sweeps;
            step_done();
        }
        return 0;
    }
    for (int s = 0; s < n_steps; ++s) {
        const int key = step_mode(), mode = key & 7;
        const bool c2v = (key & 8) != 0;
        if (mode >= 2) {
            HpKArgs a0 = make_args(h, dt, has_old);
            a0.ext_order = mode - 1;
            a0.c2_valid = c2v ? 1 : 0;
            void* p0[] = {&h->d, &h->phys, &a0};
            if (launch(h, hpk::desc_delta(h->cfg, h->d), p0, KIND_OTHER)) return -1;
        } else if (mode == 1 && hp_update_old(h)) return -1;
        for (int k = 0; k < sweeps; ++k)
            if (sweep_stream(h, dt, has_old, (k % h->opts.refactor_every) == 0, mode >= 2 && k == 0,
                             (track && k == 1) ? (c2v ? 2 : 1) : 0)) return -1;
        step_done();
    }
    return 0;
}

int hp_sweep(hp_handle h, double dt, int has_old, double* host_residual) {
    if (!h) return fail("null handle");
    if (run_steps(h, dt, 1, 1, has_old, 0)) return -1;
    if (host_residual) {
        hp_sweep_stat st;
        if (hp_get_stats(h, &st, 1) < 1) return fail("hp_sweep: no statistics available");
        *host_residual = st.residual_l2;
    }
    return 0;
}

int hp_step(hp_handle h, double dt, int sweeps, int has_old) {
    if (!h) return fail("null handle");
    return run_steps(h, dt, 1, sweeps, has_old, 1);
}

int hp_run(hp_handle h, double dt, int n_steps, int sweeps, int has_old) {
    if (!h) return fail("null handle");
    if (n_steps < 0) return fail("n_steps must be >= 0");
    return run_steps(h, dt, n_steps, sweeps, has_old, 1);
}

int hp_run_host(hp_handle h, const double* h_T_in, double* h_T_out, double dt, int n_steps, int sweeps, int has_old) {
    if (!h || !h_T_in || !h_T_out) return fail("hp_run_host: null argument");
    // the uploaded field continues the transient: the previous step start kept in Told stays the history of the
    // extrapolated PCG start (it only seeds the solver; a stale history costs iterations, never accuracy)
    const int keep = h->hist, keep_pipe = h->pipe_hist;
    const bool keep_c2 = h->c2_stashed;
    if (hp_set_T_host(h, h_T_in)) return -1;
    h->hist = keep; h->c2_stashed = keep_c2; h->pipe_hist = keep_pipe;
    if (hp_run(h, dt, n_steps, sweeps, has_old)) return -1;
    return hp_get_T_host(h, h_T_out);
}

int hp_get_stats(hp_handle h, hp_sweep_stat* out, int max_n) {
    if (!h || !out || max_n < 1) { fail("hp_get_stats: bad argument"); return -1; }
    if (fetch_state(h)) return -1;
    if (check_fault(h)) return -1;
    long long total = h->h_state->sweep_count;
    long long n = total < max_n ? total : max_n;
    if (n > HP_STATS_CAP) n = HP_STATS_CAP;
    for (long long k = 0; k < n; ++k) {
        long long slot = (total - n + k) % HP_STATS_CAP;
        if (cudaMemcpy(&out[k], h->d.stats + slot, sizeof(hp_sweep_stat), cudaMemcpyDeviceToHost) != cudaSuccess) {
            fail("hp_get_stats: copy failed");
            return -1;
        }
    }
    return (int)n;
}

int64_t hp_total_sweeps(hp_handle h) { return h ? h->sweeps_host : -1; }
int64_t hp_total_iters(hp_handle h) {
    if (!h || fetch_state(h)) return -1;
    return h->h_state->total_iters;
}
int64_t hp_kernel_launches(hp_handle h) {
    if (!h || fetch_state(h)) return -1;
    long long graph_iters = h->h_state->total_iters - h->stream_mode_iters - h->h_state->pipe_iters;
    return h->launches_host + (long long)h->graph_body_nodes * graph_iters;
}

int hp_assemble_debug(hp_handle h, double dt, int has_old, double* d_cR, double* d_cZ, double* d_diag, double* d_rhs) {
    if (!h) return fail("null handle");
    if (!(dt > 0.0)) return fail("dt must be > 0");
    HpDev& d = h->d;
    double* scratch = nullptr;
    CK(cudaMalloc((void**)&scratch, h->field_elems * sizeof(double)));
    HpKArgs a = make_args(h, dt, has_old);
    a.rhs_out = scratch;
    void* params[] = {&d, &h->phys, &a};
    int rc = 0;
    if (h->phys.gamma == 1) rc = launch(h, hpk::desc_coef(h->cfg, d), params);
    if (!rc) {
        int write_dinv = 0, tprl = hpk::diag_tpr_log2(d.nr);
        void* p4[] = {&d, &h->phys, &a, &write_dinv, &tprl};
        rc = launch(h, hpk::desc_diag(h->cfg, d), p4);
    }
    const long long n = (long long)d.nz * d.nr;
    cudaError_t e = cudaSuccess;
    if (!rc && d_cR) e = hpk::launch_copy_owned(d.cR, d_cR, d.nr, n, h->stream);
    if (!rc && e == cudaSuccess && d_cZ) e = hpk::launch_copy_owned(d.cZ, d_cZ, d.nr, n, h->stream);
    if (!rc && e == cudaSuccess && d_diag) e = hpk::launch_copy_owned(d.diag, d_diag, d.nr, n, h->stream);
    if (!rc && e == cudaSuccess && d_rhs) e = hpk::launch_copy_owned(scratch, d_rhs, d.nr, n, h->stream);
    cudaStreamSynchronize(h->stream);
    cudaFree(scratch);
    if (rc) return -1;
    if (e != cudaSuccess) return fail(std::string("hp_assemble_debug: ") + cudaGetErrorString(e));
    return 0;
}

int hp_eval_property(hp_handle h, int prop_id, const double* d_T, double* d_out, int64_t n) {
    if (!h || !d_T || !d_out) return fail("hp_eval_property: null argument");
    CK(hpk::launch_eval_property(h->phys, prop_id, d_T, d_out, n, h->stream));
    h->launches_host++;
    return 0;
}

int hp_eval_property_raw(const hp_phys* phys, int prop_id, const double* d_T, double* d_out, int64_t n, void* stream) {
    if (!phys || !d_T || !d_out) return fail("hp_eval_property_raw: null argument");
    CK(hpk::launch_eval_property(*phys, prop_id, d_T, d_out, n, (cudaStream_t)stream));
    return 0;
}

int hp_eval_cell_fields(hp_handle h, double* d_rho, double* d_cp, double* d_kavg) {
    if (!h) return fail("null handle");
    CK(hpk::launch_cell_fields(h->d, h->phys, d_rho, d_cp, d_kavg, h->stream));
    h->launches_host++;
    return 0;
}

// ---- peer memory over NVLink: every rank maps every other rank's mailbox, and its neighbours' z / T arrays -------
int hp_peer_export(hp_handle h, const int32_t* nz_of_rank, void* out256) {
    if (!h || !out256 || !nz_of_rank) return fail("hp_peer_export: null argument");
    if (h->nranks < 2) return fail("hp_peer_export: call hp_comm_init first");
    HpDev& d = h->d;
    if (!h->comm) {
        if (dev_alloc(h, (void**)&h->comm, sizeof(HpComm))) return -1;
        CK(cudaMemset(h->comm, 0, sizeof(HpComm)));
        if (dev_alloc(h, (void**)&d.seq, sizeof(unsigned long long) * HP_COMM_KINDS)) return -1;
        CK(cudaMemset(d.seq, 0, sizeof(unsigned long long) * HP_COMM_KINDS));
    }
    h->nz_of_rank.assign(nz_of_rank, nz_of_rank + h->nranks);
    if (nz_of_rank[h->rank] != d.nz) return fail("hp_peer_export: nz_of_rank[rank] differs from this handle's row count");
    if (ensure_levels(h)) return -1;            // the coarse-level block the neighbours push their boundary rows into
    CK(cudaStreamSynchronize(h->stream));
    cudaIpcMemHandle_t hd[4];
    CK(cudaIpcGetMemHandle(&hd[0], h->comm));
    CK(cudaIpcGetMemHandle(&hd[1], h->qz_block));
    CK(cudaIpcGetMemHandle(&hd[2], d.T));
    CK(cudaIpcGetMemHandle(&hd[3], h->lvblock));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(out256, hd, sizeof(hd));
    return 0;
}

int hp_peer_import(hp_handle h, const void* all_handles, const int32_t* nz_of_rank) {
    if (!h || !all_handles || !nz_of_rank) return fail("hp_peer_import: null argument");
    if (h->nranks < 2) return fail("hp_peer_import: call hp_comm_init first");
    if (h->nranks > HP_MAX_RANKS) return fail("hp_peer_import: too many ranks");
    if (!h->comm) return fail("hp_peer_import: call hp_peer_export first");
    CK(cudaStreamSynchronize(h->stream));
    const cudaIpcMemHandle_t* hd = (const cudaIpcMemHandle_t*)all_handles;
    HpDev& d = h->d;
    const size_t nr = d.nr;
    for (int r = 0; r < h->nranks; ++r) {
        if (r == h->rank) { d.comm_peer[r] = h->comm; continue; }
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, hd[4 * r + 0], cudaIpcMemLazyEnablePeerAccess));
        h->ipc_opened.push_back(p);
        d.comm_peer[r] = (HpComm*)p;
    }
    h->nz_of_rank.assign(nz_of_rank, nz_of_rank + h->nranks);
    h->lv_lo = h->lv_hi = nullptr;
    d.z_lo = d.z_hi = d.T_lo = d.T_hi = nullptr;
    auto open_nb = [&](int r, double** qz, double** T, double** lv) -> int {
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, hd[4 * r + 1], cudaIpcMemLazyEnablePeerAccess));
        h->ipc_opened.push_back(p); *qz = (double*)p;
        CK(cudaIpcOpenMemHandle(&p, hd[4 * r + 2], cudaIpcMemLazyEnablePeerAccess));
        h->ipc_opened.push_back(p); *T = (double*)p;
        CK(cudaIpcOpenMemHandle(&p, hd[4 * r + 3], cudaIpcMemLazyEnablePeerAccess));
        h->ipc_opened.push_back(p); *lv = (double*)p;
        return 0;
    };
    if (h->rank > 0) {                       // lower neighbour: its UPPER ghost row (nz_lo + 1) mirrors my row 1
        const int r = h->rank - 1;
        double *qz, *T;
        if (open_nb(r, &qz, &T, &h->lv_lo)) return -1;
        h->lay_lo = level_layout(nz_of_rank[r], d.nr);
        const size_t felems = (size_t)(nz_of_rank[r] + 2) * nr + 64;
        d.z_lo = qz + felems + (size_t)(nz_of_rank[r] + 1) * nr;
        d.T_lo = T + (size_t)(nz_of_rank[r] + 1) * nr;
    }
    if (h->rank < h->nranks - 1) {           // upper neighbour: its LOWER ghost row (0) mirrors my row nz
        const int r = h->rank + 1;
        double *qz, *T;
        if (open_nb(r, &qz, &T, &h->lv_hi)) return -1;
        h->lay_hi = level_layout(nz_of_rank[r], d.nr);
        const size_t felems = (size_t)(nz_of_rank[r] + 2) * nr + 64;
        d.z_hi = qz + felems;
        d.T_hi = T;
    }
    d.p2p = 1;
    h->ghost_dirty = true;
    destroy_graphs(h);
    return resolve_mg(h);                                    // the cycle becomes the distributed whole-mesh hierarchy
}

int hp_comm_fault(hp_handle h) { return (h && h->h_fault && *(volatile int*)h->h_fault) ? 1 : 0; }

int hp_comm_clear_fault(hp_handle h) {
    if (!h) return fail("null handle");
    CK(cudaStreamSynchronize(h->stream));
    if (h->comm) {
        CK(cudaMemset(h->comm, 0, sizeof(HpComm)));
        CK(cudaMemset(h->d.seq, 0, sizeof(unsigned long long) * HP_COMM_KINDS));
    }
    if (fetch_state(h)) return -1;
    HpState st = *h->h_state;
    st.flags &= ~8; st.cycle = 0; st.ticket = 0; st.done = 1; st.mg_skip = 1;
    CK(cudaMemcpy(h->d.st, &st, sizeof(HpState), cudaMemcpyHostToDevice));
    if (h->h_fault) *(volatile int*)h->h_fault = 0;
    h->ghost_dirty = true;
    return 0;
}

int hp_profile_begin(hp_handle h, int max_launches) {
    if (!h) return fail("null handle");
    if (max_launches < 1) max_launches = 1;
    while (h->prof_ev.size() < (size_t)2 * max_launches) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        h->prof_ev.push_back(e);
    }
    h->prof_kind.clear();
    h->prof_used = 0;
    h->prof_on = true;
    return 0;
}

int hp_profile_end(hp_handle h, hp_kernel_times* out) {
    if (!h || !out) return fail("hp_profile_end: null argument");
    h->prof_on = false;
    CK(cudaStreamSynchronize(h->stream));
    memset(out, 0, sizeof(*out));
    std::vector<float> ms(h->prof_used, 0.f);
    float longest[16] = {0};
    for (size_t k = 0; k < h->prof_used; ++k) {
        CK(cudaEventElapsedTime(&ms[k], h->prof_ev[2 * k], h->prof_ev[2 * k + 1]));
        int kind = h->prof_kind[k];
        out->count[kind] += 1;
        out->total_ms[kind] += ms[k];
        if (ms[k] > longest[kind]) longest[kind] = ms[k];
    }
    for (size_t k = 0; k < h->prof_used; ++k) {
        int kind = h->prof_kind[k];
        if (ms[k] >= 0.5f * longest[kind]) { out->full_count[kind] += 1; out->full_ms[kind] += ms[k]; }
    }
    return 0;
}

int hp_comm_unique_id(void* id128) {
    if (!id128) return fail("hp_comm_unique_id: null argument");
    if (!load_nccl()) return fail("hp_comm_unique_id: libnccl.so.2 not found");
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int hp_comm_init(hp_handle h, int rank, int nranks, const void* id128) {
    if (!h || !id128) return fail("hp_comm_init: null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail("hp_comm_init: bad rank / nranks");
    if (!load_nccl()) return fail("hp_comm_init: libnccl.so.2 not found");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NK(g_nccl.CommInitRank(&h->nccl, nranks, id, rank));
    h->rank = rank; h->nranks = nranks;
    if (dev_alloc(h, (void**)&h->d.red_local, sizeof(double) * HP_NRED)) return -1;
    if (dev_alloc(h, (void**)&h->d.red_all, sizeof(double) * HP_NRED * nranks)) return -1;
    CK(cudaMemsetAsync(h->d.red_local, 0, sizeof(double) * HP_NRED, h->stream));
    h->d.nranks = nranks; h->d.rank = rank;
    destroy_graphs(h);
    if (resolve_mg(h)) return -1;          // slabs without peer memory: the V-cycle turns rank-local
    // the snapshot couplings across slab interfaces were assembled from the T_amb-filled ghost rows at create
    // time (both neighbours compute the same value); nothing to exchange here.
    return 0;
}

}  // extern "C"
