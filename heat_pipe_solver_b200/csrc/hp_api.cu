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
    // 'mgline' hierarchy (level 0 aliases the fine fields)
    HpLevel lev[HP_MAX_LEVELS];
    int nlev = 0, nlev_alloc = 0;
    // replicated global coarse levels of mgline under peer-memory slabs (shape fixed at hp_peer_export time)
    std::vector<int> nz_of_rank;
    int g_switch = 0;                // local level whose rows are gathered into global level 0 (0 = rank-local cycle)
    int g_rows_mine = 0, g_row0 = 0; // my rows / first row inside global level 0
    int nglev = 0;
    HpLevel glev[HP_MAX_LEVELS];
    long long g_off[HP_MAX_LEVELS][8];   // element offsets of (cR, cZ, diag, dinv, rhs, x, x2, tmp) inside gblock
    double* gblock = nullptr;
    size_t gblock_elems = 0;
    bool g_active = false;
    // peer-memory exchange (hp_peer_export / hp_peer_import)
    HpComm* comm = nullptr;          // my mailbox
    double* qz_block = nullptr;      // base of the (q, z) allocation (IPC handle is per allocation)
    std::vector<void*> ipc_opened;
    unsigned long long* gbar = nullptr;   // grid-barrier words of k_mg_subcycle
    bool mg_fused = true;
    bool mg_upfused = true;      // coarse up-leg stages as one kernel (k_mg_up) instead of residual + line solve
    bool pdl = false;                 // programmatic dependent launch edges between the kernel nodes of the step graph
    int hist = 0;                    // completed steps since the field was last set: Told (and d1, d2) hold that much history
    double hist_dt = 0.0;            // time step the history was taken with
    bool c2_on = true;               // track the re-sweep correction (predicted start of the second sweep); HP_SWEEP2=0 disables
    bool c2_stashed = false;         // d.c2 holds the first-sweep solution of the step just finished
    bool ghost_dirty = true;         // ghost rows of T must be refreshed by an explicit exchange before the next sweep
    // L2 access-policy window over the contiguous (q, z) block
    bool l2_on = false;
    void* l2_base = nullptr;
    size_t l2_bytes = 0;
    cudaAccessPolicyWindow l2_window{};
    // per-kernel event timing (hp_profile_begin / hp_profile_end; stream-launch mode only)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<int> prof_kind;
    size_t prof_used = 0;
};

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
// graphs).  HP_L2_PERSIST=0 disables it (A/B measurements).
static void setup_l2_window(hp_solver_s* h, const cudaDeviceProp& prop) {
    h->l2_on = false;
    const char* env = getenv("HP_L2_PERSIST");
    const int mode = env ? atoi(env) : 1;
    if (mode == 0) return;
    if (prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) return;
    // Only when the whole window fits the persisting carve-out: a partially persisting window (hitRatio < 1) over
    // vectors larger than L2 halves the streaming rate of both PCG kernels (measured on 16384x4096, profiles/).
    size_t win = h->l2_bytes;                                   // q and z
    const size_t cap = (size_t)prop.persistingL2CacheMaxSize < (size_t)prop.accessPolicyMaxWindowSize
                           ? (size_t)prop.persistingL2CacheMaxSize : (size_t)prop.accessPolicyMaxWindowSize;
    if (win > cap) {
        if (mode == 2 && win / 2 <= cap) win = win / 2;        // experiment: q alone
        else return;
    }
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, win) != cudaSuccess) { cudaGetLastError(); return; }
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

// ---- 'mgline' hierarchy: level l+1 has ceil(nz_l / 2) rows; stops at 16 rows or mg_levels -------------------
static int local_level_count(int nz, int max_levels) {
    int n = 1;
    for (int nzl = nz; n < max_levels && nzl >= 32; nzl = (nzl + 1) / 2) ++n;
    return n;
}
static int rows_at_level(int nz, int l) {
    for (int k = 0; k < l; ++k) nz = (nz + 1) / 2;
    return nz;
}

static int setup_mg(hp_solver_s* h) {
    HpDev& d = h->d;
    int want = local_level_count(d.nz, h->opts.mg_levels);
    // under peer-memory slabs every rank stops at the same level, and the rows of that level are gathered into the
    // replicated global hierarchy whose shape was fixed at export time
    h->g_active = false;
    if (h->nranks > 1 && d.p2p && h->gblock && !h->nz_of_rank.empty()) {
        int s_min = HP_MAX_LEVELS;
        for (int r = 0; r < h->nranks; ++r) {
            int c = local_level_count(h->nz_of_rank[r], h->opts.mg_levels) - 1;
            if (c < s_min) s_min = c;
        }
        if (s_min >= 1 && s_min == h->g_switch) { want = s_min + 1; h->g_active = true; }
    }
    if (h->nlev_alloc == 0) {
        HpLevel& L0 = h->lev[0];
        L0.nz = d.nz; L0.cR = d.cR; L0.cZ = d.cZ; L0.diag = d.diag; L0.dinv = d.dinv;
        L0.rhs = d.r; L0.x = d.z; L0.tmp = d.q;
        if (dev_alloc(h, (void**)&L0.x2, h->field_elems * sizeof(double))) return -1;
        CK(cudaMemsetAsync(L0.x2, 0, h->field_elems * sizeof(double), h->stream));
        if (dev_alloc(h, (void**)&h->gbar, 2 * sizeof(unsigned long long))) return -1;
        CK(cudaMemsetAsync(h->gbar, 0, 2 * sizeof(unsigned long long), h->stream));
        const char* ef = getenv("HP_MG_FUSED");
        // measured on B200: inside the step graph the fused cycle (132 us) is no faster than its 21 separate launches
        // (~6 us each), both are bound by the chain of dependent row latencies -> off unless asked for
        h->mg_fused = (ef && ef[0] == '1') && hpk::mg_subcycle_supported(h->cfg);
        const char* eu = getenv("HP_MG_UPFUSE");
        h->mg_upfused = !(eu && eu[0] == '0') && hpk::mg_subcycle_supported(h->cfg) && h->cfg.ncolA == 1;
        h->nlev_alloc = 1;
    }
    for (int l = h->nlev_alloc; l < want; ++l) {
        HpLevel& L = h->lev[l];
        L.nz = (h->lev[l - 1].nz + 1) / 2;
        const size_t elems = (size_t)(L.nz + 2) * d.nr + 64;
        double** f[] = {&L.cR, &L.cZ, &L.diag, &L.dinv, &L.rhs, &L.x, &L.x2, &L.tmp};
        for (double** p : f) {
            if (dev_alloc(h, (void**)p, elems * sizeof(double))) return -1;
            CK(cudaMemsetAsync(*p, 0, elems * sizeof(double), h->stream));
        }
        h->nlev_alloc = l + 1;
    }
    h->nlev = want;
    return 0;
}

static HpKArgs make_args(const hp_solver_s* h, double dt, int has_old) {
    HpKArgs a{};
    a.dt = dt; a.tol = h->opts.tol; a.has_old = has_old; a.max_iter = h->opts.max_iter;
    a.criterion = h->opts.criterion; a.use_cond = 0; a.cond = 0; a.rhs_out = nullptr;
    static const int pf = [] { const char* e = getenv("HP_PREFETCH"); return e ? atoi(e) : 0; }();
    a.prefetch = pf;
    a.mg = (h->opts.precond == 2) ? 1 : 0; a.init_phase = 0; a.omega = h->opts.mg_omega; a.shift = 0;   // measured slower on B200 (profiles/): off unless asked for
    // repeated converged sweeps as bookkeeping only (HpState.noop_ok, opt-in); under slabs the reducing kernels must run
    // on every rank for the flag protocol, so single GPU only
    a.noop_enable = (h->opts.reuse_converged && h->nranks == 1) ? 1 : 0;
    return a;
}

enum { KIND_OTHER = 0, KIND_COEF, KIND_DIAG, KIND_FACTOR, KIND_INIT, KIND_A, KIND_B, KIND_MG_RES0, KIND_MG_LINE0, KIND_MG_COARSE,
       KIND_MG_SETUP, KIND_MG_L1, KIND_MG_L2, KIND_MG_L3, KIND_MG_L4, KIND_MG_L5, KIND_COUNT };   // L5 = level 5 and deeper

static int launch(hp_solver_s* h, const HpKernelLaunch& k, void** params, int kind = KIND_OTHER) {
    const bool prof = h->prof_on && h->prof_used < h->prof_ev.size() / 2;
    if (prof) CK(cudaEventRecord(h->prof_ev[2 * h->prof_used], h->stream));
    if (k.cooperative) CK(cudaLaunchCooperativeKernel(k.func, k.grid, k.block, params, k.smem, h->stream));
    else CK(cudaLaunchKernel(k.func, k.grid, k.block, params, k.smem, h->stream));
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
    h->stream = (cudaStream_t)stream;
    h->phys = *phys;
    CK(cudaGetDevice(&h->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->device));
    h->num_sms = prop.multiProcessorCount;
    const char* err = nullptr;
    if (!hpk::pick_config(m->nr, m->nz, h->num_sms, &h->cfg, &err)) { delete h; return fail(std::string("hp_create: ") + err); }
    const int nr = m->nr, nz = m->nz;
    for (int i = 0; i < nr; ++i)
        if (!(m->r_v[i + 1] > m->r_v[i])) { delete h; return fail("hp_create: r_v must be strictly increasing"); }
    for (int g = 0; g < nz + 2; ++g)
        if (!(m->z_v[g + 1] > m->z_v[g])) { delete h; return fail("hp_create: z_v must be strictly increasing"); }

    HpDev& d = h->d;
    d.nr = nr; d.nz = nz; d.ncolA = h->cfg.ncolA; d.p2p = 0; d.seq = nullptr; d.z_lo = d.z_hi = d.T_lo = d.T_hi = nullptr;
    for (int r = 0; r < HP_MAX_RANKS; ++r) { d.comm_peer[r] = nullptr; d.gblock_peer[r] = nullptr; }
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
#define UP(field, vec) if (upload(h, &d.field, vec)) { hp_destroy(h); return -1; }
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
    h->field_elems = (size_t)(nz + 2) * nr + 64;
    double** fields[] = {&d.T, &d.Told, &d.cR, &d.cZ, &d.diag, &d.dinv, &d.r, &d.p0, &d.p1};
    for (double** f : fields) {
        if (dev_alloc(h, (void**)f, h->field_elems * sizeof(double))) { hp_destroy(h); return -1; }
        CK(cudaMemsetAsync(*f, 0, h->field_elems * sizeof(double), h->stream));
    }
    // q and z are produced by one PCG kernel and consumed by the next one: one contiguous block so that a single
    // L2 access-policy window can keep both resident across kernel boundaries (setup_l2_window)
    {
        double* qz = nullptr;
        if (dev_alloc(h, (void**)&qz, 2 * h->field_elems * sizeof(double))) { hp_destroy(h); return -1; }
        CK(cudaMemsetAsync(qz, 0, 2 * h->field_elems * sizeof(double), h->stream));
        d.q = qz; d.z = qz + h->field_elems;
        h->qz_block = qz;
        h->l2_base = qz; h->l2_bytes = 2 * h->field_elems * sizeof(double);
    }
    setup_l2_window(h, prop);
    void* p = nullptr;
    if (dev_alloc(h, &p, sizeof(double) * (nz + 2))) { hp_destroy(h); return -1; }
    d.k0_out = (const double*)p;
    if (dev_alloc(h, (void**)&d.st, sizeof(HpState))) { hp_destroy(h); return -1; }
    CK(cudaMemsetAsync(d.st, 0, sizeof(HpState), h->stream));
    if (dev_alloc(h, (void**)&d.stats, sizeof(hp_sweep_stat) * HP_STATS_CAP)) { hp_destroy(h); return -1; }
    CK(cudaMemsetAsync(d.stats, 0, sizeof(hp_sweep_stat) * HP_STATS_CAP, h->stream));
    if (dev_alloc(h, (void**)&d.partials, sizeof(double) * HP_NRED * HP_MAX_PARTIALS)) { hp_destroy(h); return -1; }
    CK(cudaMallocHost((void**)&h->h_state, sizeof(HpState)));

    h->opts.precond = 1; h->opts.criterion = 0; h->opts.tol = 1e-9; h->opts.max_iter = 5000;
    // measured on B200: programmatic edges make the step SLOWER (mgline 8.97 -> 10.70 ms, rline 18.2 -> 19.4 ms; the
    // early-scheduled dependents compete for SM slots with the draining grid) -> opt-in only (HP_PDL=1)
    { const char* e = getenv("HP_PDL"); h->pdl = (e && e[0] == '1'); }
    { const char* e = getenv("HP_SWEEP2"); h->c2_on = !(e && e[0] == '0'); }
    h->opts.loop_mode = 1; h->opts.poll_chunk = 8; h->opts.refactor_every = 1; h->opts.extrapolate = 3;
    h->opts.mg_levels = 6; h->opts.mg_coarse_sweeps = 2; h->opts.mg_omega = 0.7; h->opts.reuse_converged = 0;

    // initial field + snapshots (main.py:34,133,145,152)
    CK(hpk::launch_fill_rows(d, d.T, h->stream));
    CK(hpk::launch_fill_rows(d, d.Told, h->stream));
    HpKArgs a = make_args(h, 1.0, 0);
    {
        HpKernelLaunch k = hpk::desc_snapshot_kout(h->cfg, d);
        double* k0 = (double*)d.k0_out;
        void* params[] = {&d, &h->phys, &a, &k0};
        if (launch(h, k, params)) { hp_destroy(h); return -1; }
        HpKernelLaunch kc = hpk::desc_coef(h->cfg, d);
        void* params2[] = {&d, &h->phys, &a};
        if (launch(h, kc, params2)) { hp_destroy(h); return -1; }
    }
    CK(cudaStreamSynchronize(h->stream));
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
    if (h->opts.poll_chunk < 1) h->opts.poll_chunk = 8;
    if (h->opts.refactor_every < 1) h->opts.refactor_every = 1;
    if (h->opts.mg_levels < 2) h->opts.mg_levels = 6;
    if (h->opts.mg_levels > HP_MAX_LEVELS) h->opts.mg_levels = HP_MAX_LEVELS;
    if (h->opts.mg_coarse_sweeps < 0) h->opts.mg_coarse_sweeps = 2;
    if (!(h->opts.mg_omega > 0.0 && h->opts.mg_omega <= 1.0)) h->opts.mg_omega = 0.7;
    if (h->opts.precond == 2) {
        if (setup_mg(h)) return -1;
        if (h->nlev < 2) h->opts.precond = 1;          // mesh too short to coarsen: plain line preconditioner
    }
    return invalidate_noop(h);
}

static size_t owned_bytes(hp_handle h) { return (size_t)h->d.nz * h->d.nr * sizeof(double); }


int hp_set_T_dev(hp_handle h, const double* d_T) {
    if (!h || !d_T) return fail("hp_set_T_dev: null argument");
    CK(cudaMemcpyAsync(h->d.T + h->d.nr, d_T, owned_bytes(h), cudaMemcpyDeviceToDevice, h->stream));
    h->ghost_dirty = true;
    h->hist = 0; h->c2_stashed = false;
    return invalidate_noop(h);
}
int hp_get_T_dev(hp_handle h, double* d_T) {
    if (!h || !d_T) return fail("hp_get_T_dev: null argument");
    CK(cudaMemcpyAsync(d_T, h->d.T + h->d.nr, owned_bytes(h), cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}
int hp_set_T_host(hp_handle h, const double* h_T) {
    if (!h || !h_T) return fail("hp_set_T_host: null argument");
    CK(cudaMemcpyAsync(h->d.T + h->d.nr, h_T, owned_bytes(h), cudaMemcpyHostToDevice, h->stream));
    h->ghost_dirty = true;
    h->hist = 0; h->c2_stashed = false;
    return invalidate_noop(h);
}
int hp_get_T_host(hp_handle h, double* h_T) {
    if (!h || !h_T) return fail("hp_get_T_host: null argument");
    CK(cudaMemcpyAsync(h_T, h->d.T + h->d.nr, owned_bytes(h), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
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
struct NodeChain {
    cudaGraph_t g; cudaGraphNode_t last = nullptr; int count = 0;
    bool last_is_kernel = false;      // a programmatic (PDL) edge may only connect two kernel nodes
};
static int chain_kernel(NodeChain& c, const HpKernelLaunch& k, void** params, const hp_solver_s* h = nullptr) {
    cudaGraphNode_t n;
    const bool pdl = h && h->pdl && c.last && c.last_is_kernel && !k.cooperative;
    if (pdl) {
        // programmatic edge: the node may be scheduled while its predecessor drains; every kernel starts with
        // griddepcontrol.wait (pdl_enter), which restores full completion + memory visibility
        cudaGraphNodeParams gp{};
        gp.type = cudaGraphNodeTypeKernel;
        gp.kernel.func = (void*)k.func; gp.kernel.gridDim = k.grid; gp.kernel.blockDim = k.block;
        gp.kernel.sharedMemBytes = (unsigned)k.smem; gp.kernel.kernelParams = params; gp.kernel.extra = nullptr;
        cudaGraphEdgeData ed;
        memset(&ed, 0, sizeof(ed));
        ed.from_port = cudaGraphKernelNodePortProgrammatic;
        ed.type = cudaGraphDependencyTypeProgrammatic;
        CK(cudaGraphAddNode_v2(&n, c.g, &c.last, &ed, 1, &gp));
    } else {
        cudaKernelNodeParams np{};
        np.func = (void*)k.func; np.gridDim = k.grid; np.blockDim = k.block; np.sharedMemBytes = (unsigned)k.smem;
        np.kernelParams = params; np.extra = nullptr;
        CK(cudaGraphAddKernelNode(&n, c.g, c.last ? &c.last : nullptr, c.last ? 1 : 0, &np));
    }
    if (k.cooperative) {
        cudaKernelNodeAttrValue v;
        memset(&v, 0, sizeof(v));
        v.cooperative = 1;
        CK(cudaGraphKernelNodeSetAttribute(n, cudaKernelNodeAttributeCooperative, &v));
    }
    if (h && h->l2_on) {
        cudaKernelNodeAttrValue v;
        memset(&v, 0, sizeof(v));
        v.accessPolicyWindow = h->l2_window;
        CK(cudaGraphKernelNodeSetAttribute(n, cudaKernelNodeAttributeAccessPolicyWindow, &v));
    }
    c.last = n; c.count++;
    c.last_is_kernel = !k.cooperative;
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

// mgline: Galerkin operators and line pivots of the coarse levels (after the fine diagonal / pivots are current).
// Under peer-memory slabs the rows of the last local level are gathered into the replicated global level 0 of every
// rank (with the true interface couplings), which is then coarsened further on every rank redundantly.
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
    const int s = h->nlev - 1;
    for (int l = 1; l < h->nlev; ++l) {
        void* pc[] = {&h->lev[l - 1], &h->lev[l], &nr, &skip};
        if (emit(e, hpk::desc_mg_coarsen(h->cfg, h->lev[l].nz, nr), pc, KIND_MG_SETUP)) return -1;
        if (!(h->g_active && l == s)) { ls.lev[ls.n++] = h->lev[l]; rows += h->lev[l].nz; }
    }
    if (h->g_active) {
        HpGatherArgs g{};
        g.src[0] = h->lev[s].cR; g.src[1] = h->lev[s].cZ; g.src[2] = h->lev[s].diag;
        g.dst_off[0] = h->g_off[0][0]; g.dst_off[1] = h->g_off[0][1]; g.dst_off[2] = h->g_off[0][2];
        g.top_override = d.cZ + (size_t)d.nz * nr;      // fine coupling of my last row to the next slab (0 at the top end)
        g.override_idx = 1;
        g.nsrc = 3; g.nrows = h->g_rows_mine; g.row0_dst = h->g_row0; g.kind = HP_COMM_GOP;
        void* pg[] = {&d, &g, &skip};
        if (emit(e, hpk::desc_gather_rows(h->cfg, 3LL * g.nrows * nr), pg, KIND_MG_SETUP)) return -1;
        int kind = HP_COMM_GOP;
        void* pw[] = {&d, &kind, &skip};
        if (emit(e, hpk::desc_wait_flag(), pw, KIND_MG_SETUP)) return -1;
        ls.lev[ls.n++] = h->glev[0]; rows += h->glev[0].nz;
        for (int gl = 1; gl < h->nglev; ++gl) {
            void* pc[] = {&h->glev[gl - 1], &h->glev[gl], &nr, &skip};
            if (emit(e, hpk::desc_mg_coarsen(h->cfg, h->glev[gl].nz, nr), pc, KIND_MG_SETUP)) return -1;
            ls.lev[ls.n++] = h->glev[gl]; rows += h->glev[gl].nz;
        }
    }
    if (rows == 0) return 0;
    void* pf[] = {&ls, &nr, &skip};
    return emit(e, hpk::desc_factor_levels(rows), pf, KIND_MG_SETUP);
}

// mgline V(1,1) cycle on r (level-0 rhs) starting from zline = M^-1 r in d.z (left there by k_cg_B); leaves the
// preconditioned residual in d.z and finalises (r.z, beta) in the last kernel.
static int emit_vcycle(Emitter& e, const HpKArgs& a_in, int init_phase) {
    hp_solver_s* h = e.h;
    HpDev& d = h->d;
    HpKArgs a = a_in;
    a.init_phase = init_phase;
    const int nr = d.nr;
    // profiling kind of a coarse stage: local levels 1..5+ separately, replicated global levels under "mg_coarse"
    auto ckind = [&](const HpLevel& Lv) -> int {
        const long long l = &Lv - h->lev;
        if (l < 1 || l >= h->nlev) return KIND_MG_COARSE;
        return KIND_MG_L1 + (int)(l > 5 ? 5 : l) - 1;
    };
    auto res = [&](HpLevel& Lv, bool fine, const double* xin, const double* ec, int nzc, double* xout, double* resout,
                   double* crhs, double scale) -> int {
        HpMgResArgs m{xin, ec, xout, resout, crhs, scale, nzc};
        void* p[] = {&d, &a, &Lv, &m};
        return emit(e, hpk::desc_mg_res(h->cfg, Lv.nz), p, fine ? KIND_MG_RES0 : ckind(Lv));
    };
    auto line = [&](HpLevel& Lv, bool fine, const double* rhs, const double* base, double* out, int fin) -> int {
        HpMgLineArgs m{rhs, base, out, fin};
        void* p[] = {&d, &a, &Lv, &m};
        return emit(e, hpk::desc_mg_line(h->cfg, Lv.nz), p, fine ? KIND_MG_LINE0 : ckind(Lv));
    };
    // fused up-leg stage: out = xt + omega M^-1 (rhs - L xt), xt = xin + P ec   (one launch instead of res + line)
    auto up = [&](HpLevel& Lv, const double* xin, const double* ec, int nzc, double* out) -> int {
        HpMgResArgs m{xin, ec, out, nullptr, nullptr, 1.0, nzc};
        void* p[] = {&d, &a, &Lv, &m};
        return emit(e, hpk::desc_mg_up(h->cfg, Lv.nz), p, ckind(Lv));
    };
    const bool upf = h->mg_upfused;
    // down-leg stage of level Lv: residual of xin (scaled), restricted into Ln.rhs; with `presmooth` the pre-smoothing of
    // the next level, Ln.x = omega M^-1 Ln.rhs, rides in the same launch when the fused kernels are usable
    auto down = [&](HpLevel& Lv, bool fine, const double* xin, double scale, double* xout, HpLevel& Ln, bool presmooth) -> int {
        if (upf && presmooth) {
            HpMgResArgs m{xin, nullptr, xout, nullptr, Ln.rhs, scale, Ln.nz, Ln.dinv, Ln.cR, Ln.x};
            void* p[] = {&d, &a, &Lv, &m};
            return emit(e, hpk::desc_mg_down(h->cfg, Lv.nz), p, fine ? KIND_MG_RES0 : ckind(Lv));
        }
        if (res(Lv, fine, xin, nullptr, Ln.nz, xout, nullptr, Ln.rhs, scale)) return -1;
        return presmooth ? line(Ln, false, Ln.rhs, nullptr, Ln.x, 0) : 0;
    };
    // post-smoothing of level l with coarse correction ec; returns where the smoothed correction was left
    auto up_leg = [&](HpLevel& Lv, const double* ec, int nzc) -> const double* {
        if (upf) return up(Lv, Lv.x, ec, nzc, Lv.x2) ? nullptr : Lv.x2;
        if (res(Lv, false, Lv.x, ec, nzc, Lv.x2, Lv.tmp, nullptr, 1.0)) return nullptr;
        if (line(Lv, false, Lv.tmp, Lv.x2, Lv.x, 0)) return nullptr;
        return Lv.x;
    };
    // complete V-cycle on levels lo..n-1 of Ls (rhs of level lo given, Ls[lo].x = omega M^-1 rhs too if `presmoothed`);
    // *top = where the correction of level lo is left
    auto subcycle = [&](HpLevel* Ls, int lo, int n, bool presmoothed, const double** top) -> int {
        const int c = n - 1;
        if (!presmoothed && line(Ls[lo], false, Ls[lo].rhs, nullptr, Ls[lo].x, 0)) return -1;
        for (int l = lo; l < c; ++l)
            if (down(Ls[l], false, Ls[l].x, 1.0, nullptr, Ls[l + 1], true)) return -1;
        const double* cur = Ls[c].x;
        for (int k = 0; k < h->opts.mg_coarse_sweeps; ++k) {
            if (upf) {
                double* other = (cur == Ls[c].x) ? Ls[c].x2 : Ls[c].x;
                if (up(Ls[c], cur, nullptr, 1, other)) return -1;
                cur = other;
            } else {
                if (res(Ls[c], false, Ls[c].x, nullptr, 1, nullptr, Ls[c].tmp, nullptr, 1.0)) return -1;
                if (line(Ls[c], false, Ls[c].tmp, Ls[c].x, Ls[c].x, 0)) return -1;
            }
        }
        for (int l = c - 1; l >= lo; --l) {
            cur = up_leg(Ls[l], cur, Ls[l + 1].nz);
            if (!cur) return -1;
        }
        *top = cur;
        return 0;
    };
    // same cycle in ONE cooperative launch (grid barriers instead of kernel boundaries) when the row-team shapes allow
    auto fused = [&](HpLevel* Ls, int lo, int n, int part, const double* ec_top) -> int {
        HpLevelSet ls;
        ls.n = n;
        for (int l = 0; l < n; ++l) ls.lev[l] = Ls[l];
        HpSubcycleArgs sc{lo, n, h->opts.mg_coarse_sweeps, part, ec_top, h->gbar};
        void* p[] = {&d, &a, &ls, &sc};
        return emit(e, hpk::desc_mg_subcycle(h->cfg), p, KIND_MG_COARSE);
    };
    HpLevel* L = h->lev;
    const int nl = h->nlev;
    // level 0 starts from omega * zline (its pre-smoothing is k_cg_B's line solve)
    const int s = nl - 1;                  // g_active: rows of local level s are one slice of the replicated global level 0
    // level 1 is pre-smoothed here unless the cooperative cycle does it or it is the level handed to the global hierarchy
    const bool pre1 = !h->mg_fused && !(h->g_active && s == 1);
    if (down(L[0], true, d.z, a.omega, L[0].x2, L[1], pre1)) return -1;
    const double* ec0 = L[1].x;          // coarse correction seen by level 0
    if (!h->g_active) {
        if (h->mg_fused ? fused(L, 1, nl, 0, nullptr) : subcycle(L, 1, nl, true, &ec0)) return -1;
    } else {
        if (h->mg_fused && s > 1) {
            if (fused(L, 1, nl, 1, nullptr)) return -1;
        } else {
            for (int l = 1; l < s; ++l)
                if (down(L[l], false, L[l].x, 1.0, nullptr, L[l + 1], l + 1 < s)) return -1;
        }
        const int* skip = &d.st->mg_skip;
        HpGatherArgs g{};
        g.src[0] = L[s].rhs; g.dst_off[0] = h->g_off[0][4];
        g.top_override = nullptr; g.override_idx = -1;
        g.nsrc = 1; g.nrows = h->g_rows_mine; g.row0_dst = h->g_row0; g.kind = HP_COMM_GRHS;
        void* pg[] = {&d, &g, &skip};
        if (emit(e, hpk::desc_gather_rows(h->cfg, (long long)g.nrows * nr), pg, KIND_MG_COARSE)) return -1;
        int kind = HP_COMM_GRHS;
        void* pw[] = {&d, &kind, &skip};
        if (emit(e, hpk::desc_wait_flag(), pw, KIND_MG_COARSE)) return -1;
        const double* gtop = h->glev[0].x;
        if (h->mg_fused ? fused(h->glev, 0, h->nglev, 0, nullptr) : subcycle(h->glev, 0, h->nglev, false, &gtop)) return -1;
        const double* mine = gtop + (size_t)h->g_row0 * nr;              // my slice (ghost-inclusive indexing kept)
        if (h->mg_fused && s > 1) {
            if (fused(L, 1, nl, 2, mine)) return -1;
        } else {
            const double* cur = mine;
            for (int l = s - 1; l >= 1; --l) {
                cur = up_leg(L[l], cur, L[l + 1].nz);
                if (!cur) return -1;
            }
            if (s > 1) ec0 = cur;
        }
        if (s == 1) ec0 = mine;
    }
    if (res(L[0], true, L[0].x2, ec0, L[1].nz, d.z, L[0].tmp, nullptr, 1.0)) return -1;
    if (line(L[0], true, L[0].tmp, d.z, d.z, 1)) return -1;
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
    if (emit(e, hpk::desc_cg_B(h->cfg, d, h->opts.precond != 0, h->opts.criterion), params, KIND_B)) return -1;
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
        if (emit(e, hpk::desc_factor(h->cfg, d), params, KIND_FACTOR)) return -1;
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
        if (emit(e, hpk::desc_cg_B(h->cfg, d, h->opts.precond != 0, h->opts.criterion), ps, KIND_B)) return -1;
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
           a.reuse_converged == b.reuse_converged;
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
        c.last = n; c.last_is_kernel = false;
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
        c.last = wn; c.last_is_kernel = false;
        NodeChain body{cp.conditional.phGraph_out[0]};
        Emitter eb{h, &body};
        // HP_UNROLL=U (experiment, default 1): U iterations per pass of the WHILE node -- the evaluation of the node costs
        // ~18 us, the kernels of iterations past convergence return at once (the protocol of the host-polled chunks)
        static const int unroll = [] { const char* e = getenv("HP_UNROLL"); int u = e ? atoi(e) : 1; return u < 1 ? 1 : (u > 8 ? 8 : u); }();
        for (int u = 0; u < unroll; ++u)
            if (emit_iteration(eb, a)) return -1;
        ge.body_nodes = body.count / unroll;
    }
    ge.static_nodes = c.count;
    {
        cudaError_t ie = cudaGraphInstantiate(&ge.exec, ge.graph, 0);
        if (ie != cudaSuccess && (h->pdl || (h->mg_fused && h->opts.precond == 2))) {
            // programmatic edges / cooperative nodes not accepted in this graph shape: rebuild without them
            cudaGetLastError();
            cudaGraphDestroy(ge.graph);
            if (h->pdl) h->pdl = false; else h->mg_fused = false;
            return get_graph(h, dt, sweeps, has_old, update_old, out);
        }
        CK(ie);
    }
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
    const bool track = ext && sweeps >= 2 && h->c2_on;
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
    const int keep = h->hist;
    const bool keep_c2 = h->c2_stashed;
    if (hp_set_T_host(h, h_T_in)) return -1;
    h->hist = keep; h->c2_stashed = keep_c2;
    if (hp_run(h, dt, n_steps, sweeps, has_old)) return -1;
    return hp_get_T_host(h, h_T_out);
}

static int fetch_state(hp_solver_s* h) {
    CK(cudaMemcpyAsync(h->h_state, h->d.st, sizeof(HpState), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int hp_get_stats(hp_handle h, hp_sweep_stat* out, int max_n) {
    if (!h || !out || max_n < 1) { fail("hp_get_stats: bad argument"); return -1; }
    if (fetch_state(h)) return -1;
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
    long long graph_iters = h->h_state->total_iters - h->stream_mode_iters;
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
    // shape of the replicated global coarse hierarchy (identical on every rank: it only depends on nz_of_rank, nr, opts)
    h->nz_of_rank.assign(nz_of_rank, nz_of_rank + h->nranks);
    int s_min = HP_MAX_LEVELS;
    for (int r = 0; r < h->nranks; ++r) {
        int c = local_level_count(nz_of_rank[r], h->opts.mg_levels) - 1;
        if (c < s_min) s_min = c;
    }
    h->g_switch = s_min;
    int M = 0;
    for (int r = 0; r < h->nranks; ++r) {
        if (r == h->rank) h->g_row0 = M;
        M += rows_at_level(nz_of_rank[r], s_min);
    }
    h->g_rows_mine = rows_at_level(nz_of_rank[h->rank], s_min);
    h->nglev = 0;
    size_t elems = 0;
    for (int rows = M; h->nglev < HP_MAX_LEVELS; rows = (rows + 1) / 2) {
        h->glev[h->nglev].nz = rows;
        for (int k = 0; k < 8; ++k) { h->g_off[h->nglev][k] = (long long)elems; elems += (size_t)(rows + 2) * d.nr + 64; }
        h->nglev++;
        if (rows < 32) break;
    }
    if (!h->gblock) {
        h->gblock_elems = elems;
        if (dev_alloc(h, (void**)&h->gblock, elems * sizeof(double))) return -1;
        CK(cudaMemset(h->gblock, 0, elems * sizeof(double)));
    } else if (elems != h->gblock_elems) {
        return fail("hp_peer_export: called twice with different shapes");
    }
    for (int g = 0; g < h->nglev; ++g) {
        double** f[] = {&h->glev[g].cR, &h->glev[g].cZ, &h->glev[g].diag, &h->glev[g].dinv,
                        &h->glev[g].rhs, &h->glev[g].x, &h->glev[g].x2, &h->glev[g].tmp};
        for (int k = 0; k < 8; ++k) *f[k] = h->gblock + h->g_off[g][k];
    }
    cudaIpcMemHandle_t hd[4];
    CK(cudaIpcGetMemHandle(&hd[0], h->comm));
    CK(cudaIpcGetMemHandle(&hd[1], h->qz_block));
    CK(cudaIpcGetMemHandle(&hd[2], d.T));
    CK(cudaIpcGetMemHandle(&hd[3], h->gblock));
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
        if (r == h->rank) { d.comm_peer[r] = h->comm; d.gblock_peer[r] = h->gblock; continue; }
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, hd[4 * r + 0], cudaIpcMemLazyEnablePeerAccess));
        h->ipc_opened.push_back(p);
        d.comm_peer[r] = (HpComm*)p;
        CK(cudaIpcOpenMemHandle(&p, hd[4 * r + 3], cudaIpcMemLazyEnablePeerAccess));
        h->ipc_opened.push_back(p);
        d.gblock_peer[r] = (double*)p;
    }
    d.z_lo = d.z_hi = d.T_lo = d.T_hi = nullptr;
    auto open_nb = [&](int r, double** qz, double** T) -> int {
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, hd[4 * r + 1], cudaIpcMemLazyEnablePeerAccess));
        h->ipc_opened.push_back(p); *qz = (double*)p;
        CK(cudaIpcOpenMemHandle(&p, hd[4 * r + 2], cudaIpcMemLazyEnablePeerAccess));
        h->ipc_opened.push_back(p); *T = (double*)p;
        return 0;
    };
    if (h->rank > 0) {                       // lower neighbour: its UPPER ghost row (nz_lo + 1) mirrors my row 1
        const int r = h->rank - 1;
        double *qz, *T;
        if (open_nb(r, &qz, &T)) return -1;
        const size_t felems = (size_t)(nz_of_rank[r] + 2) * nr + 64;
        d.z_lo = qz + felems + (size_t)(nz_of_rank[r] + 1) * nr;
        d.T_lo = T + (size_t)(nz_of_rank[r] + 1) * nr;
    }
    if (h->rank < h->nranks - 1) {           // upper neighbour: its LOWER ghost row (0) mirrors my row nz
        const int r = h->rank + 1;
        double *qz, *T;
        if (open_nb(r, &qz, &T)) return -1;
        const size_t felems = (size_t)(nz_of_rank[r] + 2) * nr + 64;
        d.z_hi = qz + felems;
        d.T_hi = T;
    }
    d.p2p = 1;
    h->ghost_dirty = true;
    destroy_graphs(h);
    if (h->opts.precond == 2 && setup_mg(h)) return -1;      // switch the cycle to the replicated global coarse levels
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
    // the snapshot couplings across slab interfaces were assembled from the T_amb-filled ghost rows at create
    // time (both neighbours compute the same value); nothing to exchange here.
    return 0;
}

}  // extern "C"
