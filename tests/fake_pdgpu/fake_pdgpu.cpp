// tests/fake_pdgpu/fake_pdgpu.cpp -- TEST INFRASTRUCTURE ONLY (built and used by tests/test_cpp_driver_host_logic.py).
//
// A stand-in `libpdgpu.so` whose pdamr_* entry points are served by the compiled reference
// (oracle/_ref/libpdrefimp2d.so, through the C entry points of oracle/ref_shim.cpp).  Put in front of the real library
// with LD_LIBRARY_PATH, it lets the HOST logic of the C++ driver host/pd_corrosion_gpu (host/amr_run.cpp: cycle
// structure, batching between output points, snapshot cadence, PVD / CSV / VTU writing, D_map bookkeeping) run on a
// machine without a GPU; its output files must then equal those of the reference's own main() byte for byte.
// The operators themselves are what the -m gpu tests check on the device.  The lattice entry points (pdgpu_*) the
// drivers host/main.cpp + host/coupling.cpp call are served the same way for 2D single-rank runs; the remaining
// ones exist only so that the binary loads (it is linked BIND_NOW): fake_stubs.cpp, they fail when called.
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pdgpu.h"

namespace {
void* g_lib = nullptr;
std::string g_err = "fake libpdgpu: lattice entry points are not served";
template <class F>
F sym(const char* name) {
    if (!g_lib) {
        const char* p = std::getenv("PD_FAKE_REF_LIB");
        g_lib = dlopen(p ? p : "libpdrefimp2d.so", RTLD_NOW | RTLD_GLOBAL);
        if (!g_lib) { std::fprintf(stderr, "fake libpdgpu: %s\n", dlerror()); std::abort(); }
    }
    void* f = dlsym(g_lib, name);
    if (!f) { std::fprintf(stderr, "fake libpdgpu: missing %s\n", name); std::abort(); }
    return reinterpret_cast<F>(f);
}
#define REF0(name) sym<void (*)(void*)>(#name)
typedef void (*v_h)(void*);
}  // namespace

struct pdamr_ctx {                // also stands in for pdgpu_ctx (the lattice context)
    void* h = nullptr;
    std::string cfg_path;
    PdConfig cfg{};
};

static void apply(pdamr_ctx* c, const std::string& text) {
    const std::string p = c->cfg_path + ".apply";
    FILE* f = std::fopen(p.c_str(), "w");
    std::fputs(text.c_str(), f);
    std::fclose(f);
    sym<void (*)(void*, const char*)>("ref_config_apply")(c->h, p.c_str());
    std::remove(p.c_str());
}
static void dims(pdamr_ctx* c, long long* d) { sym<void (*)(void*, long long*)>("ref_get_dims")(c->h, d); }
static void* ptr(pdamr_ctx* c, const char* n) { return sym<void* (*)(void*, const char*)>("ref_ptr")(c->h, n); }

extern "C" {

const char* pdgpu_last_error(void) { return g_err.c_str(); }

static std::string write_cfg(const PdConfig* k, const std::string& extra) {
    char tmpl[] = "/tmp/pdfake_XXXXXX";
    const int fd = mkstemp(tmpl);
    if (fd < 0) return "";
    FILE* f = fdopen(fd, "w");
#define D(key) std::fprintf(f, #key " = %.17g\n", k->key)
#define I(key) std::fprintf(f, #key " = %d\n", k->key)
    D(dx); I(m_ratio); D(R_wire); D(L_wire); D(R_tube); D(L_upstream); D(L_downstream); D(rho_f); D(mu_f); D(gamma_eos);
    D(c0); D(eta_density); D(Q_flow); D(rho_m); D(D_liquid); D(D_grain); D(D_gb); D(D_precip); D(C_solid_init);
    D(C_liquid_init); D(C_thresh); D(C_sat); D(alpha_art_diff); D(corrosion_decay_l); D(cfl_factor); D(cfl_factor_corr);
    D(flow_conv_tol); D(T_final); I(flow_max_iters); I(corrosion_steps_per_check); I(output_every_flow);
    I(output_every_corr); I(channel_flow_corrections); I(use_implicit);
#undef D
#undef I
    std::fputs(extra.c_str(), f);
    std::fclose(f);
    sym<void (*)(int)>("ref_set_threads")(1);                       // the reference's in-place loops are order dependent
    return tmpl;
}

int pdamr_create(const PdConfig* k, int amr_ratio, double amr_buffer, pdamr_ctx** out) {
    pdamr_ctx* c = new pdamr_ctx();
    char extra[128];
    std::snprintf(extra, sizeof extra, "use_amr = 1\namr_ratio = %d\namr_buffer = %.17g\n", amr_ratio, amr_buffer);
    c->cfg_path = write_cfg(k, extra);
    if (c->cfg_path.empty()) return 1;
    c->h = sym<void* (*)(const char*)>("ref_create")(c->cfg_path.c_str());
    *out = c;
    return 0;
}
int pdamr_build(pdamr_ctx* c) { REF0(ref_grid_build_amr)(c->h); return 0; }
int pdamr_build_neighbors(pdamr_ctx* c) { REF0(ref_build_neighbors_celllist)(c->h); return 0; }
int pdamr_info(pdamr_ctx* c, PdAmrInfo* o) {
    std::memset(o, 0, sizeof(*o));
    long long d[5];
    dims(c, d);
    o->N_total = d[3]; o->nnz = d[4];
    o->n_fict_entries = sym<long long (*)(void*)>("ref_fict_entries")(c->h);
    const uint8_t* t = (const uint8_t*)ptr(c, "node_type");
    const int* lvl = (const int*)ptr(c, "grid_level");
    for (long long i = 0; i < d[3]; ++i) {
        o->counts[t[i]]++;
        if (t[i] == 6) o->n_fict++;
        else if (lvl[i] == 0) o->n_fine++;
        else o->n_coarse++;
    }
    return 0;
}
int pdamr_get(pdamr_ctx* c, const char* name, void* out) {
    long long d[5];
    dims(c, d);
    const long long N = d[3], nnz = d[4], nf = sym<long long (*)(void*)>("ref_fict_entries")(c->h);
    const std::string n(name);
    size_t bytes = 0;
    if (n == "pos") bytes = 16 * N; else if (n == "node_type") bytes = N;
    else if (n == "dx_local" || n == "delta_local") bytes = 8 * N; else if (n == "grid_level") bytes = 4 * N;
    else if (n == "fict_offset" || n == "nbr_offset") bytes = 4 * (N + 1);
    else if (n == "fict_source") bytes = 4 * nf; else if (n == "fict_weight") bytes = 8 * nf;
    else if (n == "nbr_index") bytes = 4 * nnz; else if (n == "nbr_dist" || n == "nbr_vol") bytes = 8 * nnz;
    else if (n == "nbr_evec") bytes = 16 * nnz;
    else { g_err = "fake pdamr_get: " + n; return 1; }
    std::memcpy(out, ptr(c, name), bytes);
    return 0;
}
int pdamr_device_init(pdamr_ctx* c, int) {
    REF0(ref_fields_init)(c->h);        // allocates the Fields (and draws the reference's own grains; the driver overwrites them)
    REF0(ref_ns_init)(c->h); REF0(ref_ard_init)(c->h); REF0(ref_imp_init)(c->h);
    return 0;
}
static size_t field_bytes(pdamr_ctx* c, const std::string& n) {
    long long d[5];
    dims(c, d);
    const size_t N = (size_t)d[3];
    if (n == "vel" || n == "vel_new") return 16 * N;
    if (n == "phase" || n == "is_gb" || n == "is_precip" || n == "node_type") return N;
    if (n == "rho" || n == "rho_new" || n == "C" || n == "C_new" || n == "pressure") return 8 * N;
    return 0;
}
int pdamr_field_set(pdamr_ctx* c, const char* name, const void* src) {
    const size_t b = field_bytes(c, name);
    if (!b) { g_err = std::string("fake pdamr_field_set: ") + name; return 1; }
    std::memcpy(ptr(c, name), src, b);
    return 0;
}
int pdamr_field_get(pdamr_ctx* c, const char* name, void* dst) {
    const size_t b = field_bytes(c, name);
    if (!b) { g_err = std::string("fake pdamr_field_get: ") + name; return 1; }
    std::memcpy(dst, ptr(c, name), b);
    return 0;
}
int pdamr_update_fictitious(pdamr_ctx* c) { REF0(ref_update_fictitious)(c->h); return 0; }
int pdamr_bc(pdamr_ctx* c, int which) {
    static const char* fn[] = {"ref_apply_inlet_bc", "ref_apply_outlet_bc", "ref_apply_wall_bc", "ref_apply_solid_surface_bc",
                               "ref_apply_wall_concentration_bc", "ref_apply_wall_bc_new", "ref_smooth_boundary_concentration"};
    if (which < 0 || which > 6) return 1;
    sym<v_h>(fn[which])(c->h);
    return 0;
}
int pdamr_ns_compute_dt(pdamr_ctx* c, double* dt) { *dt = sym<double (*)(void*)>("ref_ns_compute_dt")(c->h); return 0; }
int pdamr_ns_step(pdamr_ctx* c, double dt) { sym<void (*)(void*, double)>("ref_ns_step")(c->h, dt); return 0; }
int pdamr_ns_iterate(pdamr_ctx* c, int n, double dt) { sym<void (*)(void*, int, double)>("ref_ns_iterate_amr")(c->h, n, dt); return 0; }
int pdamr_ns_solve_steady(pdamr_ctx* c, PdSteadyResult* out, int) {
    std::memset(out, 0, sizeof(*out));
    out->iters = sym<int (*)(void*)>("ref_ns_solve_steady")(c->h);
    return 0;
}
int pdamr_ard_set_volume_loss(pdamr_ctx* c, double v) {
    sym<void (*)(void*, double)>("ref_ard_set_volume_loss")(c->h, v);
    sym<void (*)(void*, double)>("ref_imp_set_volume_loss")(c->h, v);
    return 0;
}
int pdamr_ard_compute_dt(pdamr_ctx* c, double* dt) { *dt = sym<double (*)(void*)>("ref_ard_compute_dt")(c->h); return 0; }
int pdamr_ard_step(pdamr_ctx* c, double dt) { sym<void (*)(void*, double)>("ref_ard_step")(c->h, dt); return 0; }
int pdamr_ard_iterate(pdamr_ctx* c, int n, double dt) { sym<void (*)(void*, int, double)>("ref_ard_iterate")(c->h, n, dt); return 0; }
int pdamr_phase_change(pdamr_ctx* c, int* n_dissolved) {
    const int n = sym<int (*)(void*)>("ref_ard_phase_change")(c->h);
    if (n > 0) { REF0(ref_update_node_types)(c->h); REF0(ref_build_neighbors_celllist)(c->h); }   // src/coupling.cpp:262-268
    if (n_dissolved) *n_dissolved = n;
    return 0;
}
int pdamr_implicit_assemble(pdamr_ctx* c) { REF0(ref_imp_assemble)(c->h); return 0; }
int pdamr_implicit_matvec(pdamr_ctx*, double, const double*, double*) { return 1; }
int pdamr_implicit_rhs(pdamr_ctx*, double, double*) { return 1; }
int pdamr_implicit_compute_dt(pdamr_ctx* c, double frac, double dmax, double* dt) {
    char buf[160];
    std::snprintf(buf, sizeof buf, "implicit_dt_fraction = %.17g\nimplicit_dt_max = %.17g\n", frac, dmax);
    apply(c, buf);
    *dt = sym<double (*)(void*)>("ref_imp_compute_adaptive_dt")(c->h);
    return 0;
}
int pdamr_implicit_step(pdamr_ctx* c, double dt, double, int, int, int, PdLinSolveInfo* info) {
    sym<int (*)(void*, double)>("ref_imp_step")(c->h, dt);
    if (info) { info->iters = 0; info->converged = 1; info->rel_res = 0.0; info->pad = 0; }
    return 0;
}
int pdamr_destroy(pdamr_ctx* c) {
    if (!c) return 0;
    REF0(ref_destroy)(c->h);
    std::remove(c->cfg_path.c_str());
    delete c;
    return 0;
}

// ---- lattice context (2D, one rank): what host/main.cpp + host/coupling.cpp call ---------------------------------
static const char* kFieldName[] = {"rho", "vel", "pressure", "C", "rho_new", "vel_new", "C_new", "phase", "is_gb", "is_precip",
                                   "node_type"};
static long long lat_N(pdgpu_ctx* c) { long long d[5]; dims((pdamr_ctx*)c, d); return d[3]; }

int pdgpu_grid_extents(const PdConfig* k, int dim, int* Nx, int* Ny, int* Nz, double origin[3]) {
    if (dim != 2) { g_err = "fake libpdgpu serves 2D only"; return 1; }
    const std::string p = write_cfg(k, "use_amr = 0\n");
    void* h = sym<void* (*)(const char*)>("ref_create")(p.c_str());
    REF0(ref_grid_build)(h);
    long long d[5];
    sym<void (*)(void*, long long*)>("ref_get_dims")(h, d);
    *Nx = (int)d[0]; *Ny = (int)d[1]; *Nz = (int)d[2];
    sym<void (*)(void*, double*)>("ref_get_origin")(h, origin);
    REF0(ref_destroy)(h);
    std::remove(p.c_str());
    return 0;
}
int pdgpu_create_slab(const PdConfig* k, int dim, int, int, int nranks, pdgpu_ctx** out) {
    if (dim != 2 || nranks != 1) { g_err = "fake libpdgpu serves 2D single-rank runs only"; return 1; }
    pdamr_ctx* c = new pdamr_ctx();
    c->cfg_path = write_cfg(k, "use_amr = 0\n");
    c->h = sym<void* (*)(const char*)>("ref_create")(c->cfg_path.c_str());
    c->cfg = *k;
    // members that do not cross the C ABI (grain parameters): from the run's own configuration file, so that the
    // reference's grains can be compared with the driver's in pdgpu_fields_init
    if (const char* p = std::getenv("PD_FAKE_CFG")) sym<void (*)(void*, const char*)>("ref_config_apply")(c->h, p);
    *out = (pdgpu_ctx*)c;
    return 0;
}
int pdgpu_destroy(pdgpu_ctx* c) { return pdamr_destroy((pdamr_ctx*)c); }
int pdgpu_grid_build(pdgpu_ctx* c) { REF0(ref_grid_build)(((pdamr_ctx*)c)->h); REF0(ref_build_neighbors)(((pdamr_ctx*)c)->h); return 0; }
int pdgpu_grid_info(pdgpu_ctx* c, PdGridInfo* o) {
    std::memset(o, 0, sizeof(*o));
    long long d[5];
    dims((pdamr_ctx*)c, d);
    o->dim = 2; o->Nx = (int)d[0]; o->Ny = (int)d[1]; o->Nz = (int)d[2]; o->N_total = d[3]; o->nnz = d[4];
    o->m = ((pdamr_ctx*)c)->cfg.m_ratio; o->n_off = 36; o->reach = o->m + 1; o->a0 = 0; o->a1 = o->Ny; o->plane = o->Nx;
    const uint8_t* t = (const uint8_t*)ptr((pdamr_ctx*)c, "node_type");
    for (long long i = 0; i < d[3]; ++i) o->counts[t[i]]++;
    sym<void (*)(void*, double*)>("ref_get_origin")(((pdamr_ctx*)c)->h, o->origin);
    return 0;
}
int pdgpu_fields_download_all(pdgpu_ctx* c, int field, void* out) {
    if (field < 0 || field > 10) return 1;
    const size_t b = field_bytes((pdamr_ctx*)c, kFieldName[field]);
    std::memcpy(out, ptr((pdamr_ctx*)c, kFieldName[field]), b);
    return 0;
}
int pdgpu_fields_init(pdgpu_ctx* c, const uint8_t* gb, const uint8_t* pr) {
    pdamr_ctx* a = (pdamr_ctx*)c;
    REF0(ref_fields_init)(a->h);                 // the reference's own grains and initialize_fields ...
    const size_t N = (size_t)lat_N(c);
    if (std::memcmp(ptr(a, "is_gb"), gb, N) != 0 || std::memcmp(ptr(a, "is_precip"), pr, N) != 0) {
        g_err = "fake pdgpu_fields_init: the driver's grain flags differ from the reference's";   // ... must agree with the driver's
        return 1;
    }
    REF0(ref_ns_init)(a->h); REF0(ref_ard_init)(a->h); REF0(ref_imp_init)(a->h);
    return 0;
}
int pdgpu_gather(pdgpu_ctx* c, int field, const int* idx, long long n, double* out) {
    if (field != 3 && field != 0) return 1;
    const double* a = (const double*)ptr((pdamr_ctx*)c, kFieldName[field]);
    for (long long q = 0; q < n; ++q) out[q] = a[idx[q]];
    return 0;
}
int pdgpu_diag(pdgpu_ctx* c, PdDiag* o) {        // reductions of write_diagnostics (src/coupling.cpp:20-49)
    pdamr_ctx* a = (pdamr_ctx*)c;
    const long long N = lat_N(c);
    const uint8_t* t = (const uint8_t*)ptr(a, "node_type");
    const double *v = (const double*)ptr(a, "vel"), *C = (const double*)ptr(a, "C");
    o->solid_count = 0; o->v_max = 0.0; o->C_max_fluid = 0.0;
    for (long long i = 0; i < N; ++i) {
        if (t[i] == 1) o->solid_count++;
        if (t[i] != 0) continue;
        const double m = std::sqrt(v[2 * i] * v[2 * i] + v[2 * i + 1] * v[2 * i + 1]);
        if (m > o->v_max) o->v_max = m;
        if (C[i] > o->C_max_fluid) o->C_max_fluid = C[i];
    }
    return 0;
}
int pdgpu_vti_write(pdgpu_ctx* c, const char* path, const int* grain_id, const double* D_map, long long* bytes_out, float* ms) {
    pdamr_ctx* a = (pdamr_ctx*)c;                // the reference's writer on its state with the DRIVER's host-side arrays
    const size_t N = (size_t)lat_N(c);
    if (grain_id) std::memcpy(ptr(a, "grain_id"), grain_id, 4 * N);
    if (D_map) std::memcpy(ptr(a, "D_map"), D_map, 8 * N);
    sym<double (*)(void*, const char*)>("ref_write_vti")(a->h, path);
    if (bytes_out) *bytes_out = 0;
    if (ms) *ms = 0.0f;
    return 0;
}
int pdgpu_ns_solve_steady(pdgpu_ctx* c, PdSteadyResult* out, int v) { return pdamr_ns_solve_steady((pdamr_ctx*)c, out, v); }
int pdgpu_ard_set_volume_loss(pdgpu_ctx* c, double v) { return pdamr_ard_set_volume_loss((pdamr_ctx*)c, v); }
int pdgpu_ard_compute_dt(pdgpu_ctx* c, double* dt) { return pdamr_ard_compute_dt((pdamr_ctx*)c, dt); }
int pdgpu_ard_iterate(pdgpu_ctx* c, int n, double dt) { return pdamr_ard_iterate((pdamr_ctx*)c, n, dt); }
int pdgpu_implicit_assemble(pdgpu_ctx* c) { return pdamr_implicit_assemble((pdamr_ctx*)c); }
int pdgpu_implicit_compute_dt(pdgpu_ctx* c, double f, double m, double* dt) { return pdamr_implicit_compute_dt((pdamr_ctx*)c, f, m, dt); }
int pdgpu_implicit_step(pdgpu_ctx* c, double dt, double tol, int r, int mi, int, PdLinSolveInfo* info) {
    return pdamr_implicit_step((pdamr_ctx*)c, dt, tol, r, mi, 1, info);
}
int pdgpu_bc_inlet(pdgpu_ctx* c) { return pdamr_bc((pdamr_ctx*)c, 0); }
int pdgpu_bc_outlet(pdgpu_ctx* c) { return pdamr_bc((pdamr_ctx*)c, 1); }
int pdgpu_bc_wall_conc(pdgpu_ctx* c) { return pdamr_bc((pdamr_ctx*)c, 4); }
int pdgpu_bc_smooth_conc(pdgpu_ctx* c) { return pdamr_bc((pdamr_ctx*)c, 6); }
int pdgpu_solid_below_thresh(pdgpu_ctx* c, int* count) {
    pdamr_ctx* a = (pdamr_ctx*)c;
    const long long N = lat_N(c);
    const uint8_t* t = (const uint8_t*)ptr(a, "node_type");
    const double* C = (const double*)ptr(a, "C");
    int n = 0;
    for (long long i = 0; i < N; ++i) n += (t[i] == 1 && C[i] < a->cfg.C_thresh);
    *count = n;
    return 0;
}
int pdgpu_phase_change(pdgpu_ctx* c, int* n_dissolved, int* dissolved, int cap) {
    pdamr_ctx* a = (pdamr_ctx*)c;
    const long long N = lat_N(c);
    const uint8_t* t = (const uint8_t*)ptr(a, "node_type");
    std::vector<uint8_t> before(t, t + N);
    const int n = sym<int (*)(void*)>("ref_ard_phase_change")(a->h);
    if (n > 0) { REF0(ref_update_node_types)(a->h); REF0(ref_build_neighbors)(a->h); }            // src/coupling.cpp:262-270
    t = (const uint8_t*)ptr(a, "node_type");
    int k = 0;
    for (long long i = 0; i < N && dissolved && k < cap; ++i)
        if (before[i] == 1 && t[i] == 0) dissolved[k++] = (int)i;
    if (n_dissolved) *n_dissolved = n;
    return 0;
}
int pdgpu_comm_allreduce(pdgpu_ctx*, double*, int, int) { return 0; }      // one rank

}  // extern "C"
