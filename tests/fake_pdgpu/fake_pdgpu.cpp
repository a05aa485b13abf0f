// tests/fake_pdgpu/fake_pdgpu.cpp -- TEST INFRASTRUCTURE ONLY (built and used by tests/test_cpp_driver_host_logic.py).
//
// A stand-in `libpdgpu.so` whose pdamr_* entry points are served by the compiled reference
// (oracle/_ref/libpdrefimp2d.so, through the C entry points of oracle/ref_shim.cpp).  Put in front of the real library
// with LD_LIBRARY_PATH, it lets the HOST logic of the C++ driver host/pd_corrosion_gpu (host/amr_run.cpp: cycle
// structure, batching between output points, snapshot cadence, PVD / CSV / VTU writing, D_map bookkeeping) run on a
// machine without a GPU; its output files must then equal those of the reference's own main() byte for byte.
// The operators themselves are what the -m gpu tests check on the device.  The lattice entry points (pdgpu_*) are
// present only so that the driver binary loads (it is linked BIND_NOW): fake_stubs.cpp, they fail when called.
#include <dlfcn.h>

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

struct pdamr_ctx {
    void* h = nullptr;
    std::string cfg_path;
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

int pdamr_create(const PdConfig* k, int amr_ratio, double amr_buffer, pdamr_ctx** out) {
    pdamr_ctx* c = new pdamr_ctx();
    char tmpl[] = "/tmp/pdfake_XXXXXX";
    const int fd = mkstemp(tmpl);
    if (fd < 0) return 1;
    c->cfg_path = tmpl;
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
    std::fprintf(f, "use_amr = 1\namr_ratio = %d\namr_buffer = %.17g\n", amr_ratio, amr_buffer);
    std::fclose(f);
    sym<void (*)(int)>("ref_set_threads")(1);                       // the reference's in-place loops are order dependent
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

}  // extern "C"
