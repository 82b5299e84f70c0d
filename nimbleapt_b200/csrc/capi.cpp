// capi.cpp — the public C ABI (include/aptb200.h) on top of the front end and the lowered-model modules.
#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <chrono>
#include <string>
#include <vector>
#include "engine/apt_module_abi.h"
#include "front.h"

using namespace aptf;

static thread_local std::string g_err;

struct apt_model {
    ModelSpec spec;
};

struct ModFns {
    int (*abi_version)(void);
    int (*create)(const aptmod_init*, void**, char*, int);
    void (*destroy)(void*);
    int (*run)(void*, const aptmod_run_args*, char*, int);
    int (*get)(void*, int, int, double*, int64_t, int64_t*, char*, int);
    int (*get_cm)(void*, int, int, double*, int64_t, int64_t*, char*, int);
    long long (*info)(void*, int);
    int (*logprob)(void*, const double*, int, double*, char*, int);
    int (*set_model_theta)(void*, const double*, int, char*, int);
    int (*set_array)(void*, int, const double*, int64_t, char*, int);
    void (*timing_get)(void*, aptmod_timing*);
    void* (*stream)(void*);
    int (*device_ptr)(void*, int, void**, int64_t*, void**);
    int (*waic)(void*, long long, int, double*, char*, int);
};

struct apt_engine {
    void* dl = nullptr;
    void* mod = nullptr;
    ModFns fn{};
    Lowered low;
    int n_ladders = 1, device = 0, enable_waic = 0, thin_conf = 1;
    // NCCL (lazy)
    void* nccl_dl = nullptr;
    void* comm = nullptr;
    int world = 1, rank = 0;
    void* gather_buf = nullptr;  // device destination of the trace gathers, kept between calls (grow only)
    size_t gather_cap = 0;
    void* scratch = nullptr;     // small persistent device buffer: pooled statistics (all-reduce), ladder counts
    size_t scratch_cap = 0;
    std::vector<int64_t> peer_ladders;  // ladders held by every rank (ranks may hold different numbers)
    void* copy_stream = nullptr;        // device-to-host copies of gathered pieces, overlapping the next piece's sampling
    std::vector<void*> events;          // one per piece of a gathering run
    apt_timing last{};
};

#define API_BEGIN try {
#define API_END                          \
    return 0;                            \
    }                                    \
    catch (const std::exception& e) {    \
        g_err = e.what();                \
        return 1;                        \
    }                                    \
    catch (...) {                        \
        g_err = "unknown error";         \
        return 1;                        \
    }

extern "C" {

const char* apt_last_error(void) { return g_err.c_str(); }
const char* apt_version(void) { return "aptb200 0.1 (sm_100a)"; }

int64_t apt_abi_sizeof(int which) {
    switch (which) {
        case 0: return (int64_t)sizeof(apt_sampler_control);
        case 1: return (int64_t)sizeof(apt_build_opts);
        case 2: return (int64_t)sizeof(apt_run_args);
        case 3: return (int64_t)sizeof(apt_timing);
        default: return -1;
    }
}
void apt_sampler_control_default(apt_sampler_control* c) {
    c->log = 0; c->reflective = 0; c->adaptive = 1; c->adaptScaleOnly = 0; c->adaptInterval = 200;
    c->adaptFactorExponent = 0.8; c->scale = 1.0; c->temperPriors = -1; c->propCov = nullptr; c->propCov_dim = 0;
    c->width = 1.0; c->sliceMaxSteps = 100; c->useTempering = -1;
}
void apt_build_opts_default(apt_build_opts* o) {
    memset(o, 0, sizeof *o);
    o->monitorTmax = 1; o->ULT = 1e6; o->n_ladders = 1; o->ladder0 = 0; o->seed = 11; o->device = 0;
    o->record_decisions = 0; o->fuse_links = 1;
}
void apt_run_args_default(apt_run_args* a) {
    memset(a, 0, sizeof *a);
    a->reset = 1; a->adaptTemps = 1; a->tuneTemper1 = 10; a->tuneTemper2 = 1; a->progressBar = 1; a->thin = -1; a->thin2 = -1;
    a->gather = 0; a->gather_root = 0;
}

int apt_model_create(const char* bugs_code, apt_model** out) {
    API_BEGIN
    if (!bugs_code || !out) throw Error("null argument");
    auto* m = new apt_model();
    try {
        m->spec.code = bugs_code;
        m->spec.decls = parse_bugs(bugs_code);
        if (m->spec.decls.empty()) throw Error("the model has no declarations");
    } catch (...) {
        delete m;
        throw;
    }
    *out = m;
    API_END
}
void apt_model_free(apt_model* m) { delete m; }

int apt_model_set_array(apt_model* m, const char* name, int role, const double* values, const int64_t* dims, int ndim) {
    API_BEGIN
    if (!m || !name || (!values && ndim >= 0)) throw Error("null argument");
    Array a;
    a.name = name; a.role = role;
    int64_t n = 1;
    for (int i = 0; i < ndim; i++) {
        if (dims[i] < 1) throw Error("non-positive dimension");
        a.dims.push_back(dims[i]);
        n *= dims[i];
    }
    if (ndim == 1 && n == 1) a.dims.clear();  // length-1 vector == scalar
    a.v.assign(values, values + n);
    auto& dst = role == APT_ROLE_CONSTANT ? m->spec.constants : role == APT_ROLE_DATA ? m->spec.data : m->spec.inits;
    if (role < 0 || role > 2) throw Error("unknown role");
    dst[name] = std::move(a);
    API_END
}

int apt_model_add_sampler(apt_model* m, const char* target, const char* type, const apt_sampler_control* control) {
    API_BEGIN
    if (!m || !target || !type) throw Error("null argument");
    SamplerConf sc;
    sc.target = target; sc.type = type;
    if (control) sc.ctl = *control; else apt_sampler_control_default(&sc.ctl);
    if (sc.ctl.propCov) {
        if (sc.ctl.propCov_dim < 1) throw Error("propCov must be a matrix\n");  // APT:814
        sc.propCov.assign(sc.ctl.propCov, sc.ctl.propCov + (size_t)sc.ctl.propCov_dim * sc.ctl.propCov_dim);
        for (double v : sc.propCov)
            if (!std::isfinite(v)) throw Error("propCov matrix must be numeric\n");  // APT:815
    }
    sc.ctl.propCov = nullptr;
    m->spec.samplers.push_back(std::move(sc));
    API_END
}
int apt_model_add_function(apt_model* m, const char* name, const char* r_source) {
    API_BEGIN
    if (!m || !name || !r_source) throw Error("null argument");
    (void)parse_user_function(name, r_source);  // syntax errors are reported here, not at buildAPT time
    m->spec.functions[name] = r_source;
    API_END
}
int apt_model_remove_samplers(apt_model* m) {
    API_BEGIN
    m->spec.samplers.clear();
    API_END
}
int apt_model_set_monitors(apt_model* m, int which, const char* const* names, int n) {
    API_BEGIN
    auto& dst = which == 2 ? m->spec.monitors2 : m->spec.monitors;
    dst.clear();
    for (int i = 0; i < n; i++) dst.push_back(names[i]);
    API_END
}
int apt_model_set_thin(apt_model* m, int thin, int thin2) {
    API_BEGIN
    if (thin < 1 || thin2 < 1) throw Error("thin must be >= 1");
    m->spec.thin = thin; m->spec.thin2 = thin2;
    API_END
}

static ModelSpec with_default_monitors(const ModelSpec& s) {
    ModelSpec spec = s;
    if (spec.monitors.empty()) {  // configureMCMC default: monitor the top-level stochastic nodes; here: all sampled nodes
        std::vector<std::string> seen;
        for (auto& d : spec.decls)
            if (d.stoch && !spec.data.count(d.lhs->name)) {
                bool dup = false;
                for (auto& x : seen) dup |= (x == d.lhs->name);
                if (!dup) seen.push_back(d.lhs->name);
            }
        spec.monitors = seen;
    }
    return spec;
}

static void copy_out(const std::string& s, char* buf, int64_t buflen, int64_t* needed) {
    if (needed) *needed = (int64_t)s.size() + 1;
    if (buf && buflen > 0) {
        size_t n = std::min<size_t>(s.size(), (size_t)buflen - 1);
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
}

int apt_model_names(apt_model* m, int which, char* buf, int64_t buflen) {
    API_BEGIN
    apt_build_opts o;
    apt_build_opts_default(&o);
    Lowered L = lower_model(with_default_monitors(m->spec), 2, o);
    const auto& v = which == 0 ? L.param_names : which == 2 ? L.mon2_names : L.mon_names;
    std::string s;
    for (size_t i = 0; i < v.size(); i++) s += (i ? "\n" : "") + v[i];
    if ((int64_t)s.size() + 1 > buflen) throw Error("buffer too small");
    copy_out(s, buf, buflen, nullptr);
    API_END
}

int apt_model_lower(apt_model* m, int n_temps, const apt_build_opts* opts, char* buf, int64_t buflen, int64_t* needed) {
    API_BEGIN
    apt_build_opts o;
    if (opts) o = *opts; else apt_build_opts_default(&o);
    Lowered L = lower_model(with_default_monitors(m->spec), n_temps, o);
    copy_out(L.source, buf, buflen, needed);
    API_END
}

int apt_model_compile(apt_model* m, int n_temps, const apt_build_opts* opts, char* path_buf, int64_t buflen) {
    API_BEGIN
    apt_build_opts o;
    if (opts) o = *opts; else apt_build_opts_default(&o);
    const ModelSpec spec = with_default_monitors(m->spec);
    Lowered L = lower_model(spec, n_temps, o);
    std::string so = jit_compile(L.source, o.verbose != 0);
    copy_out(so, path_buf, buflen, nullptr);
    API_END
}

int apt_build(apt_model* m, const double* temps, int n_temps, const apt_build_opts* opts, apt_engine** out) {
    API_BEGIN
    if (!m || !temps || !out) throw Error("null argument");
    apt_build_opts o;
    if (opts) o = *opts; else apt_build_opts_default(&o);
    for (int i = 0; i < n_temps; i++)
        if (!(temps[i] > 0) || !std::isfinite(temps[i])) throw Error("temperatures must be positive and finite");
    auto* e = new apt_engine();
    // the lowered description borrows the arrays of this copy until they have been uploaded
    const ModelSpec spec = with_default_monitors(m->spec);
    try {
        e->low = lower_model(spec, n_temps, o);
        if (o.verbose) {  // message(paste("Initial temperatures:", ...))  APT:279
            std::vector<double> t(temps, temps + n_temps);
            std::sort(t.begin(), t.end());
            fprintf(stderr, "Initial temperatures:");
            for (double x : t) fprintf(stderr, " %g", x);
            fprintf(stderr, "\n");
        }
        std::string so = jit_compile(e->low.source, o.verbose != 0);
        // Modules stay loaded for the life of the process: unloading a library that embeds the CUDA runtime and
        // registered kernels is not safe, and a second engine for the same model reuses the loaded code.
        static std::map<std::string, void*> loaded;
        auto it = loaded.find(so);
        if (it != loaded.end()) e->dl = it->second;
        else {
            e->dl = dlopen(so.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (!e->dl) throw Error(std::string("cannot load the compiled module: ") + dlerror());
            loaded[so] = e->dl;
        }
        auto sym = [&](const char* n) {
            void* p = dlsym(e->dl, n);
            if (!p) throw Error(std::string("compiled module lacks symbol ") + n);
            return p;
        };
        e->fn.abi_version = (int (*)(void))sym("aptmod_abi_version");
        if (e->fn.abi_version() != APTMOD_ABI_VERSION) throw Error("stale module in the JIT cache (ABI version mismatch)");
        e->fn.create = (decltype(e->fn.create))sym("aptmod_create");
        e->fn.destroy = (decltype(e->fn.destroy))sym("aptmod_destroy");
        e->fn.run = (decltype(e->fn.run))sym("aptmod_run");
        e->fn.get = (decltype(e->fn.get))sym("aptmod_get");
        e->fn.get_cm = (decltype(e->fn.get_cm))sym("aptmod_get_cm");
        e->fn.info = (decltype(e->fn.info))sym("aptmod_info");
        e->fn.logprob = (decltype(e->fn.logprob))sym("aptmod_logprob");
        e->fn.set_model_theta = (decltype(e->fn.set_model_theta))sym("aptmod_set_model_theta");
        e->fn.set_array = (decltype(e->fn.set_array))sym("aptmod_set_array");
        e->fn.timing_get = (decltype(e->fn.timing_get))sym("aptmod_timing_get");
        e->fn.device_ptr = (decltype(e->fn.device_ptr))dlsym(e->dl, "aptmod_device_ptr");
        e->fn.stream = (decltype(e->fn.stream))sym("aptmod_stream");
        e->fn.waic = (decltype(e->fn.waic))sym("aptmod_waic");
        aptmod_init in{};
        in.n_ladders = o.n_ladders; in.ladder0 = o.ladder0; in.seed = o.seed; in.device = o.device;
        in.temps = temps; in.n_temps = n_temps; in.ULT = o.ULT; in.monitor_tmax = o.monitorTmax;
        std::vector<const double*> ap;
        std::vector<int64_t> al;
        for (auto* a : e->low.arrays) { ap.push_back(a->v.data()); al.push_back(a->size()); }
        std::vector<const int*> ip;
        std::vector<int64_t> il;
        for (auto& v : e->low.iarrays) { ip.push_back(v.data()); il.push_back((int64_t)v.size()); }
        in.n_arrays = (int)ap.size(); in.arrays = ap.data(); in.array_len = al.data();
        in.n_iarrays = (int)ip.size(); in.iarrays = ip.data(); in.iarray_len = il.data();
        in.theta0 = e->low.theta0.data();
        in.record = o.record_decisions;
        in.thin_conf = m->spec.thin; in.thin2_conf = m->spec.thin2;
        char err[2048] = {0};
        if (e->fn.create(&in, &e->mod, err, sizeof err)) throw Error(err);
        e->n_ladders = o.n_ladders; e->device = o.device; e->enable_waic = o.enable_waic;
        e->thin_conf = m->spec.thin > 0 ? m->spec.thin : 1;
        // arrays were copied to the device: drop the borrowed pointers
        e->low.arrays.clear();
    } catch (...) {
        if (e->mod && e->fn.destroy) e->fn.destroy(e->mod);
        delete e;
        throw;
    }
    *out = e;
    API_END
}

int apt_comm_finalize(apt_engine* e);

void apt_free(apt_engine* e) {
    if (!e) return;
    apt_comm_finalize(e);
    if (e->mod) e->fn.destroy(e->mod);
    delete e;
}

struct GatherRun;
static int gather_chunk_cb(void* arg, const double* d_packed, int64_t n, int64_t row0, int64_t row1, void* stream);
static void gather_run_begin(apt_engine* e, const apt_run_args* args, GatherRun& g, aptmod_run_args& ra);
static void gather_run_end(apt_engine* e, GatherRun& g, bool run_ok);
struct GatherRun {
    apt_engine* e = nullptr;
    int root = 0, piece = 0;
    bool active = false, deliver = false;
    double* out = nullptr;
    int64_t nr = 0, cols = 0, sumL = 0;
    std::vector<int64_t> loff;  // first ladder of every rank in the gathered order
    std::string err;
};

int apt_run(apt_engine* e, const apt_run_args* args) {
    API_BEGIN
    if (!e || !args) throw Error("null argument");
    if (args->simulateAll) throw Error("simulateAll = TRUE is not supported by the device engine (set initial values with apt_set_inits)");
    const auto t0 = std::chrono::steady_clock::now();
    aptmod_run_args ra{};
    ra.niter = args->niter; ra.reset = args->reset; ra.resetMV = args->resetMV; ra.resetTempering = args->resetTempering;
    ra.adaptTemps = args->adaptTemps; ra.tuneTemper1 = args->tuneTemper1; ra.tuneTemper2 = args->tuneTemper2;
    ra.thin = args->thin; ra.thin2 = args->thin2;
    ra.tabZ = args->inject_normals; ra.tabU = args->inject_accept_u; ra.tabS = args->inject_swap_u;
    ra.should_abort = args->should_abort; ra.abort_arg = args->abort_arg; ra.iters_per_launch = args->iters_per_launch;
    ra.samples_out = args->samples_out; ra.samples_out_cap = args->samples_out_capacity;
    GatherRun g;
    if (args->gather) gather_run_begin(e, args, g, ra);
    char err[2048] = {0};
    const int rc = e->fn.run(e->mod, &ra, err, sizeof err);
    const auto t1 = std::chrono::steady_clock::now();
    if (g.active) gather_run_end(e, g, rc == 0);
    const auto t2 = std::chrono::steady_clock::now();
    if (rc) throw Error(g.err.empty() ? std::string(err) : g.err);
    if (!g.err.empty()) throw Error(g.err);
    aptmod_timing mt{};
    e->fn.timing_get(e->mod, &mt);
    e->last.kernel_ms = mt.kernel_ms; e->last.main_kernel_ms = mt.main_kernel_ms; e->last.launches = mt.launches; e->last.main_launches = mt.main_launches;
    e->last.run_wall_ms = std::chrono::duration<double, std::milli>(t2 - t0).count();
    e->last.gather_tail_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
    if (args->printTemps) {  // nimPrint(iter, " ", totalIters-1, " ", size_mvSamples, " ", size_mvSamples2, asRow(Temps))  APT:524
        std::vector<double> t((size_t)e->low.K);
        int64_t n = 0;
        if (!e->fn.get(e->mod, APTMOD_TEMPS, 0, t.data(), (int64_t)t.size(), &n, err, sizeof err)) {
            printf("%lld %lld %lld %lld", (long long)args->niter, e->fn.info(e->mod, APTMOD_INFO_TOTALITERS) - 1,
                   e->fn.info(e->mod, APTMOD_INFO_ROWS), e->fn.info(e->mod, APTMOD_INFO_ROWS2));
            for (double x : t) printf(" %g", x);
            printf("\n");
        }
    }
    API_END
}

int64_t apt_info(apt_engine* e, int what) {
    if (!e) return -1;
    return e->fn.info(e->mod, what);
}

int apt_get(apt_engine* e, int what, int ladder, double* out, int64_t capacity, int64_t* n_written) {
    API_BEGIN
    if (!e || !out) throw Error("null argument");
    char err[2048] = {0};
    int64_t n = 0;
    if (e->fn.get(e->mod, what, ladder, out, capacity, &n, err, sizeof err)) throw Error(err);
    if (n_written) *n_written = n;
    API_END
}

int apt_get_layout(apt_engine* e, int what, int ladder, int layout, double* out, int64_t capacity, int64_t* n_written) {
    API_BEGIN
    if (!e || !out) throw Error("null argument");
    if (layout != APT_ROW_MAJOR && layout != APT_COLUMN_MAJOR) throw Error("layout must be APT_ROW_MAJOR or APT_COLUMN_MAJOR");
    char err[2048] = {0};
    int64_t n = 0;
    if (layout == APT_COLUMN_MAJOR) {
        if (ladder < 0) throw Error("column-major output is per ladder (ladder >= 0)");
        if (e->fn.get_cm(e->mod, what, ladder, out, capacity, &n, err, sizeof err)) throw Error(err);
    } else if (e->fn.get(e->mod, what, ladder, out, capacity, &n, err, sizeof err)) throw Error(err);
    if (n_written) *n_written = n;
    API_END
}

int apt_get_timing(apt_engine* e, apt_timing* t) {
    API_BEGIN
    aptmod_timing mt{};
    e->fn.timing_get(e->mod, &mt);
    *t = e->last;  // host-clock figures of the last apt_run
    t->kernel_ms = mt.kernel_ms; t->main_kernel_ms = mt.main_kernel_ms; t->launches = mt.launches; t->main_launches = mt.main_launches;
    API_END
}

int apt_logprob(apt_engine* e, const double* theta, int n, double* out_total, double* out_groups) {
    API_BEGIN
    if (!e || !theta || !out_total) throw Error("null argument");
    const int ND = e->low.ND;
    std::vector<double> g((size_t)n * ND);
    char err[2048] = {0};
    if (e->fn.logprob(e->mod, theta, n, g.data(), err, sizeof err)) throw Error(err);
    for (int i = 0; i < n; i++) {
        double s = 0;
        for (int d = 0; d < ND; d++) s += g[(size_t)i * ND + d];
        out_total[i] = s;
    }
    if (out_groups) memcpy(out_groups, g.data(), g.size() * sizeof(double));
    API_END
}

int apt_waic(apt_engine* e, int64_t nburnin, int ladder, double out[3]) {
    API_BEGIN
    if (!e || !out) throw Error("null argument");
    out[0] = out[1] = out[2] = NAN;
    if (!e->enable_waic)  // APT:596-599 (the reference prints this and returns NaN)
        throw Error("must set enableWAIC = TRUE in buildAPT to use the method calcualteWAIC. See help(buildAPT) and help(buildMCMC) for additional information.");
    char err[2048] = {0};
    if (e->fn.waic(e->mod, (long long)nburnin, ladder, out, err, sizeof err)) throw Error(err);
    API_END
}

int apt_set_inits(apt_engine* e, const double* theta, int ladder) {
    API_BEGIN
    char err[2048] = {0};
    if (e->fn.set_model_theta(e->mod, theta, ladder, err, sizeof err)) throw Error(err);
    API_END
}

int apt_set_array(apt_engine* e, const char* name, const double* values, int64_t n) {
    API_BEGIN
    int idx = -1;
    for (size_t i = 0; i < e->low.array_names.size(); i++)
        if (e->low.array_names[i] == name) idx = (int)i;
    if (idx < 0) throw Error(std::string("array '") + name + "' is not resident on the device (unknown, unused, or folded into the code as a small constant)");
    char err[2048] = {0};
    if (e->fn.set_array(e->mod, idx, values, n, err, sizeof err)) throw Error(err);
    API_END
}

// ------------------------------------------------------------------------------------------------ NCCL
// Loaded lazily so that the library itself has no link-time dependency on NCCL or CUDA (it must load on a
// host without a GPU).  Only gathers of traces and sums of pooled statistics cross NVLink.
typedef struct { char internal[128]; } nccl_uid;
struct NcclFns {
    int (*GetUniqueId)(nccl_uid*);
    int (*CommInitRank)(void**, int, nccl_uid, int);
    int (*CommDestroy)(void*);
    int (*AllGather)(const void*, void*, size_t, int, void*, void*);
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, void*);
    int (*Send)(const void*, size_t, int, int, void*, void*);
    int (*Recv)(void*, size_t, int, int, void*, void*);
    int (*GroupStart)(void);
    int (*GroupEnd)(void);
    const char* (*GetErrorString)(int);
};
static NcclFns g_nccl{};
static void* g_nccl_dl = nullptr;
static void* g_cudart_dl = nullptr;
struct RtFns {
    int (*SetDevice)(int);
    int (*Malloc)(void**, size_t);
    int (*Free)(void*);
    int (*Memcpy)(void*, const void*, size_t, int);
    int (*MemcpyAsync)(void*, const void*, size_t, int, void*);
    int (*Memcpy2DAsync)(void*, size_t, const void*, size_t, size_t, size_t, int, void*);
    int (*DeviceSynchronize)(void);
    int (*StreamSynchronize)(void*);
    int (*StreamCreateWithFlags)(void**, unsigned);
    int (*StreamDestroy)(void*);
    int (*EventCreateWithFlags)(void**, unsigned);
    int (*EventDestroy)(void*);
    int (*EventRecord)(void*, void*);
    int (*StreamWaitEvent)(void*, void*, unsigned);
};
static RtFns g_rt{};

static void load_nccl() {
    if (g_nccl_dl) return;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (auto n : names) {
        g_nccl_dl = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl_dl) break;
    }
    if (!g_nccl_dl) throw Error(std::string("cannot load NCCL: ") + dlerror());
    auto s = [&](const char* n) {
        void* p = dlsym(g_nccl_dl, n);
        if (!p) throw Error(std::string("NCCL lacks ") + n);
        return p;
    };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))s("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))s("ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))s("ncclCommDestroy");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))s("ncclAllGather");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))s("ncclAllReduce");
    g_nccl.Send = (decltype(g_nccl.Send))s("ncclSend");
    g_nccl.Recv = (decltype(g_nccl.Recv))s("ncclRecv");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))s("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))s("ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))s("ncclGetErrorString");
    const char* rts[] = {"libcudart.so.12", "libcudart.so"};
    for (auto n : rts) {
        g_cudart_dl = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_cudart_dl) break;
    }
    if (!g_cudart_dl) throw Error(std::string("cannot load libcudart: ") + dlerror());
    auto r = [&](const char* n) {
        void* p = dlsym(g_cudart_dl, n);
        if (!p) throw Error(std::string("libcudart lacks ") + n);
        return p;
    };
    g_rt.SetDevice = (decltype(g_rt.SetDevice))r("cudaSetDevice");
    g_rt.Malloc = (decltype(g_rt.Malloc))r("cudaMalloc");
    g_rt.Free = (decltype(g_rt.Free))r("cudaFree");
    g_rt.Memcpy = (decltype(g_rt.Memcpy))r("cudaMemcpy");
    g_rt.MemcpyAsync = (decltype(g_rt.MemcpyAsync))r("cudaMemcpyAsync");
    g_rt.DeviceSynchronize = (decltype(g_rt.DeviceSynchronize))r("cudaDeviceSynchronize");
    g_rt.StreamSynchronize = (decltype(g_rt.StreamSynchronize))r("cudaStreamSynchronize");
    g_rt.Memcpy2DAsync = (decltype(g_rt.Memcpy2DAsync))r("cudaMemcpy2DAsync");
    g_rt.StreamCreateWithFlags = (decltype(g_rt.StreamCreateWithFlags))r("cudaStreamCreateWithFlags");
    g_rt.StreamDestroy = (decltype(g_rt.StreamDestroy))r("cudaStreamDestroy");
    g_rt.EventCreateWithFlags = (decltype(g_rt.EventCreateWithFlags))r("cudaEventCreateWithFlags");
    g_rt.EventDestroy = (decltype(g_rt.EventDestroy))r("cudaEventDestroy");
    g_rt.EventRecord = (decltype(g_rt.EventRecord))r("cudaEventRecord");
    g_rt.StreamWaitEvent = (decltype(g_rt.StreamWaitEvent))r("cudaStreamWaitEvent");
}
#define NCCL_OK(x)                                                                             \
    do {                                                                                       \
        int r_ = (x);                                                                          \
        if (r_ != 0) throw Error(std::string("NCCL: ") + g_nccl.GetErrorString(r_) + " in " #x); \
    } while (0)
#define RT_OK(x)                                                          \
    do {                                                                  \
        int r_ = (x);                                                     \
        if (r_ != 0) throw Error("CUDA runtime error " + std::to_string(r_) + " in " #x); \
    } while (0)

int apt_comm_unique_id(unsigned char id_out[128]) {
    API_BEGIN
    load_nccl();
    nccl_uid id;
    NCCL_OK(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, 128);
    API_END
}

static void ensure_dev(void*& buf, size_t& cap, size_t bytes) {
    if (bytes <= cap) return;
    if (buf) g_rt.Free(buf);
    buf = nullptr; cap = 0;
    RT_OK(g_rt.Malloc(&buf, bytes));
    cap = bytes;
}
// rows x columns of 8-byte elements between two pitched layouts; cudaMemcpy2D rejects pitches of 2 GiB and more
static void copy_2d(double* dst, size_t dpitch_d, const double* src, size_t spitch_d, size_t width_d, size_t height, int kind, void* stream) {
    if (!width_d || !height) return;
    const size_t lim = (size_t)1 << 31;
    if (height > 1 && dpitch_d * 8 < lim && spitch_d * 8 < lim)
        RT_OK(g_rt.Memcpy2DAsync(dst, dpitch_d * 8, src, spitch_d * 8, width_d * 8, height, kind, stream));
    else
        for (size_t h = 0; h < height; h++) RT_OK(g_rt.MemcpyAsync(dst + h * dpitch_d, src + h * spitch_d, width_d * 8, kind, stream));
}

int apt_comm_init(apt_engine* e, int world_size, int rank, const unsigned char unique_id[128]) {
    API_BEGIN
    load_nccl();
    if (e->comm) throw Error("communicator already initialised");
    if (world_size < 1 || rank < 0 || rank >= world_size) throw Error("rank is not in [0, world_size)");
    nccl_uid id;
    memcpy(&id, unique_id, 128);
    RT_OK(g_rt.SetDevice(e->device));
    NCCL_OK(g_nccl.CommInitRank(&e->comm, world_size, id, rank));
    e->world = world_size; e->rank = rank;
    // Ranks may hold different numbers of ladders (sharding.py::shard_ladders gives the first ranks one more when the
    // total does not divide): every rank learns every rank's count once, here; the gathers size their pieces from it.
    ensure_dev(e->scratch, e->scratch_cap, std::max<size_t>(64 * 1024, (size_t)world_size * 16));
    void* stream = e->fn.stream(e->mod);
    std::vector<double> cnt((size_t)world_size, 0.0);
    cnt[(size_t)rank] = (double)e->n_ladders;
    double* d = (double*)e->scratch;
    RT_OK(g_rt.MemcpyAsync(d + world_size + rank, &cnt[(size_t)rank], 8, 1 /* H2D */, stream));
    NCCL_OK(g_nccl.AllGather(d + world_size + rank, d, 1, 8 /* ncclFloat64 */, e->comm, stream));
    RT_OK(g_rt.MemcpyAsync(cnt.data(), d, (size_t)world_size * 8, 2 /* D2H */, stream));
    RT_OK(g_rt.StreamSynchronize(stream));
    e->peer_ladders.assign((size_t)world_size, 0);
    for (int r = 0; r < world_size; r++) e->peer_ladders[(size_t)r] = (int64_t)cnt[(size_t)r];
    if (e->peer_ladders[(size_t)rank] != e->n_ladders) throw Error("NCCL all-gather of the ladder counts returned a wrong own count");
    API_END
}

// doubles of a trace held by rank r, given this rank's n_local doubles for its own ladders (same rows x columns everywhere)
static int64_t peer_count(const apt_engine* e, int r, int64_t n_local) {
    return n_local / std::max(1, e->n_ladders) * e->peer_ladders[(size_t)r];
}

// NCCL type/op codes: ncclFloat64 = 8, ncclSum = 0
int apt_comm_allgather_samples(apt_engine* e, int what, double* out, int64_t capacity, int64_t* n_written) {
    API_BEGIN
    if (!e->comm) throw Error("apt_comm_init has not been called");
    RT_OK(g_rt.SetDevice(e->device));
    // staging through the getter (device -> host) is avoided: gather straight from the device-resident trace
    if (!e->fn.device_ptr) throw Error("module lacks aptmod_device_ptr");
    void* dptr = nullptr; int64_t n_local = 0; void* stream = nullptr;
    if (e->fn.device_ptr(e->mod, what, &dptr, &n_local, &stream)) throw Error("trace not available on the device");
    int64_t total = 0;
    bool equal = true;
    std::vector<int64_t> off((size_t)e->world);
    for (int r = 0; r < e->world; r++) {
        off[(size_t)r] = total;
        total += peer_count(e, r, n_local);
        equal &= e->peer_ladders[(size_t)r] == e->n_ladders;
    }
    if (total > capacity) throw Error("output buffer too small");
    // the gathered traces land in a device buffer the engine keeps (a cudaMalloc / cudaFree pair per call costs
    // milliseconds and makes NCCL register a new buffer every time); the copy to the caller's host buffer is queued
    // on the same stream right behind the collective
    ensure_dev(e->gather_buf, e->gather_cap, (size_t)std::max<int64_t>(total, 1) * 8);
    if (equal) {
        NCCL_OK(g_nccl.AllGather(dptr, e->gather_buf, (size_t)n_local, 8 /* ncclFloat64 */, e->comm, stream));
    } else {  // different ladder counts: every rank sends its share to every other one (grouped point-to-point)
        RT_OK(g_rt.MemcpyAsync((double*)e->gather_buf + off[(size_t)e->rank], dptr, (size_t)n_local * 8, 3 /* D2D */, stream));
        NCCL_OK(g_nccl.GroupStart());
        for (int r = 0; r < e->world; r++)
            if (r != e->rank) {
                NCCL_OK(g_nccl.Send(dptr, (size_t)n_local, 8, r, e->comm, stream));
                NCCL_OK(g_nccl.Recv((double*)e->gather_buf + off[(size_t)r], (size_t)peer_count(e, r, n_local), 8, r, e->comm, stream));
            }
        NCCL_OK(g_nccl.GroupEnd());
    }
    RT_OK(g_rt.MemcpyAsync(out, e->gather_buf, (size_t)total * 8, 2 /* D2H */, stream));
    RT_OK(g_rt.StreamSynchronize(stream));
    if (n_written) *n_written = total;
    API_END
}

// Gather to ONE rank: the host process that hands results to the user (the reference's user sits in one R session)
// is the only one that needs every rank's traces in host memory.  With the all-gather above every rank copies
// world x its own share over PCIe into the same host memory at the same time, which is what limited the end-to-end
// figure at 8 GPUs; here the other ranks only send their device-resident trace over NVLink (grouped ncclSend/ncclRecv)
// and the device-to-host copy happens once.  This entry point gathers a WHOLE trace after the run; a run that should
// hand its rows over while it is still sampling uses apt_run_args.gather instead (below).
int apt_comm_gather_samples(apt_engine* e, int what, int root, double* out, int64_t capacity, int64_t* n_written) {
    API_BEGIN
    if (!e->comm) throw Error("apt_comm_init has not been called");
    if (root < 0 || root >= e->world) throw Error("root is not a rank of the communicator");
    RT_OK(g_rt.SetDevice(e->device));
    if (!e->fn.device_ptr) throw Error("module lacks aptmod_device_ptr");
    void* dptr = nullptr; int64_t n_local = 0; void* stream = nullptr;
    if (e->fn.device_ptr(e->mod, what, &dptr, &n_local, &stream)) throw Error("trace not available on the device");
    if (n_written) *n_written = 0;
    if (e->rank != root) {
        NCCL_OK(g_nccl.Send(dptr, (size_t)n_local, 8 /* ncclFloat64 */, root, e->comm, stream));
        RT_OK(g_rt.StreamSynchronize(stream));
        return 0;
    }
    int64_t total = 0;
    std::vector<int64_t> off((size_t)e->world);
    for (int r = 0; r < e->world; r++) { off[(size_t)r] = total; total += peer_count(e, r, n_local); }
    // a root that cannot deliver still receives what its peers are sending (they would wait for ever otherwise)
    const bool deliver = out && total <= capacity;
    ensure_dev(e->gather_buf, e->gather_cap, (size_t)std::max<int64_t>(total, 1) * 8);
    // the root's own share goes straight to the caller's buffer while the peers' shares arrive over NVLink
    if (deliver) RT_OK(g_rt.MemcpyAsync(out + off[(size_t)root], dptr, (size_t)n_local * 8, 2 /* D2H */, stream));
    NCCL_OK(g_nccl.GroupStart());
    for (int r = 0; r < e->world; r++)
        if (r != root)
            NCCL_OK(g_nccl.Recv((double*)e->gather_buf + off[(size_t)r], (size_t)peer_count(e, r, n_local), 8, r, e->comm, stream));
    NCCL_OK(g_nccl.GroupEnd());
    if (deliver && root > 0) RT_OK(g_rt.MemcpyAsync(out, e->gather_buf, (size_t)off[(size_t)root] * 8, 2, stream));
    if (deliver && root + 1 < e->world)
        RT_OK(g_rt.MemcpyAsync(out + off[(size_t)root + 1], (double*)e->gather_buf + off[(size_t)root + 1],
                               (size_t)(total - off[(size_t)root + 1]) * 8, 2, stream));
    RT_OK(g_rt.StreamSynchronize(stream));
    if (!deliver) throw Error("output buffer too small");
    if (n_written) *n_written = total;
    API_END
}

// ---- apt_run_args.gather: the rows a run writes travel to the root WHILE the run goes on.  After every piece of the run
// (the module splits it into 4-8 launches) the module packs the piece's rows on its stream and calls gather_chunk_cb:
// a non-root rank enqueues one ncclSend of the packed piece, the root one group of ncclRecv into its gather buffer, both
// on the engine's stream BETWEEN two sampling launches -- an NCCL kernel running beside the sampling kernel would take
// SM slots the ladder-resident kernel needs to keep every ladder resident.  The root then copies the piece to the caller's
// host buffer on a second stream (copy engine), which does overlap the next piece's sampling.  What is left when the
// last launch ends is one piece's send + copy.
static void gather_run_begin(apt_engine* e, const apt_run_args* args, GatherRun& g, aptmod_run_args& ra) {
    if (!e->comm) throw Error("apt_run_args.gather needs a communicator: call apt_comm_init first");
    if (args->gather_root < 0 || args->gather_root >= e->world) throw Error("gather_root is not a rank of the communicator");
    RT_OK(g_rt.SetDevice(e->device));
    const int thin = args->thin != -1 ? args->thin : e->thin_conf;
    if (thin < 1) throw Error("thin and thin2 must be >= 1");
    g.e = e; g.root = args->gather_root; g.active = true; g.piece = 0;
    g.nr = args->niter / thin;
    g.cols = e->fn.info(e->mod, APTMOD_INFO_NMON);
    g.loff.assign((size_t)e->world + 1, 0);
    for (int r = 0; r < e->world; r++) g.loff[(size_t)r + 1] = g.loff[(size_t)r] + e->peer_ladders[(size_t)r];
    g.sumL = g.loff[(size_t)e->world];
    g.out = args->samples_out;
    if (e->rank == g.root) {
        g.deliver = g.out && g.sumL * g.nr * g.cols <= args->samples_out_capacity;
        if (!g.deliver) g.err = "samples_out: buffer too small for (ladders of all ranks) x (niter / thin) x n_monitors doubles";
        ensure_dev(e->gather_buf, e->gather_cap, (size_t)std::max<int64_t>(g.sumL * g.nr * g.cols, 1) * 8);
        if (!e->copy_stream) RT_OK(g_rt.StreamCreateWithFlags(&e->copy_stream, 1 /* cudaStreamNonBlocking */));
    }
    ra.samples_out = nullptr; ra.samples_out_cap = 0;
    ra.chunk_cb = gather_chunk_cb; ra.chunk_arg = &g;
}

static int gather_chunk_cb(void* arg, const double* d_packed, int64_t n, int64_t row0, int64_t row1, void* stream) {
    GatherRun& g = *static_cast<GatherRun*>(arg);
    apt_engine* e = g.e;
    try {
        const int64_t w = (row1 - row0) * g.cols;  // doubles per ladder in this piece
        if (n != w * e->n_ladders) throw Error("gather: unexpected piece size");
        if (e->rank != g.root) {
            NCCL_OK(g_nccl.Send(d_packed, (size_t)n, 8 /* ncclFloat64 */, g.root, e->comm, stream));
        } else {
            // piece layout in the gather buffer: [all ladders, rank-major][row1 - row0][cols], pieces one after the other
            double* base = (double*)e->gather_buf + (size_t)g.sumL * row0 * g.cols;
            RT_OK(g_rt.MemcpyAsync(base + (size_t)g.loff[(size_t)g.root] * w, d_packed, (size_t)n * 8, 3 /* D2D */, stream));
            if (e->world > 1) {
                NCCL_OK(g_nccl.GroupStart());
                for (int r = 0; r < e->world; r++)
                    if (r != g.root)
                        NCCL_OK(g_nccl.Recv(base + (size_t)g.loff[(size_t)r] * w, (size_t)(e->peer_ladders[(size_t)r] * w), 8, r, e->comm, stream));
                NCCL_OK(g_nccl.GroupEnd());
            }
            if (g.deliver) {
                while ((int)e->events.size() <= g.piece) {
                    void* ev = nullptr;
                    RT_OK(g_rt.EventCreateWithFlags(&ev, 2 /* cudaEventDisableTiming */));
                    e->events.push_back(ev);
                }
                RT_OK(g_rt.EventRecord(e->events[(size_t)g.piece], stream));
                RT_OK(g_rt.StreamWaitEvent(e->copy_stream, e->events[(size_t)g.piece], 0));
                copy_2d(g.out + (size_t)row0 * g.cols, (size_t)g.nr * g.cols, base, (size_t)w, (size_t)w, (size_t)g.sumL, 2 /* D2H */, e->copy_stream);
            }
        }
        g.piece++;
        return 0;
    } catch (const std::exception& x) {
        g.err = x.what();
        return 1;
    }
}

static void gather_run_end(apt_engine* e, GatherRun& g, bool run_ok) {
    (void)run_ok;
    if (e->rank == g.root && e->copy_stream) g_rt.StreamSynchronize(e->copy_stream);
}

int apt_comm_allreduce_sum(apt_engine* e, double* inout, int64_t n) {
    API_BEGIN
    if (!e->comm) throw Error("apt_comm_init has not been called");
    if (n < 0 || (n > 0 && !inout)) throw Error("null argument");
    RT_OK(g_rt.SetDevice(e->device));
    // persistent device scratch and the engine's own stream: no allocation, no device-wide synchronisation per call
    ensure_dev(e->scratch, e->scratch_cap, std::max<size_t>(64 * 1024, (size_t)n * 8));
    void* stream = e->fn.stream(e->mod);
    RT_OK(g_rt.MemcpyAsync(e->scratch, inout, (size_t)n * 8, 1 /* H2D */, stream));
    NCCL_OK(g_nccl.AllReduce(e->scratch, e->scratch, (size_t)n, 8 /* ncclFloat64 */, 0 /* ncclSum */, e->comm, stream));
    RT_OK(g_rt.MemcpyAsync(inout, e->scratch, (size_t)n * 8, 2 /* D2H */, stream));
    RT_OK(g_rt.StreamSynchronize(stream));
    API_END
}

int apt_comm_finalize(apt_engine* e) {
    API_BEGIN
    if (e && (e->gather_buf || e->scratch || e->copy_stream || !e->events.empty())) {
        g_rt.SetDevice(e->device);
        if (e->gather_buf) g_rt.Free(e->gather_buf);
        if (e->scratch) g_rt.Free(e->scratch);
        for (void* ev : e->events) g_rt.EventDestroy(ev);
        if (e->copy_stream) g_rt.StreamDestroy(e->copy_stream);
        e->gather_buf = nullptr; e->gather_cap = 0; e->scratch = nullptr; e->scratch_cap = 0; e->copy_stream = nullptr;
        e->events.clear();
    }
    if (e && e->comm) {
        g_nccl.CommDestroy(e->comm);
        e->comm = nullptr;
    }
    API_END
}

}  // extern "C"
