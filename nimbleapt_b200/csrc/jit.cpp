// jit.cpp — "compiled with nvcc at buildAPT time": turns the lowered CUDA source into a loadable module.
// The reference compiles its generated C++ with g++ through nimble at the same point (compileNimble,
// /root/reference/nimbleAPT/R/APT_functions.R:157-162, vignette:137).  Modules are cached by content hash
// next to the library (nimbleapt_b200/_jit_cache), so a model that was built once (for instance on a machine
// without a GPU: nvcc cross-compiles for sm_100a) is not compiled again.
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include "front.h"

namespace aptf {

static std::string lib_dir() {
    Dl_info info;
    if (dladdr((void*)&lib_dir, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t s = p.find_last_of('/');
        if (s != std::string::npos) return p.substr(0, s);
    }
    return ".";
}

std::string engine_include_dir() {
    if (const char* e = getenv("APTB200_ENGINE_INCLUDE")) return e;
    return lib_dir() + "/csrc/engine";
}

static std::string cache_dir() {
    if (const char* e = getenv("APTB200_CACHE")) return e;
    return lib_dir() + "/_jit_cache";
}

static uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ull) {
    for (unsigned char c : s) {
        h ^= c;
        h *= 1099511628211ull;
    }
    return h;
}

static std::string slurp(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    std::ostringstream o;
    o << f.rdbuf();
    return o.str();
}

// single-quoted for the shell: paths with spaces or metacharacters neither break nor inject
static std::string shq(const std::string& s) {
    std::string o = "'";
    for (char c : s) {
        if (c == '\'') o += "'\\''";
        else o += c;
    }
    return o + "'";
}

// `nvcc --version`, recorded in the module's build log (the cache key stays the hash of source + headers + flags, so that a
// module cross-compiled on a machine without a GPU and shipped is found); asked once per process
static const std::string& nvcc_version(const std::string& nvcc) {
    static std::string v;
    static bool done = false;
    if (!done) {
        done = true;
        if (FILE* p = popen((shq(nvcc) + " --version 2>/dev/null").c_str(), "r")) {
            char buf[256];
            while (fgets(buf, sizeof buf, p)) v += buf;
            pclose(p);
        }
    }
    return v;
}

std::string jit_compile(const std::string& source, bool verbose) {
    const std::string inc = engine_include_dir();
    const std::string nvcc = getenv("APTB200_NVCC") ? getenv("APTB200_NVCC") : "/usr/local/cuda/bin/nvcc";
    const std::string extra = getenv("APTB200_EXTRA_NVCC_FLAGS") ? std::string(" ") + getenv("APTB200_EXTRA_NVCC_FLAGS") : std::string();
    const std::string flags = extra + " -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xcompiler -fno-gnu-unique -diag-suppress 177 -shared";
    uint64_t h = fnv1a(source);
    for (const char* hdr : {"apt_engine.cuh", "apt_host.cuh", "apt_dists.cuh", "apt_module_abi.h", "apt_sferr_table.cuh", "apt_grid.cuh", "apt_softplus.cuh", "apt_softplus_tables.h", "apt_waic.cuh"}) {
        std::string txt = slurp(inc + "/" + hdr);
        h = fnv1a(txt, h);
    }
    h = fnv1a(flags, h);
    char hex[32];
    snprintf(hex, sizeof hex, "%016llx", (unsigned long long)h);
    const std::string dir = cache_dir();
    mkdir(dir.c_str(), 0755);
    const std::string base = dir + "/aptmod_" + hex;
    const std::string so = base + ".so", cu = base + ".cu", log = base + ".log";
    struct stat st;
    if (stat(so.c_str(), &st) == 0 && st.st_size > 0) return so;
    if (access(nvcc.c_str(), X_OK) != 0) throw Error("nvcc not found at " + nvcc + " (set APTB200_NVCC); the model cannot be compiled");
    // Several processes may build the same model at once (8 ranks calling buildAPT): everything is written under names unique
    // to this process and renamed into place, so nobody ever reads a half-written source, log or module.
    const std::string uniq = ".tmp." + std::to_string((long)getpid());
    const std::string cu_tmp = base + uniq + ".cu", log_tmp = base + uniq + ".log", so_tmp = base + uniq + ".so";
    {
        std::ofstream f(cu_tmp);
        if (!f) throw Error("cannot write " + cu_tmp);
        f << source;
    }
    const std::string cmd = shq(nvcc) + " " + flags + " -I" + shq(inc) + " " + shq(cu_tmp) + " -o " + shq(so_tmp) + " > " + shq(log_tmp) + " 2>&1";
    if (verbose) fprintf(stderr, "[aptb200] %s\n", cmd.c_str());
    int rc = system(cmd.c_str());
    rename(cu_tmp.c_str(), cu.c_str());
    {
        std::ofstream f(log_tmp, std::ios::app);
        f << "\n# " << cmd << "\n# " << nvcc_version(nvcc);
    }
    rename(log_tmp.c_str(), log.c_str());
    if (rc != 0) {
        std::string l = slurp(log);
        if (l.size() > 4000) l = l.substr(0, 4000) + "\n...";
        unlink(so_tmp.c_str());
        throw Error("nvcc failed for the lowered model (source kept at " + cu + "):\n" + l);
    }
    if (rename(so_tmp.c_str(), so.c_str()) != 0) throw Error("cannot move the compiled module into the cache");
    return so;
}

}  // namespace aptf
