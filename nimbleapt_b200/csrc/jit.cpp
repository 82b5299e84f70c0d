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
    {
        std::ofstream f(cu);
        if (!f) throw Error("cannot write " + cu);
        f << source;
    }
    if (access(nvcc.c_str(), X_OK) != 0) throw Error("nvcc not found at " + nvcc + " (set APTB200_NVCC); the model cannot be compiled");
    const std::string tmp = base + ".tmp." + std::to_string((long)getpid()) + ".so";
    const std::string cmd = nvcc + " " + flags + " -I" + inc + " " + cu + " -o " + tmp + " > " + log + " 2>&1";
    if (verbose) fprintf(stderr, "[aptb200] %s\n", cmd.c_str());
    int rc = system(cmd.c_str());
    if (rc != 0) {
        std::string l = slurp(log);
        if (l.size() > 4000) l = l.substr(0, 4000) + "\n...";
        unlink(tmp.c_str());
        throw Error("nvcc failed for the lowered model (source kept at " + cu + "):\n" + l);
    }
    if (rename(tmp.c_str(), so.c_str()) != 0) throw Error("cannot move the compiled module into the cache");
    return so;
}

}  // namespace aptf
