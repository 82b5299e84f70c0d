// front.h — host-side front end of libaptb200.so: BUGS-subset AST, model description, lowering, JIT.
// Replaces (for the supported subset) what nimble does between nimbleModel() and compileNimble():
// declarations -> nodes, getDependencies(target) (APT_functions.R:671,793), code generation + compile
// (APT_functions.R:157-162 documents that the reference must be compiled by nimble at this point).
#pragma once
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "../../include/aptb200.h"

namespace aptf {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Expr;
using EP = std::shared_ptr<Expr>;
struct Expr {
    enum Kind { NUM, VAR, NEG, BIN, CALL, RANGE } k = NUM;
    double num = 0;
    std::string name;  // VAR / CALL
    char op = 0;       // BIN: + - * / ^
    std::vector<EP> a; // VAR: indices; BIN: 2; NEG: 1; CALL: positional args; RANGE: lo, hi
    std::vector<std::pair<std::string, EP>> kw;
};
EP mk_num(double v);
EP mk_var(const std::string& n, std::vector<EP> idx = {});
EP mk_bin(char op, EP l, EP r);
EP mk_call(const std::string& f, std::vector<EP> args);
std::string to_string(const EP& e);

struct Loop {
    std::string var;
    EP lo, hi;
};
struct RawDecl {
    std::vector<Loop> loops;
    EP lhs;  // VAR (possibly wrapped link already removed)
    bool stoch = false;
    EP rhs;
    std::string link;
    int line = 0;
};
std::vector<RawDecl> parse_bugs(const std::string& code);
// A user-supplied deterministic function (nimbleFunction with a `run` function of scalar / vector double arguments:
// assignments, if / else, return, arithmetic, comparisons, ceiling / floor / abs ..., indexed loads; demo/APT_demo.R:42-71).
// It is inlined where the model calls it: `call` executes the body symbolically on the argument expressions.
struct UserFunction {
    struct Param { std::string name; int ndim = 0; };
    std::string name;
    std::vector<Param> params;
    std::function<EP(const std::vector<EP>&, const std::vector<std::string>&)> call;
};
UserFunction parse_user_function(const std::string& name, const std::string& r_source);
EP expand_user_calls(const EP& e, const std::map<std::string, UserFunction>& fns, int depth = 0);
EP parse_expr_text(const std::string& s);  // e.g. a sampler target "beta[1:50]"

struct Array {
    std::string name;
    int role = 0;
    std::vector<int64_t> dims;
    std::vector<double> v;
    int64_t size() const { return (int64_t)v.size(); }
    int64_t nrow() const { return dims.empty() ? 1 : dims[0]; }
};

struct SamplerConf {
    std::string target, type;
    apt_sampler_control ctl;
    std::vector<double> propCov;
};

struct ModelSpec {
    std::string code;
    std::vector<RawDecl> decls;  // user-supplied functions already inlined
    std::map<std::string, std::string> functions;  // name -> R source, as given to apt_model_add_function
    std::map<std::string, Array> constants, data, inits;
    std::vector<SamplerConf> samplers;
    std::vector<std::string> monitors, monitors2;
    int thin = 1, thin2 = 1;
};

// Result of the lowering step: CUDA source + everything the module needs at create time.
struct Lowered {
    std::string source;
    int P = 0, K = 0, NU = 0, ND = 0, NMON = 0, NMON2 = 0, DMAX = 1, W = 1, engine = 0;
    std::vector<std::string> array_names;       // order of Data::a[]
    std::vector<const Array*> arrays;
    std::vector<std::vector<int>> iarrays;      // order of Data::ia[]
    std::vector<double> theta0;
    std::vector<std::string> param_names, mon_names, mon2_names;
};
Lowered lower_model(const ModelSpec& spec, int n_temps, const apt_build_opts& opts);

// JIT: compile `source` with nvcc for sm_100a (cached by content hash) and return the path of the .so
std::string jit_compile(const std::string& source, bool verbose);
std::string engine_include_dir();

}  // namespace aptf
