// lower.cpp — the lowering step: BUGS subset -> CUDA device log-probability code for sm_100a.
//
// Stands in for nimble's model compilation and for the sampler `setup` functions of the reference
// (/root/reference/nimbleAPT/R/APT_functions.R:652-689, 773-820): canonical parameterisations, deterministic
// nodes inlined into their consumers, `model$getDependencies(target)` computed concretely per sampler, and the
// result emitted as C++/CUDA that is instantiated into the engine kernels (engine/apt_engine.cuh).
//
// Per sampler the dependent stochastic nodes are grouped by declaration:
//   FULL    – every instance of the declaration depends on the target: the engine keeps the declaration's
//             log-probability SUM per chain and compares new sum vs stored sum (1 pass over the data);
//   PARTIAL – a subset of instances (a prior term, the observations of one random-effect group, ...):
//             old and new terms are both evaluated over the explicit instance list.
// Runs of scalar samplers on consecutive elements of one variable whose dependent sets are disjoint (e.g. the
// 1000 random effects of a GLMM) are recognised as a FAMILY and executed in parallel; this is exactly equivalent
// to the reference's serial order (APT:440-444) because no member's update reads another member's target.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <set>
#include <sstream>
#include "front.h"

namespace aptf {
namespace {

struct Var {
    std::string name;
    enum Role { PARAM, DATA, CONST, DET } role = CONST;
    std::vector<int64_t> dims;
    int64_t size = 1;
    int64_t theta_off = -1;
    const Array* arr = nullptr;
    int arr_idx = -1;
    int64_t nrow() const { return dims.empty() ? 1 : dims[0]; }
};
struct LoopC {
    std::string var;
    long lo, hi;
};
struct CDecl {
    std::vector<LoopC> loops;
    EP lhs;
    bool stoch = false;
    std::string dist;
    std::vector<EP> args;          // canonical (not inlined)
    std::vector<EP> iargs;         // deterministic nodes inlined (scalar args)
    std::vector<EP> mean_elems;    // dmnorm
    std::vector<EP> chol_elems;    // dmnorm, column-major n*n
    int mvn_n = 0;
    bool mvn_factor = false;       // dmnorm(prec = / cov = constant matrix): chol_elems are computed at build time
    int lpidx = -1;
    bool lhs_param = false;
    long ninst = 1;
    int line = 0;
};
struct Pattern {  // one read of (or the write to) a PARAM variable by a stochastic declaration
    const Var* v;
    EP ref;
    bool own;
    bool loop_dep;
    std::vector<std::vector<int>> by_elem;  // element -> instances (loop-dependent, non-range)
    bool built = false;
};
struct Group {
    int decl;
    bool full, prior;
    std::vector<int> inst;  // linear instance indices when !full
    bool big = false;       // grid engine: evaluated by the grid-wide pass over the data
};
struct Unit {  // one generated sampler struct: a single sampler or a family
    int kind = 0, d = 1, unit0 = 0, count = 1;
    bool discrete = false;  // slice sampler: model$isDiscrete(target) (APT:926)
    double ntotal = 0;      // multinomial sampler: Ntotal (APT:1024)
    std::vector<int> tgt;  // theta indices of member 0
    const SamplerConf* conf = nullptr;
    std::vector<Group> groups;
    // family CSR arrays (Data::ia indices) per group: ptr, idx
    std::vector<int> ia_ptr, ia_idx;
    // explicit lists for non-family partial groups
    std::vector<int> ia_list;
    int blkoff = 0, empoff = 0;
    EP lower, upper;
    std::map<std::string, double> bound_env;
};

std::string fmt_d(double v) {
    if (std::isinf(v)) return v > 0 ? "CUDART_INF" : "(-CUDART_INF)";
    if (std::isnan(v)) return "CUDART_NAN";
    char buf[64];
    snprintf(buf, sizeof buf, "%.17g", v);
    std::string s = buf;
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
    return s;
}

struct Lowerer {
    const ModelSpec& spec;
    int K;
    apt_build_opts opts;
    std::map<std::string, Var> vars;
    std::vector<std::string> param_order;
    std::vector<CDecl> decls;
    std::vector<int> stoch_decls;
    std::vector<std::vector<Pattern>> patterns;  // per stochastic decl (index = position in decls)
    int P = 0, ND = 0;
    std::vector<const Array*> used_arrays;
    std::vector<std::vector<int>> iarrays;
    std::vector<Unit> units;
    std::ostringstream out;

    Lowerer(const ModelSpec& s, int k, const apt_build_opts& o) : spec(s), K(k), opts(o) {}

    // ------------------------------------------------------------------ constants
    const Array* find_const(const std::string& n) const {
        auto it = spec.constants.find(n);
        if (it != spec.constants.end()) return &it->second;
        return nullptr;
    }
    double ceval(const EP& e, const std::map<std::string, double>& env) const {
        switch (e->k) {
            case Expr::NUM: return e->num;
            case Expr::NEG: return -ceval(e->a[0], env);
            case Expr::BIN: {
                double a = ceval(e->a[0], env), b = ceval(e->a[1], env);
                switch (e->op) {
                    case '+': return a + b;
                    case '-': return a - b;
                    case '*': return a * b;
                    case '/': return a / b;
                    case '^': return std::pow(a, b);
                }
                break;
            }
            case Expr::VAR: {
                if (e->a.empty()) {
                    auto it = env.find(e->name);
                    if (it != env.end()) return it->second;
                }
                const Array* c = find_const(e->name);
                if (!c) {
                    auto it = spec.data.find(e->name);
                    if (it != spec.data.end() && !vars.count(e->name)) c = &it->second;
                    else if (it != spec.data.end() && vars.at(e->name).role == Var::CONST) c = &it->second;
                }
                if (c) {
                    if (e->a.empty()) {
                        if (c->size() != 1) throw Error("array constant '" + e->name + "' used as a scalar");
                        return c->v[0];
                    }
                    int64_t off = 0, stride = 1;
                    if (e->a.size() != c->dims.size()) throw Error("wrong number of indices for '" + e->name + "'");
                    for (size_t k = 0; k < e->a.size(); k++) {
                        if (e->a[k]->k == Expr::RANGE) throw Error("range where a scalar index is required: " + to_string(e));
                        long idx = (long)std::llround(ceval(e->a[k], env));
                        if (idx < 1 || idx > c->dims[k]) throw Error("index out of range: " + to_string(e));
                        off += (idx - 1) * stride;
                        stride *= c->dims[k];
                    }
                    return c->v[off];
                }
                throw Error("not a constant: " + e->name);
            }
            case Expr::CALL: {
                if (e->name == "length" && e->a.size() == 1 && e->a[0]->k == Expr::VAR) {
                    const Array* c = find_const(e->a[0]->name);
                    if (c) return (double)c->size();
                }
                if (e->a.size() == 1) {
                    double x = ceval(e->a[0], env);
                    if (e->name == "exp") return std::exp(x);
                    if (e->name == "log") return std::log(x);
                    if (e->name == "sqrt") return std::sqrt(x);
                    if (e->name == "abs") return std::fabs(x);
                    if (e->name == "ceiling") return std::ceil(x);
                    if (e->name == "floor") return std::floor(x);
                    if (e->name == "round") return std::nearbyint(x);
                    if (e->name == "trunc") return std::trunc(x);
                    if (e->name == "__not") return x == 0.0 ? 1.0 : 0.0;
                }
                if (e->a.size() == 2 && e->name.rfind("__", 0) == 0) {
                    double x = ceval(e->a[0], env), y = ceval(e->a[1], env);
                    if (e->name == "__lt") return x < y;
                    if (e->name == "__gt") return x > y;
                    if (e->name == "__le") return x <= y;
                    if (e->name == "__ge") return x >= y;
                    if (e->name == "__eq") return x == y;
                    if (e->name == "__ne") return x != y;
                    if (e->name == "__or") return (x != 0.0) || (y != 0.0);
                    if (e->name == "__and") return (x != 0.0) && (y != 0.0);
                }
                if (e->name == "__if" && e->a.size() == 3) return ceval(e->a[0], env) != 0.0 ? ceval(e->a[1], env) : ceval(e->a[2], env);
                break;
            }
            default: break;
        }
        throw Error("expression is not a compile-time constant: " + to_string(e));
    }
    bool try_ceval(const EP& e, const std::map<std::string, double>& env, double& v) const {
        try {
            v = ceval(e, env);
            return true;
        } catch (const Error&) {
            return false;
        }
    }

    // ------------------------------------------------------------------ canonical declarations
    static EP kwarg(const EP& call, const char* name, size_t pos) {
        for (auto& kv : call->kw)
            if (kv.first == name) return kv.second;
        // positional arguments fill the formals that were not named, in order
        return pos < call->a.size() ? call->a[pos] : nullptr;
    }
    static bool has_kw(const EP& call, const char* name) {
        for (auto& kv : call->kw)
            if (kv.first == name) return true;
        return false;
    }

    EP substitute(const EP& e, const std::map<std::string, EP>& m) const {
        if (e->k == Expr::VAR && e->a.empty()) {
            auto it = m.find(e->name);
            if (it != m.end()) return it->second;
            return e;
        }
        auto r = std::make_shared<Expr>(*e);
        for (auto& x : r->a) x = substitute(x, m);
        for (auto& x : r->kw) x.second = substitute(x.second, m);
        return r;
    }

    // user-supplied deterministic functions (apt_model_add_function) are inlined where the model calls them
    std::vector<RawDecl> expanded_decls;
    void expand_functions() {
        expanded_decls = spec.decls;
        if (spec.functions.empty()) return;
        std::map<std::string, UserFunction> fns;
        for (auto& kv : spec.functions) fns[kv.first] = parse_user_function(kv.first, kv.second);
        for (auto& rd : expanded_decls) rd.rhs = expand_user_calls(rd.rhs, fns);
    }

    void canonicalise() {
        int idx = 0;
        expand_functions();
        for (const RawDecl& rd : expanded_decls) {
            CDecl d;
            d.line = rd.line;
            for (auto& l : rd.loops) {
                LoopC lc;
                lc.var = l.var;
                lc.lo = (long)std::llround(ceval(l.lo, {}));
                lc.hi = (long)std::llround(ceval(l.hi, {}));
                if (lc.hi < lc.lo) throw Error("empty loop range for '" + l.var + "'");
                d.loops.push_back(lc);
            }
            d.lhs = rd.lhs;
            d.stoch = rd.stoch;
            if (rd.stoch) {
                EP c = rd.rhs;
                std::string dist = c->name;
                EP one = mk_num(1.0);
                auto need = [&](EP e, const char* what) {
                    if (!e) throw Error("line " + std::to_string(rd.line) + ": " + dist + " is missing argument '" + what + "'");
                    return e;
                };
                if (dist == "dbern") {
                    d.dist = "dbinom";
                    d.args = {need(kwarg(c, "prob", 0), "prob"), one};
                } else if (dist == "dnorm") {
                    d.dist = dist;
                    EP mean = need(kwarg(c, "mean", 0), "mean"), sd;
                    if (has_kw(c, "sd")) sd = kwarg(c, "sd", 99);
                    else if (has_kw(c, "var")) sd = mk_call("sqrt", {kwarg(c, "var", 99)});
                    else sd = mk_bin('/', one, mk_call("sqrt", {need(kwarg(c, "tau", 1), "tau")}));
                    d.args = {mean, sd};
                } else if (dist == "dunif") {
                    d.dist = dist;
                    d.args = {need(kwarg(c, "min", 0), "min"), need(kwarg(c, "max", 1), "max")};
                } else if (dist == "dgamma") {
                    d.dist = dist;
                    EP shape = need(kwarg(c, "shape", 0), "shape");
                    EP scale = has_kw(c, "scale") ? kwarg(c, "scale", 99) : mk_bin('/', one, need(kwarg(c, "rate", 1), "rate"));
                    d.args = {shape, scale};
                } else if (dist == "dbeta") {
                    d.dist = dist;
                    d.args = {need(kwarg(c, "shape1", 0), "shape1"), need(kwarg(c, "shape2", 1), "shape2")};
                } else if (dist == "dbinom") {
                    d.dist = dist;
                    EP prob, size;
                    const bool kp = has_kw(c, "prob"), ks = has_kw(c, "size");
                    if (kp && ks) { prob = kwarg(c, "prob", 99); size = kwarg(c, "size", 99); }
                    else if (kp) { prob = kwarg(c, "prob", 99); size = c->a.empty() ? nullptr : c->a[0]; }
                    else if (ks) { size = kwarg(c, "size", 99); prob = c->a.empty() ? nullptr : c->a[0]; }
                    else { prob = c->a.size() > 0 ? c->a[0] : nullptr; size = c->a.size() > 1 ? c->a[1] : nullptr; }
                    d.args = {need(prob, "prob"), need(size, "size")};
                } else if (dist == "dpois") {
                    d.dist = dist;
                    d.args = {need(kwarg(c, "lambda", 0), "lambda")};
                } else if (dist == "dexp") {
                    d.dist = dist;
                    d.args = {has_kw(c, "scale") ? kwarg(c, "scale", 99) : mk_bin('/', one, need(kwarg(c, "rate", 0), "rate"))};
                } else if (dist == "dcat") {  // demo/APT_demo.R:84
                    d.dist = dist;
                    d.args = {need(kwarg(c, "prob", 0), "prob")};
                } else if (dist == "dmulti") {  // target distribution of sampler_RW_multinomial_tempered (APT:1046)
                    d.dist = dist;
                    d.args = {need(kwarg(c, "prob", 0), "prob"), need(kwarg(c, "size", 1), "size")};
                } else if (dist == "dmnorm") {
                    d.dist = dist;
                    if (has_kw(c, "cholesky")) {
                        d.args = {need(kwarg(c, "mean", 0), "mean"), kwarg(c, "cholesky", 99),
                                  has_kw(c, "prec_param") ? kwarg(c, "prec_param", 99) : one};
                    } else {
                        // dmnorm(mean, prec) / dmnorm(mean, prec = ) / dmnorm(mean, cov = ) with a CONSTANT matrix: nimble takes
                        // chol() of it once and calls dmnorm_chol; so does the lowering (inline_all), at build time
                        const bool cov = has_kw(c, "cov");
                        EP m = cov ? kwarg(c, "cov", 99) : has_kw(c, "prec") ? kwarg(c, "prec", 99) : (c->a.size() > 1 ? c->a[1] : nullptr);
                        if (!m) throw Error("line " + std::to_string(rd.line) + ": dmnorm needs cholesky = , prec = or cov =");
                        d.args = {need(kwarg(c, "mean", 0), "mean"), m, cov ? mk_num(0) : one};
                        d.mvn_factor = true;
                    }
                } else {
                    throw Error("line " + std::to_string(rd.line) + ": unsupported distribution '" + dist +
                                "' (supported: dnorm dunif dgamma dbeta dbinom dbern dpois dexp dcat dmulti dmnorm)");
                }
            } else {
                EP rhs = rd.rhs;
                if (!rd.link.empty()) {
                    static const std::map<std::string, std::string> inv = {{"log", "exp"}, {"logit", "ilogit"}, {"probit", "iprobit"}, {"cloglog", "icloglog"}};
                    auto it = inv.find(rd.link);
                    if (it == inv.end()) throw Error("unsupported link function '" + rd.link + "'");
                    rhs = mk_call(it->second, {rhs});
                }
                // scalarise element-wise vector declarations  v[a:b] <- f(w[c:d])
                int rdim = -1;
                for (size_t k = 0; k < d.lhs->a.size(); k++)
                    if (d.lhs->a[k]->k == Expr::RANGE) {
                        if (rdim >= 0) throw Error("matrix-valued deterministic declarations are not supported");
                        rdim = (int)k;
                    }
                if (rdim >= 0) {
                    long lo = (long)std::llround(ceval(d.lhs->a[rdim]->a[0], {}));
                    long hi = (long)std::llround(ceval(d.lhs->a[rdim]->a[1], {}));
                    std::string ev = "_e" + std::to_string(idx);
                    d.loops.push_back({ev, lo, hi});
                    auto lhs = std::make_shared<Expr>(*d.lhs);
                    lhs->a[rdim] = mk_var(ev);
                    d.lhs = lhs;
                    std::function<EP(const EP&)> scal = [&](const EP& e) -> EP {
                        auto r = std::make_shared<Expr>(*e);
                        if (e->k == Expr::CALL && (e->name == "inprod" || e->name == "sum"))
                            throw Error("vector reductions inside vector declarations are not supported");
                        for (auto& x : r->a) {
                            if (e->k == Expr::VAR && x->k == Expr::RANGE) {
                                long l2 = (long)std::llround(ceval(x->a[0], {}));
                                x = mk_bin('+', mk_var(ev), mk_num((double)(l2 - lo)));
                            } else {
                                x = scal(x);
                            }
                        }
                        return r;
                    };
                    rhs = scal(rhs);
                }
                d.args = {rhs};
            }
            d.ninst = 1;
            for (auto& l : d.loops) d.ninst *= (l.hi - l.lo + 1);
            if (d.loops.size() > 3) throw Error("more than 3 nested loops are not supported");
            decls.push_back(d);
            idx++;
        }
    }

    // ------------------------------------------------------------------ variables
    void build_vars() {
        for (auto& d : decls) {
            const std::string& name = d.lhs->name;
            Var::Role role = d.stoch ? (spec.data.count(name) ? Var::DATA : Var::PARAM) : Var::DET;
            const Array* src = nullptr;
            if (role == Var::DATA) src = &spec.data.at(name);
            if (role == Var::PARAM && spec.inits.count(name)) src = &spec.inits.at(name);
            std::vector<int64_t> dims;
            if (src && !(src->dims.empty())) dims = src->dims;
            else {
                // extent of the LHS indices over the loop ranges
                for (auto& ix : d.lhs->a) {
                    if (ix->k == Expr::RANGE) { dims.push_back((int64_t)std::llround(ceval(ix->a[1], {}))); continue; }
                    double best = 0;
                    // evaluate at loop corners
                    size_t nl = d.loops.size();
                    for (size_t mask = 0; mask < (1u << nl); mask++) {
                        std::map<std::string, double> env;
                        for (size_t l = 0; l < nl; l++) env[d.loops[l].var] = (mask >> l & 1) ? d.loops[l].hi : d.loops[l].lo;
                        best = std::max(best, ceval(ix, env));
                    }
                    dims.push_back((int64_t)std::llround(best));
                }
            }
            auto it = vars.find(name);
            if (it == vars.end()) {
                Var v;
                v.name = name; v.role = role; v.dims = dims; v.arr = src;
                vars[name] = v;
                if (role == Var::PARAM) param_order.push_back(name);
            } else {
                if (it->second.role != role) throw Error("variable '" + name + "' is declared with mixed roles");
                if (!src)
                    for (size_t k = 0; k < dims.size() && k < it->second.dims.size(); k++)
                        it->second.dims[k] = std::max(it->second.dims[k], dims[k]);
            }
        }
        auto add_const = [&](const std::map<std::string, Array>& m) {
            for (auto& kv : m)
                if (!vars.count(kv.first) && kv.second.size() > 1) {
                    Var v;
                    v.name = kv.first; v.role = Var::CONST; v.dims = kv.second.dims; v.arr = &kv.second;
                    vars[kv.first] = v;
                }
        };
        add_const(spec.constants);
        add_const(spec.data);  // right-hand-side-only "data" (covariates) behave as constants
        P = 0;
        for (auto& kv : vars) {
            Var& v = kv.second;
            v.size = 1;
            for (auto dd : v.dims) v.size *= dd;
            if (v.arr && v.arr->size() != v.size)
                throw Error("array '" + v.name + "' has " + std::to_string(v.arr->size()) + " values but the model needs " + std::to_string(v.size));
        }
        for (auto& n : param_order) {
            Var& v = vars[n];
            if (!v.arr) throw Error("no initial values for parameter '" + n + "' (pass them as inits)");
            v.theta_off = P;
            P += (int)v.size;
        }
        ND = 0;
        for (size_t i = 0; i < decls.size(); i++)
            if (decls[i].stoch) {
                decls[i].lpidx = ND++;
                decls[i].lhs_param = vars[decls[i].lhs->name].role == Var::PARAM;
                stoch_decls.push_back((int)i);
            }
    }

    // ------------------------------------------------------------------ deterministic-node inlining
    EP inline_det(const EP& e, int depth = 0) const {
        if (depth > 32) throw Error("deterministic nodes nest too deeply (cycle?)");
        if (e->k == Expr::VAR) {
            auto it = vars.find(e->name);
            if (it != vars.end() && it->second.role == Var::DET) {
                for (auto& ix : e->a)
                    if (ix->k == Expr::RANGE) return e;  // vector context: expanded element-wise by the caller
                // find the declaration that defines this element
                const CDecl* best = nullptr;
                std::map<std::string, EP> bind;
                int nmatch = 0;
                for (auto& d : decls) {
                    if (d.stoch || d.lhs->name != e->name) continue;
                    if (d.lhs->a.size() != e->a.size()) continue;
                    std::map<std::string, EP> b;
                    bool ok = true;
                    for (size_t k = 0; k < e->a.size() && ok; k++) {
                        const EP& li = d.lhs->a[k];
                        if (li->k == Expr::VAR && li->a.empty()) {
                            bool isloop = false;
                            for (auto& l : d.loops)
                                if (l.var == li->name) isloop = true;
                            if (isloop) { b[li->name] = inline_det(e->a[k], depth + 1); continue; }
                        }
                        double lv, rv;
                        if (try_ceval(li, {}, lv)) {
                            if (try_ceval(e->a[k], {}, rv)) ok = (lv == rv);
                            else ok = false;  // cannot decide statically
                        } else {
                            throw Error("unsupported index pattern on the left-hand side of deterministic node '" + e->name + "'");
                        }
                    }
                    if (ok) { best = &d; bind = b; nmatch++; }
                }
                if (!best) throw Error("no deterministic declaration matches " + to_string(e));
                if (nmatch > 1) {
                    // several loop-declared pieces: choose by constant index range if possible
                    nmatch = 0;
                    for (auto& d : decls) {
                        if (d.stoch || d.lhs->name != e->name) continue;
                        bool ok = true;
                        std::map<std::string, EP> b;
                        for (size_t k = 0; k < e->a.size() && ok; k++) {
                            const EP& li = d.lhs->a[k];
                            double rv;
                            if (li->k == Expr::VAR && li->a.empty()) {
                                const LoopC* lc = nullptr;
                                for (auto& l : d.loops) if (l.var == li->name) lc = &l;
                                if (lc) {
                                    if (try_ceval(e->a[k], {}, rv)) ok = rv >= lc->lo && rv <= lc->hi;
                                    b[li->name] = inline_det(e->a[k], depth + 1);
                                    continue;
                                }
                            }
                            double lv;
                            if (try_ceval(li, {}, lv) && try_ceval(e->a[k], {}, rv)) ok = lv == rv; else ok = false;
                        }
                        if (ok) { best = &d; bind = b; nmatch++; }
                    }
                    if (nmatch != 1) throw Error("ambiguous deterministic declaration for " + to_string(e));
                }
                EP body = substitute(best->args[0], bind);
                return inline_det(body, depth + 1);
            }
        }
        auto r = std::make_shared<Expr>(*e);
        for (auto& x : r->a) x = inline_det(x, depth);
        for (auto& x : r->kw) x.second = inline_det(x.second, depth);
        return r;
    }

    // element e (0-based) of a vector-valued argument (a slice  v[.., lo:hi, ..]  or a whole vector)
    std::vector<EP> vector_elements(const EP& e) const {
        if (e->k != Expr::VAR) throw Error("vector argument must be a variable slice: " + to_string(e));
        auto it = vars.find(e->name);
        std::vector<EP> out;
        int rdim = -1;
        for (size_t k = 0; k < e->a.size(); k++)
            if (e->a[k]->k == Expr::RANGE) rdim = (int)k;
        if (e->a.empty()) {
            int64_t n = it != vars.end() ? it->second.size : 0;
            const Array* c = find_const(e->name);
            if (it == vars.end() && c) n = c->size();
            for (int64_t i = 0; i < n; i++) out.push_back(mk_var(e->name, {mk_num((double)(i + 1))}));
            return out;
        }
        if (rdim < 0) throw Error("expected a vector slice: " + to_string(e));
        long lo = (long)std::llround(ceval(e->a[rdim]->a[0], {})), hi = (long)std::llround(ceval(e->a[rdim]->a[1], {}));
        for (long i = lo; i <= hi; i++) {
            auto r = std::make_shared<Expr>(*e);
            r->a[rdim] = mk_num((double)i);
            out.push_back(r);
        }
        return out;
    }

    void inline_all() {
        for (auto& d : decls) {
            if (!d.stoch) continue;
            if (d.dist == "dmnorm") {
                for (auto& el : vector_elements(d.args[0])) d.mean_elems.push_back(inline_det(el));
                d.mvn_n = (int)d.mean_elems.size();
                if (d.mvn_n < 1 || d.mvn_n > 16) throw Error("dmnorm dimension must be between 1 and 16");
                const EP& ch = d.args[1];
                if (ch->k != Expr::VAR) throw Error("cholesky= / prec= / cov= must be a matrix variable");
                long r0 = 1, c0 = 1;
                if (!ch->a.empty()) {
                    if (ch->a.size() != 2 || ch->a[0]->k != Expr::RANGE || ch->a[1]->k != Expr::RANGE)
                        throw Error("cholesky= must be given as M[a:b, c:d]");
                    r0 = (long)std::llround(ceval(ch->a[0]->a[0], {}));
                    c0 = (long)std::llround(ceval(ch->a[1]->a[0], {}));
                }
                for (int j = 0; j < d.mvn_n; j++)
                    for (int i = 0; i < d.mvn_n; i++)
                        d.chol_elems.push_back(inline_det(mk_var(ch->name, {mk_num((double)(r0 + i)), mk_num((double)(c0 + j))})));
                // reorder: chol_elems is filled column by column -> index i + j*n
                if (d.mvn_factor) {
                    // upper factor U of the constant matrix M (U'U = M, R's chol()), column-major like the elements above
                    const int n = d.mvn_n;
                    std::vector<double> M((size_t)n * n), U((size_t)n * n, 0.0);
                    for (int e = 0; e < n * n; e++)
                        if (!try_ceval(d.chol_elems[e], {}, M[e]))
                            throw Error("dmnorm(prec = / cov = ) needs a constant matrix (pass cholesky = for one that depends on parameters): " + to_string(ch));
                    for (int j = 0; j < n; j++) {
                        double sjj = M[j + j * n];
                        for (int k = 0; k < j; k++) sjj -= U[k + j * n] * U[k + j * n];
                        if (!(sjj > 0.0)) throw Error("dmnorm: the matrix is not positive definite: " + to_string(ch));
                        const double dj = std::sqrt(sjj);
                        U[j + j * n] = dj;
                        for (int i = j + 1; i < n; i++) {
                            double t = M[j + i * n];
                            for (int k = 0; k < j; k++) t -= U[k + j * n] * U[k + i * n];
                            U[j + i * n] = t / dj;
                        }
                    }
                    for (int e = 0; e < n * n; e++) d.chol_elems[e] = mk_num(U[e]);
                }
                d.iargs = {inline_det(d.args[2])};
                int ln = 0;
                for (auto& ix : d.lhs->a)
                    if (ix->k == Expr::RANGE) ln = (int)(std::llround(ceval(ix->a[1], {})) - std::llround(ceval(ix->a[0], {})) + 1);
                if (ln != d.mvn_n) throw Error("dmnorm: node and mean have different lengths");
            } else if (d.dist == "dmulti") {
                for (auto& el : vector_elements(d.args[0])) d.mean_elems.push_back(inline_det(el));
                d.mvn_n = (int)d.mean_elems.size();
                if (d.mvn_n < 2 || d.mvn_n > 16) throw Error("dmulti: the probability vector must have between 2 and 16 elements");
                d.iargs = {inline_det(d.args[1])};
                int ln = 0;
                for (auto& ix : d.lhs->a)
                    if (ix->k == Expr::RANGE) ln = (int)(std::llround(ceval(ix->a[1], {})) - std::llround(ceval(ix->a[0], {})) + 1);
                if (ln != d.mvn_n) throw Error("dmulti: node and prob have different lengths");
            } else if (d.dist == "dcat") {
                // the probability vector, element by element (kept in mean_elems so that the dependency analysis sees it)
                for (auto& el : vector_elements(d.args[0])) d.mean_elems.push_back(inline_det(el));
                d.mvn_n = (int)d.mean_elems.size();
                if (d.mvn_n < 1 || d.mvn_n > 64) throw Error("dcat: the probability vector must have between 1 and 64 elements");
            } else {
                for (auto& a : d.args) d.iargs.push_back(inline_det(a));
            }
        }
    }

    // ------------------------------------------------------------------ dependency analysis
    void collect_param_reads(const EP& e, std::vector<EP>& acc) const {
        if (e->k == Expr::VAR) {
            auto it = vars.find(e->name);
            if (it != vars.end() && it->second.role == Var::PARAM) acc.push_back(e);
        }
        for (auto& x : e->a) collect_param_reads(x, acc);
        for (auto& x : e->kw) collect_param_reads(x.second, acc);
    }
    bool depends_on_loops(const EP& e, const CDecl& d) const {
        if (e->k == Expr::VAR && e->a.empty())
            for (auto& l : d.loops)
                if (l.var == e->name) return true;
        for (auto& x : e->a)
            if (depends_on_loops(x, d)) return true;
        return false;
    }
    void inst_env(const CDecl& d, long inst, std::map<std::string, double>& env) const {
        for (auto& l : d.loops) {
            long n = l.hi - l.lo + 1;
            env[l.var] = (double)(l.lo + inst % n);
            inst /= n;
        }
    }
    // theta elements (relative to the variable) touched by a reference at a given instance
    void ref_elements(const Var& v, const EP& ref, const std::map<std::string, double>& env, std::vector<int64_t>& out) const {
        out.clear();
        if (ref->a.empty()) {
            for (int64_t i = 0; i < v.size; i++) out.push_back(i);
            return;
        }
        if (ref->a.size() != v.dims.size()) throw Error("wrong number of indices: " + to_string(ref));
        std::vector<int64_t> cur = {0};
        int64_t stride = 1;
        for (size_t k = 0; k < ref->a.size(); k++) {
            long lo, hi;
            if (ref->a[k]->k == Expr::RANGE) {
                lo = (long)std::llround(ceval(ref->a[k]->a[0], env));
                hi = (long)std::llround(ceval(ref->a[k]->a[1], env));
            } else {
                lo = hi = (long)std::llround(ceval(ref->a[k], env));
            }
            if (lo < 1 || hi > v.dims[k]) throw Error("index out of range: " + to_string(ref));
            std::vector<int64_t> nxt;
            for (auto base : cur)
                for (long i = lo; i <= hi; i++) nxt.push_back(base + (i - 1) * stride);
            cur.swap(nxt);
            stride *= v.dims[k];
        }
        out = cur;
    }

    void build_patterns() {
        patterns.resize(decls.size());
        for (int di : stoch_decls) {
            CDecl& d = decls[di];
            std::vector<EP> reads;
            for (auto& a : d.iargs) collect_param_reads(a, reads);
            for (auto& a : d.mean_elems) collect_param_reads(a, reads);
            for (auto& a : d.chol_elems) collect_param_reads(a, reads);
            std::set<std::string> seen;
            for (auto& r : reads) {
                std::string key = to_string(r);
                if (!seen.insert(key).second) continue;
                Pattern p;
                p.v = &vars.at(r->name);
                p.ref = r;
                p.own = false;
                p.loop_dep = depends_on_loops(r, d);
                patterns[di].push_back(p);
            }
            if (d.lhs_param) {
                Pattern p;
                p.v = &vars.at(d.lhs->name);
                p.ref = d.lhs;
                p.own = true;
                p.loop_dep = depends_on_loops(d.lhs, d);
                patterns[di].push_back(p);
            }
        }
    }
    void build_inverse(int di, Pattern& p) {
        if (p.built) return;
        const CDecl& d = decls[di];
        p.by_elem.assign((size_t)p.v->size, {});
        std::map<std::string, double> env;
        std::vector<int64_t> els;
        for (long inst = 0; inst < d.ninst; inst++) {
            inst_env(d, inst, env);
            ref_elements(*p.v, p.ref, env, els);
            for (auto e : els) p.by_elem[(size_t)e].push_back((int)inst);
        }
        p.built = true;
    }

    std::vector<Group> dependencies(const std::vector<int>& tgt) {
        std::vector<char> inT((size_t)P, 0);
        for (int t : tgt) inT[(size_t)t] = 1;
        std::vector<Group> groups;
        for (int di : stoch_decls) {
            const CDecl& d = decls[di];
            std::vector<char> own((size_t)d.ninst, 0), dep((size_t)d.ninst, 0);
            bool all_dep = false;
            for (auto& p : patterns[di]) {
                const int64_t base = p.v->theta_off;
                bool touches = false;
                for (int t : tgt)
                    if (t >= base && t < base + p.v->size) touches = true;
                if (!touches) continue;
                if (!p.loop_dep) {
                    std::vector<int64_t> els;
                    ref_elements(*p.v, p.ref, {}, els);
                    bool hit = false;
                    for (auto e : els) hit |= inT[(size_t)(base + e)] != 0;
                    if (hit) {
                        if (p.own) std::fill(own.begin(), own.end(), 1);
                        else all_dep = true;
                    }
                } else {
                    build_inverse(di, p);
                    for (int t : tgt) {
                        if (t < base || t >= base + p.v->size) continue;
                        for (int inst : p.by_elem[(size_t)(t - base)]) (p.own ? own : dep)[(size_t)inst] = 1;
                    }
                }
            }
            Group gp{di, false, true, {}}, gd{di, false, false, {}};
            for (long i = 0; i < d.ninst; i++) {
                if (own[(size_t)i]) gp.inst.push_back((int)i);
                else if (all_dep || dep[(size_t)i]) gd.inst.push_back((int)i);
            }
            if (!gp.inst.empty()) {
                gp.full = (long)gp.inst.size() == d.ninst;
                groups.push_back(gp);
            }
            if (!gd.inst.empty()) {
                gd.full = (long)gd.inst.size() == d.ninst;
                groups.push_back(gd);
            }
        }
        return groups;
    }

    // ------------------------------------------------------------------ targets
    std::vector<int> expand_target(const std::string& text) const {
        // "a[1:3]", "a", "c('a','b[2]')" and comma-separated lists
        std::string s = text;
        std::vector<std::string> parts;
        {
            std::string t = s;
            size_t c = t.find("c(");
            if (c == 0 && t.back() == ')') t = t.substr(2, t.size() - 3);
            int depth = 0;
            std::string cur;
            for (char ch : t) {
                if (ch == '[' || ch == '(') depth++;
                if (ch == ']' || ch == ')') depth--;
                if (ch == ',' && depth == 0) { parts.push_back(cur); cur.clear(); continue; }
                if (ch == '\'' || ch == '"' || ch == ' ') continue;
                cur += ch;
            }
            if (!cur.empty()) parts.push_back(cur);
        }
        std::vector<int> out;
        for (auto& part : parts) {
            EP e = parse_expr_text(part);
            if (e->k != Expr::VAR) throw Error("cannot parse node name '" + part + "'");
            auto it = vars.find(e->name);
            if (it == vars.end()) throw Error("unknown node '" + e->name + "'");
            if (it->second.role != Var::PARAM) throw Error("'" + e->name + "' is not a sampled (non-data stochastic) node");
            std::vector<int64_t> els;
            ref_elements(it->second, e, {}, els);
            for (auto el : els) out.push_back((int)(it->second.theta_off + el));
        }
        return out;
    }
    std::vector<std::string> theta_names() const {
        std::vector<std::string> out;
        for (auto& n : param_order) {
            const Var& v = vars.at(n);
            for (int64_t e = 0; e < v.size; e++) {
                if (v.size == 1 && v.dims.empty()) out.push_back(v.name);
                else if (v.dims.size() <= 1) out.push_back(v.name + "[" + std::to_string(e + 1) + "]");
                else out.push_back(v.name + "[" + std::to_string(e % v.dims[0] + 1) + ", " + std::to_string(e / v.dims[0] + 1) + "]");
            }
        }
        return out;
    }

    // ------------------------------------------------------------------ samplers -> units
    static bool same_ctl(const apt_sampler_control& a, const apt_sampler_control& b) {
        return a.log == b.log && a.reflective == b.reflective && a.adaptive == b.adaptive && a.adaptScaleOnly == b.adaptScaleOnly &&
               a.adaptInterval == b.adaptInterval && a.adaptFactorExponent == b.adaptFactorExponent && a.scale == b.scale &&
               a.temperPriors == b.temperPriors;
    }
    int add_iarray(const std::vector<int>& v) {
        iarrays.push_back(v);
        return (int)iarrays.size() - 1;
    }

    void set_bounds(Unit& u, int tgt) {
        // model$getBound(target, 'lower' | 'upper')  (APT:702-703) from the target's own distribution
        for (int di : stoch_decls) {
            const CDecl& d = decls[di];
            if (!d.lhs_param) continue;
            const Var& v = vars.at(d.lhs->name);
            if (tgt < v.theta_off || tgt >= v.theta_off + v.size) continue;
            std::map<std::string, double> env;
            std::vector<int64_t> els;
            for (long inst = 0; inst < d.ninst; inst++) {
                inst_env(d, inst, env);
                ref_elements(v, d.lhs, env, els);
                for (auto e : els)
                    if (v.theta_off + e == tgt) {
                        u.bound_env = env;
                        if (d.dist == "dunif") { u.lower = d.iargs[0]; u.upper = d.iargs[1]; }
                        else if (d.dist == "dgamma" || d.dist == "dexp") { u.lower = mk_num(0); u.upper = mk_num(INFINITY); }
                        else if (d.dist == "dbeta") { u.lower = mk_num(0); u.upper = mk_num(1); }
                        else { u.lower = mk_num(-INFINITY); u.upper = mk_num(INFINITY); }
                        return;
                    }
            }
        }
        throw Error("target has no stochastic declaration");
    }

    bool target_is_discrete(int tgt) const {  // model$isDiscrete(target): the target's own distribution
        for (int di : stoch_decls) {
            const CDecl& d = decls[di];
            if (!d.lhs_param) continue;
            const Var& v = vars.at(d.lhs->name);
            if (tgt >= v.theta_off && tgt < v.theta_off + v.size) return d.dist == "dcat" || d.dist == "dbinom" || d.dist == "dpois" || d.dist == "dmulti";
        }
        throw Error("target has no stochastic declaration");
    }

    void build_units() {
        const auto& S = spec.samplers;
        if (S.empty()) throw Error("no samplers configured: add sampler_RW_tempered / sampler_RW_block_tempered samplers");
        int unit = 0, blkoff = 0, empoff = 0;
        size_t s = 0;
        while (s < S.size()) {
            const SamplerConf& sc = S[s];
            int kind;
            if (sc.type == "sampler_RW_tempered" || sc.type == "RW_tempered") kind = 0;
            else if (sc.type == "sampler_RW_block_tempered" || sc.type == "RW_block_tempered") kind = 1;
            else if (sc.type == "sampler_slice_tempered" || sc.type == "slice_tempered") kind = 2;
            else if (sc.type == "sampler_RW_multinomial_tempered" || sc.type == "RW_multinomial_tempered") kind = 3;
            else throw Error("unsupported sampler type '" + sc.type + "' (supported: sampler_RW_tempered, sampler_RW_block_tempered, sampler_slice_tempered, sampler_RW_multinomial_tempered)");
            if (kind == 3 && sc.ctl.useTempering < 0) throw Error("control$useTempering unspecified in sampler_RW_multinomial_tempered");  // APT:1015
            if (sc.ctl.temperPriors < 0 && kind != 2 && kind != 3)  // APT:660, APT:781 (the slice sampler never reads it, APT:909-912)
                throw Error("control$temperPriors unspecified in " + std::string(kind == 0 ? "sampler_RW_tempered" : "sampler_RW_block_tempered"));
            std::vector<int> tgt = expand_target(sc.target);
            if (kind == 0 && tgt.size() > 1) throw Error("cannot use RW sampler on more than one target; try RW_block sampler");  // APT:682
            if (kind == 2 && tgt.size() > 1) throw Error("cannot use slice sampler on more than one target node");  // APT:928
            if (kind == 2 && !(sc.ctl.width > 0)) throw Error("slice sampler: control$width must be positive");
            if (kind == 2 && sc.ctl.sliceMaxSteps < 1) throw Error("slice sampler: control$sliceMaxSteps must be >= 1");
            if (kind == 0 && sc.ctl.log && sc.ctl.reflective)
                throw Error("cannot use reflective RW sampler on a log scale (i.e. with options log=TRUE and reflective=TRUE");  // APT:684
            if (sc.ctl.adaptFactorExponent < 0) throw Error("cannot use RW sampler with adaptFactorExponent control parameter less than 0");  // APT:685
            if (sc.ctl.scale < 0) throw Error("cannot use RW sampler with scale control parameter less than 0");  // APT:686
            if (sc.ctl.adaptInterval < 1) throw Error("adaptInterval must be >= 1");
            Unit u;
            u.kind = kind; u.d = (int)tgt.size(); u.unit0 = unit; u.count = 1; u.tgt = tgt; u.conf = &sc;
            if (kind == 1) {
                if (!sc.propCov.empty()) {
                    if (sc.ctl.propCov_dim != u.d || (int)sc.propCov.size() != u.d * u.d)
                        throw Error("propCov matrix must have dimension " + std::to_string(u.d) + "x" + std::to_string(u.d));  // APT:816
                    for (int i = 0; i < u.d; i++)
                        for (int j = 0; j < i; j++)
                            if (std::fabs(sc.propCov[i + j * u.d] - sc.propCov[j + i * u.d]) > 1e-10 * (std::fabs(sc.propCov[i + j * u.d]) + 1e-300))
                                throw Error("propCov matrix must be symmetric");  // APT:817
                }
                u.blkoff = blkoff; blkoff += 3 * u.d * u.d;
                u.empoff = empoff;
                if (sc.ctl.adaptive && !sc.ctl.adaptScaleOnly) empoff += sc.ctl.adaptInterval * u.d;
            }
            if (kind == 3) {
                // checks of the setup code (APT:1046-1048); state: u + 8 matrices d x d (APT:1029-1040) in the per-rung block state
                int owner = -1;
                long owner_inst = -1;
                std::map<std::string, double> env;
                std::vector<int64_t> els;
                for (int t : tgt)
                    for (int di : stoch_decls) {
                        const CDecl& dd = decls[di];
                        if (!dd.lhs_param) continue;
                        const Var& v = vars.at(dd.lhs->name);
                        if (t < v.theta_off || t >= v.theta_off + v.size) continue;
                        for (long inst = 0; inst < dd.ninst; inst++) {
                            inst_env(dd, inst, env);
                            ref_elements(v, dd.lhs, env, els);
                            for (auto e : els)
                                if (v.theta_off + e == t) {
                                    if (dd.dist != "dmulti") throw Error("can only use RW_multinomial sampler for multinomial distributions");  // APT:1046
                                    if (owner >= 0 && (owner != di || owner_inst != inst)) throw Error("cannot use RW_multinomial sampler on more than one target");  // APT:1047
                                    owner = di; owner_inst = inst;
                                }
                        }
                    }
                if (owner < 0) throw Error("target has no stochastic declaration");
                if ((int)tgt.size() != decls[owner].mvn_n) throw Error("RW_multinomial sampler: the target must be the whole multinomial node");
                if (sc.ctl.adaptive && sc.ctl.adaptInterval < 100) throw Error("adaptInterval < 100 is not recommended for RW_multinomial sampler");  // APT:1048
                u.blkoff = blkoff; blkoff += 1 + 8 * u.d * u.d;
                u.discrete = true;
                u.ntotal = 0;  // Ntotal <- sum(values(model, target)) at setup (APT:1024)
                for (int t : tgt)
                    for (auto& n : param_order) {
                        const Var& v = vars.at(n);
                        if (t >= v.theta_off && t < v.theta_off + v.size) u.ntotal += v.arr->v[(size_t)(t - v.theta_off)];
                    }
                if (!(u.ntotal >= 0 && u.ntotal <= 1000)) throw Error("RW_multinomial sampler: sum of the target must be between 0 and 1000 (rbinom by inversion)");
            }
            if (kind == 2) {  // sumJumps (APT:925) lives in the per-rung sampler block state
                u.blkoff = blkoff; blkoff += 1;
                u.discrete = target_is_discrete(tgt[0]);
            }
            // family detection: a run of scalar samplers on consecutive theta elements with identical control
            size_t e = s + 1;
            std::vector<std::vector<Group>> member_groups;
            if (kind == 0 && !sc.ctl.reflective) {
                while (e < S.size()) {
                    const SamplerConf& nx = S[e];
                    if (nx.type != sc.type || !same_ctl(nx.ctl, sc.ctl)) break;
                    std::vector<int> t2;
                    try { t2 = expand_target(nx.target); } catch (const Error&) { break; }
                    if (t2.size() != 1 || t2[0] != tgt[0] + (int)(e - s)) break;
                    e++;
                }
                if (e - s >= 4) {
                    bool ok = true;
                    for (size_t m = s; m < e && ok; m++) {
                        auto g = dependencies({tgt[0] + (int)(m - s)});
                        for (auto& gg : g)
                            if (gg.full && decls[gg.decl].ninst > 1) ok = false;
                        if (!member_groups.empty()) {
                            if (g.size() != member_groups[0].size()) ok = false;
                            else
                                for (size_t q = 0; q < g.size(); q++)
                                    if (g[q].decl != member_groups[0][q].decl || g[q].prior != member_groups[0][q].prior) ok = false;
                        }
                        member_groups.push_back(std::move(g));
                    }
                    if (ok) {  // disjointness of the members' instance sets, declaration by declaration
                        for (size_t q = 0; q < member_groups[0].size() && ok; q++) {
                            std::vector<char> seen((size_t)decls[member_groups[0][q].decl].ninst, 0);
                            // the same declaration may appear as prior and as dependent: share the marks
                            for (size_t q2 = 0; q2 < member_groups[0].size(); q2++) {
                                if (member_groups[0][q2].decl != member_groups[0][q].decl) continue;
                                for (auto& mg : member_groups)
                                    for (int inst : mg[q2].inst) {
                                        if (seen[(size_t)inst] && q2 == q) ok = false;
                                        if (q2 == q) seen[(size_t)inst] = 1;
                                    }
                            }
                        }
                    }
                    if (!ok) { member_groups.clear(); e = s + 1; }
                } else {
                    e = s + 1;
                }
            }
            if (!member_groups.empty()) {
                u.count = (int)(e - s);
                u.groups = member_groups[0];
                for (size_t q = 0; q < u.groups.size(); q++) {
                    std::vector<int> ptr = {0}, idx;
                    for (auto& mg : member_groups) {
                        for (int inst : mg[q].inst) idx.push_back(inst);
                        ptr.push_back((int)idx.size());
                    }
                    u.groups[q].full = false;
                    u.ia_ptr.push_back(add_iarray(ptr));
                    u.ia_idx.push_back(add_iarray(idx));
                }
            } else {
                u.groups = dependencies(tgt);
                for (auto& g : u.groups) {
                    int id = -1;
                    if (!g.full && g.inst.size() > 1) {
                        bool ap = true;
                        int step = g.inst[1] - g.inst[0];
                        for (size_t i = 2; i < g.inst.size(); i++) ap &= (g.inst[i] - g.inst[i - 1] == step);
                        if (!(ap && decls[g.decl].loops.size() == 1)) id = add_iarray(g.inst);
                    }
                    u.ia_list.push_back(id);
                }
                if (kind == 0 && sc.ctl.reflective) set_bounds(u, tgt[0]);
            }
            if (u.groups.empty()) throw Error("sampler target '" + sc.target + "' has no stochastic dependencies");
            unit += u.count;
            units.push_back(std::move(u));
            s = e;
        }
        NU = unit; BLKSZ = blkoff; EMPSZ = empoff;
    }
    int NU = 0, BLKSZ = 0, EMPSZ = 0;

    // ------------------------------------------------------------------ expression code generation
    using Env = std::map<std::string, std::string>;  // loop variable -> C int expression

    int use_array(const Var& v) {
        Var& mv = vars.at(v.name);
        if (mv.arr_idx < 0) {
            mv.arr_idx = (int)used_arrays.size();
            used_arrays.push_back(mv.arr);
        }
        return mv.arr_idx;
    }
    bool const_indices(const EP& e, long* idx) const {
        for (size_t k = 0; k < e->a.size(); k++) {
            double v;
            if (e->a[k]->k == Expr::RANGE || !try_ceval(e->a[k], {}, v)) return false;
            idx[k] = (long)std::llround(v);
        }
        return true;
    }
    std::string gen_int(const EP& e, const Env& env) {
        double cv;
        if (try_ceval(e, {}, cv)) {
            if (cv != std::floor(cv)) throw Error("non-integer index: " + to_string(e));
            return std::to_string((long)cv);
        }
        switch (e->k) {
            case Expr::VAR:
                if (e->a.empty()) {
                    auto it = env.find(e->name);
                    if (it != env.end()) return it->second;
                    // a scalar sampled node used as an index (demo/APT_demo.R:42-71, rfx[k]): clamped by gen_offset
                    auto pv = vars.find(e->name);
                    if (pv != vars.end() && pv->second.role == Var::PARAM && pv->second.size == 1 && dyn_dim > 0 && !cv_mode)
                        return "apt::dyn_index(th[" + std::to_string(pv->second.theta_off) + "], " + std::to_string(dyn_dim) + ")";
                    throw Error("unknown index variable '" + e->name + "'");
                } else {
                    auto it = vars.find(e->name);
                    if (it == vars.end() || (it->second.role != Var::CONST && it->second.role != Var::DATA))
                        throw Error("index expressions may only use loop variables, constants and scalar sampled nodes: " + to_string(e));
                    return "(int)" + gen_load(it->second, e, env);
                }
            case Expr::BIN:
                if (e->op == '+' || e->op == '-' || e->op == '*')
                    return "(" + gen_int(e->a[0], env) + " " + e->op + " " + gen_int(e->a[1], env) + ")";
                break;
            case Expr::NEG: return "(-" + gen_int(e->a[0], env) + ")";
            case Expr::CALL:
                // ceiling(k) etc. of a sampled node as an index (demo/APT_demo.R:49-53): evaluated in double, clamped into the
                // array like a bare sampled-node index (the function's own range check decides what the value is used for)
                if ((e->name == "ceiling" || e->name == "floor" || e->name == "round" || e->name == "trunc") && e->a.size() == 1 && dyn_dim > 0 && !cv_mode) {
                    const int64_t dd = dyn_dim;
                    const std::string v = gen_dbl(e, env);
                    dyn_dim = dd;
                    return "apt::dyn_index(" + v + ", " + std::to_string(dd) + ")";
                }
                break;
            default: break;
        }
        throw Error("unsupported index expression: " + to_string(e));
    }
    std::string gen_offset(const Var& v, const EP& e, const Env& env) {
        if (e->a.empty()) {
            if (v.size != 1) throw Error("whole-array use of '" + v.name + "' where a scalar is required");
            return "0";
        }
        if (e->a.size() != v.dims.size()) throw Error("wrong number of indices: " + to_string(e));
        long ci[4];
        if (const_indices(e, ci)) {
            int64_t off = 0, stride = 1;
            for (size_t k = 0; k < e->a.size(); k++) {
                if (ci[k] < 1 || ci[k] > v.dims[k]) throw Error("index out of range: " + to_string(e));
                off += (ci[k] - 1) * stride;
                stride *= v.dims[k];
            }
            return std::to_string(off);
        }
        std::string s;
        int64_t stride = 1;
        for (size_t k = 0; k < e->a.size(); k++) {
            if (e->a[k]->k == Expr::RANGE) throw Error("vector slice used where a scalar is required: " + to_string(e));
            dyn_dim = (v.role == Var::CONST || v.role == Var::DATA) ? v.dims[k] : 0;  // only constant arrays may be indexed by a sampled node
            std::string term = "(" + gen_int(e->a[k], env) + " - 1)";
            dyn_dim = 0;
            if (stride != 1) term += " * " + std::to_string(stride);
            s += (k ? " + " : "") + term;
            stride *= v.dims[k];
        }
        return s;
    }
    int64_t dyn_dim = 0;  // extent of the array dimension whose index expression is being generated (0: no sampled-node index allowed)
    // chain-vectorised mode (grid engine): parameters of TB chains live in shared memory as thT[param][chain]
    bool cv_mode = false;
    std::map<const Expr*, std::string> hoisted;  // inprod nodes already emitted as per-(row, chain) arrays
    std::string gen_load(const Var& v, const EP& e, const Env& env) {
        if (v.role == Var::PARAM) {
            std::string off = gen_offset(v, e, env);
            if (cv_mode) return "thT[(" + std::to_string(v.theta_off) + " + " + off + ") * TBS + c_]";
            return "th[" + std::to_string(v.theta_off) + " + " + off + "]";
        }
        if (v.role == Var::CONST && v.size <= 64) {  // small constants are folded into the code
            long ci[4];
            if (const_indices(e, ci)) return fmt_d(ceval(e, {}));
        }
        int ai = use_array(v);
        return "D.a[" + std::to_string(ai) + "][" + gen_offset(v, e, env) + "]";
    }
    std::string gen_slice_elem(const EP& sl, const std::string& ev, const Env& env, int& len) {
        // element `ev` (C int, 0-based) of the slice  v[.., lo:hi, ..]
        if (sl->k != Expr::VAR) throw Error("inprod/sum arguments must be variable slices: " + to_string(sl));
        auto it = vars.find(sl->name);
        if (it == vars.end()) throw Error("unknown variable '" + sl->name + "'");
        const Var& v = it->second;
        if (v.role == Var::DET) throw Error("slices of deterministic nodes inside inprod/sum are not supported: " + to_string(sl));
        auto ref = std::make_shared<Expr>(*sl);
        int rdim = -1;
        if (ref->a.empty()) {
            if (v.dims.size() > 1) throw Error("whole-matrix argument in inprod/sum");
            len = (int)v.size;
            ref->a = {mk_bin('+', mk_var("_sl"), mk_num(1))};
        } else {
            for (size_t k = 0; k < ref->a.size(); k++)
                if (ref->a[k]->k == Expr::RANGE) rdim = (int)k;
            if (rdim < 0) throw Error("expected a slice: " + to_string(sl));
            long lo = (long)std::llround(ceval(ref->a[rdim]->a[0], {})), hi = (long)std::llround(ceval(ref->a[rdim]->a[1], {}));
            len = (int)(hi - lo + 1);
            ref->a[rdim] = mk_bin('+', mk_var("_sl"), mk_num((double)lo));
        }
        Env e2 = env;
        e2["_sl"] = ev;
        return gen_load(v, ref, e2);
    }
    int tmp_id = 0;
    // a condition as a C boolean expression
    std::string gen_cond(const EP& e, const Env& env) {
        if (e->k == Expr::CALL && e->name.rfind("__", 0) == 0) {
            static const std::map<std::string, std::string> c2 = {{"__lt", "<"}, {"__gt", ">"}, {"__le", "<="}, {"__ge", ">="}, {"__eq", "=="}, {"__ne", "!="}};
            auto it = c2.find(e->name);
            if (it != c2.end() && e->a.size() == 2) return "(" + gen_dbl(e->a[0], env) + " " + it->second + " " + gen_dbl(e->a[1], env) + ")";
            if (e->name == "__or" && e->a.size() == 2) return "(" + gen_cond(e->a[0], env) + " || " + gen_cond(e->a[1], env) + ")";
            if (e->name == "__and" && e->a.size() == 2) return "(" + gen_cond(e->a[0], env) + " && " + gen_cond(e->a[1], env) + ")";
            if (e->name == "__not" && e->a.size() == 1) return "(!" + gen_cond(e->a[0], env) + ")";
        }
        return "(" + gen_dbl(e, env) + " != 0.0)";
    }
    std::string gen_dbl(const EP& e, const Env& env) {
        double cv;
        if (try_ceval(e, {}, cv)) return fmt_d(cv);
        switch (e->k) {
            case Expr::NUM: return fmt_d(e->num);
            case Expr::NEG: return "(-" + gen_dbl(e->a[0], env) + ")";
            case Expr::BIN: {
                if (e->op == '^') {
                    double ex;
                    if (try_ceval(e->a[1], {}, ex) && ex == 2.0) return "apt_sq(" + gen_dbl(e->a[0], env) + ")";
                    return "pow(" + gen_dbl(e->a[0], env) + ", " + gen_dbl(e->a[1], env) + ")";
                }
                return "(" + gen_dbl(e->a[0], env) + " " + e->op + " " + gen_dbl(e->a[1], env) + ")";
            }
            case Expr::VAR: {
                if (e->a.empty()) {
                    auto it = env.find(e->name);
                    if (it != env.end()) return "(double)" + it->second;
                }
                auto it = vars.find(e->name);
                if (it == vars.end()) throw Error("unknown variable '" + e->name + "'");
                if (it->second.role == Var::DET) return gen_dbl(inline_det(e), env);
                return gen_load(it->second, e, env);
            }
            case Expr::CALL: {
                static const std::map<std::string, std::string> f1 = {
                    {"exp", "apt_exp"}, {"log", "log"}, {"abs", "fabs"}, {"sqrt", "sqrt"}, {"ilogit", "apt::ilogit"}, {"expit", "apt::ilogit"},
                    {"logit", "apt::logit"}, {"phi", "apt::iprobit"}, {"iprobit", "apt::iprobit"}, {"icloglog", "apt::icloglog"},
                    {"cloglog", "apt::cloglog"}, {"log1p", "log1p"}, {"step", "apt::stepfn"}, {"lgamma", "lgamma"}, {"loggam", "lgamma"}};
                auto it = f1.find(e->name);
                if (it != f1.end()) {
                    if (e->a.size() != 1) throw Error(e->name + "() takes one argument");
                    return it->second + "(" + gen_dbl(e->a[0], env) + ")";
                }
                {
                    static const std::map<std::string, std::string> r1 = {{"ceiling", "ceil"}, {"floor", "floor"}, {"round", "nearbyint"}, {"trunc", "trunc"}};
                    auto ir = r1.find(e->name);
                    if (ir != r1.end() && e->a.size() == 1) return ir->second + "(" + gen_dbl(e->a[0], env) + ")";
                }
                if (e->name == "__if" && e->a.size() == 3)  // lazily evaluated, like the `if (...) return(...)` it came from
                    return "(" + gen_cond(e->a[0], env) + " ? " + gen_dbl(e->a[1], env) + " : " + gen_dbl(e->a[2], env) + ")";
                if (e->name.rfind("__", 0) == 0) return "(" + gen_cond(e, env) + " ? 1.0 : 0.0)";
                if (e->name == "pow" && e->a.size() == 2) return "pow(" + gen_dbl(e->a[0], env) + ", " + gen_dbl(e->a[1], env) + ")";
                if ((e->name == "min" || e->name == "max") && e->a.size() == 2)
                    return "f" + e->name + "(" + gen_dbl(e->a[0], env) + ", " + gen_dbl(e->a[1], env) + ")";
                if (e->name == "inprod" && e->a.size() == 2) {
                    auto hit = hoisted.find(e.get());
                    if (hit != hoisted.end()) return hit->second;
                    int la = 0, lb = 0;
                    std::string ev = "e" + std::to_string(tmp_id++);
                    std::string A = gen_slice_elem(e->a[0], ev, env, la), B = gen_slice_elem(e->a[1], ev, env, lb);
                    if (la != lb) throw Error("inprod arguments have different lengths: " + to_string(e));
                    // chains per tile: a short product is unrolled completely, so that the loads of the data row are common
                    // subexpressions of the CH products eval_multi forms from it (config 5: 1.26e8 -> 1.6e8 updates/s; the
                    // single-chain code, which holds 8 rows per lane in 64 registers, loses by it: 8.8e7 -> 6.6e7)
                    const std::string unroll = (multi > 1 && la <= 16) ? "unroll" : "unroll 5";
                    return "[&]() { double s_ = 0.0;\n_Pragma(\"" + unroll + "\")\n for (int " + ev + " = 0; " + ev + " < " + std::to_string(la) + "; ++" + ev +
                           ") s_ += " + A + " * " + B + "; return s_; }()";
                }
                if (e->name == "sum" && e->a.size() == 1) {
                    int la = 0;
                    std::string ev = "e" + std::to_string(tmp_id++);
                    std::string A = gen_slice_elem(e->a[0], ev, env, la);
                    return "[&]() { double s_ = 0.0; for (int " + ev + " = 0; " + ev + " < " + std::to_string(la) + "; ++" + ev + ") s_ += " + A +
                           "; return s_; }()";
                }
                throw Error("unsupported function '" + e->name + "' in model code");
            }
            default: break;
        }
        throw Error("unsupported expression: " + to_string(e));
    }

    // ------------------------------------------------------------------ emission
    Env decl_env(const CDecl& d) const {
        Env env;
        for (size_t l = 0; l < d.loops.size(); l++) env[d.loops[l].var] = "L" + std::to_string(l);
        return env;
    }
    // call expression of lp_d<di> for linear instance expression `lin` (C int, 0-based)
    std::string lp_call(int di, const std::string& lin) const {
        const CDecl& d = decls[di];
        std::string a[3] = {"0", "0", "0"};
        if (d.loops.size() == 1) a[0] = "(" + std::to_string(d.loops[0].lo) + " + " + lin + ")";
        else if (d.loops.size() >= 2) {
            long n0 = d.loops[0].hi - d.loops[0].lo + 1;
            a[0] = "(" + std::to_string(d.loops[0].lo) + " + (" + lin + ") % " + std::to_string(n0) + ")";
            if (d.loops.size() == 2) a[1] = "(" + std::to_string(d.loops[1].lo) + " + (" + lin + ") / " + std::to_string(n0) + ")";
            else {
                long n1 = d.loops[1].hi - d.loops[1].lo + 1;
                a[1] = "(" + std::to_string(d.loops[1].lo) + " + ((" + lin + ") / " + std::to_string(n0) + ") % " + std::to_string(n1) + ")";
                a[2] = "(" + std::to_string(d.loops[2].lo) + " + (" + lin + ") / " + std::to_string(n0 * n1) + ")";
            }
        }
        return "lp_d" + std::to_string(di) + "(D, th, " + a[0] + ", " + a[1] + ", " + a[2] + ")";
    }

    // 0-based linear instance index of a declaration from the loop values L0, L1, L2 of lp_d
    std::string lin_of_decl(const CDecl& d) const {
        if (d.loops.empty()) return "0";
        std::string r;
        long mul = 1;
        for (size_t l = 0; l < d.loops.size() && l < 3; l++) {
            std::string t = "(L" + std::to_string(l) + " - " + std::to_string(d.loops[l].lo) + ")";
            if (mul != 1) t += " * " + std::to_string(mul);
            r += (l ? " + " : "") + t;
            mul *= d.loops[l].hi - d.loops[l].lo + 1;
        }
        return r;
    }
    bool is_call(const EP& e, const char* name) const { return e->k == Expr::CALL && e->name == name && e->a.size() == 1; }

    // C expression of the log-density of a (non-multivariate) stochastic declaration at value expression `x`
    // derived arrays: per-instance values that depend on data only, filled on the device when data is uploaded
    struct Derived { int decl; std::string kind; long len; };
    std::vector<Derived> derived;
    std::map<int, int> derived_of_decl;
    bool lhs_is_data(const CDecl& d) const {
        auto it = vars.find(d.lhs->name);
        return it != vars.end() && it->second.role == Var::DATA;
    }
    std::string density_expr(const CDecl& d, const std::string& x, const Env& env, const std::string& lin = "", int di = -1) {
        if (d.dist == "dpois" && !lin.empty() && di >= 0 && lhs_is_data(d) && !getenv("APTB200_NO_DERIVED")) {
            // y ~ dpois(lambda) with y data: -0.5 log(2 pi y) and stirlerr(y) come from a derived array
            int k;
            auto f = derived_of_decl.find(di);
            if (f == derived_of_decl.end()) {
                k = (int)derived.size();
                derived.push_back({di, "dpois", 2 * d.ninst});
                derived_of_decl[di] = k;
            } else k = f->second;
            const std::string base = "D.a[NARR + " + std::to_string(k) + "]";
            return "apt::dpois_log_pre(" + x + ", " + gen_dbl(d.iargs[0], env) + ", " + base + "[2 * (" + lin + ")], " + base + "[2 * (" + lin + ") + 1])";
        }
        if (d.dist == "dbinom" && opts.fuse_links && (is_call(d.iargs[0], "ilogit") || is_call(d.iargs[0], "expit"))) {
            double sz;
            if (try_ceval(d.iargs[1], {}, sz) && sz == 1.0)  // fused: dbinom(prob = ilogit(eta), size = 1)
                return "apt::bern_logit_log(" + x + ", " + gen_dbl(d.iargs[0]->a[0], env) + ")";
        }
        if (d.dist == "dnorm") {  // constant sd: pass log(sd) as a literal
            double sd;
            if (try_ceval(d.iargs[1], {}, sd) && sd > 0 && std::isfinite(sd))
                return "apt::dnorm_log_csd(" + x + ", " + gen_dbl(d.iargs[0], env) + ", " + fmt_d(sd) + ", " + fmt_d(std::log(sd)) + ")";
        }
        std::string fn = d.dist == "dnorm" ? "apt::dnorm_log" : d.dist == "dunif" ? "apt::dunif_log" : d.dist == "dgamma" ? "apt::dgamma_log"
                       : d.dist == "dbeta" ? "apt::dbeta_log" : d.dist == "dbinom" ? "apt::dbinom_log" : d.dist == "dpois" ? "apt::dpois_log"
                       : "apt::dexp_log";
        std::string r = fn + "(" + x;
        for (auto& a : d.iargs) r += ", " + gen_dbl(a, env);
        return r + ")";
    }

    void collect_inprods(const EP& e, std::vector<EP>& acc) const {
        if (e->k == Expr::CALL && e->name == "inprod" && e->a.size() == 2) { acc.push_back(e); return; }
        for (auto& x : e->a) collect_inprods(x, acc);
    }
    bool slice_is_param(const EP& sl) const {
        if (sl->k != Expr::VAR) return false;
        auto it = vars.find(sl->name);
        return it != vars.end() && it->second.role == Var::PARAM;
    }

    // Grid engine: evaluate R consecutive instances of a big declaration for TB chains at once.  Inner products of
    // a parameter slice with a data row are interchanged so that each data element is loaded once and reused by
    // all chains (the loop over the TB chains is innermost and fully unrolled into registers).
    // tile mode of the grid engine (see emit_big_eval_fn): lanes of a warp = (row group, chain group)
    bool tile_want = false, unit_tile = false, tile_mma = false;
    std::string mma_stage_fn, mma_y_hoisted;
    int frag_doubles = 0;
    int tile_cs = 1, tile_rt = 2, tile_rows = 64, tile_tb = 1;
    void emit_big_eval(Unit& u) {
        std::vector<size_t> bigs;
        for (size_t g = 0; g < u.groups.size(); g++)
            if (u.groups[g].big) bigs.push_back(g);
        out << "    static constexpr int NBIG = " << bigs.size() << ";\n";
        unit_tile = false;
        mma_stage_fn.clear(); mma_y_hoisted.clear();
        if (!bigs.empty()) {
            emit_big_eval_fn(u, bigs, "big_eval", true);
            if (unit_tile) emit_big_eval_fn(u, bigs, "big_eval_rem", false);
        } else emit_big_eval_fn(u, bigs, "big_eval", true);
        out << "    static constexpr bool TILE = " << (unit_tile ? "true" : "false") << ", MMA = " << (mma_stage_fn.empty() ? "false" : "true") << ";\n";
        if (mma_stage_fn.empty()) out << "    __device__ static __forceinline__ void mma_stage(const double*, double*, int) {}\n";
        else out << mma_stage_fn;
        out << "    static constexpr int CS = " << (unit_tile ? tile_cs : 1) << ", RT = " << (unit_tile ? tile_rt : 1) << ", TROWS = " << (unit_tile ? tile_rows : 1) << ";\n";
    }
    void emit_big_eval_fn(Unit& u, const std::vector<size_t>& bigs, const std::string& fname, bool allow_ring) {
        const bool primary = fname == "big_eval";
        if (bigs.empty()) {
            out << "    __device__ static constexpr int big_group(int gi) { return 0; }\n";
            out << "    __host__ __device__ static constexpr long big_n(int gi) { return 0; }\n";
            out << "    template <int GI, int TB, int R> __device__ static __forceinline__ void big_eval(const Data& D, const double* __restrict__ thT, int i0, int i_next, double (&acc)[TB], double* xbuf, int& ring_pos, int tid_) {}\n";
            out << "    template <int GI> __device__ static __forceinline__ void ring_prime(const Data& D, double* xbuf, int tid_, int i0) {}\n";
            out << "    static constexpr bool RING = false;\n";
            return;
        }
        if (primary) {
        out << "    __device__ static constexpr int big_group(int gi) { return ";
        for (size_t i = 0; i + 1 < bigs.size(); i++) out << "gi == " << i << " ? " << bigs[i] << " : ";
        out << bigs.back() << "; }\n";
        out << "    __host__ __device__ static constexpr long big_n(int gi) { return ";
        for (size_t i = 0; i + 1 < bigs.size(); i++) out << "gi == " << i << " ? " << decls[u.groups[bigs[i]].decl].ninst << "L : ";
        out << decls[u.groups[bigs.back()].decl].ninst << "L; }\n";
        }
        out << "    template <int GI, int TB, int R> __device__ static __forceinline__ void " << fname << "(const Data& D, const double* __restrict__ thT, int i0, int i_next, double (&acc)[TB], double* xbuf, int& ring_pos, int tid_) {\n";
        out << "      constexpr int TBS = Model::TB;  // chain stride of thT (a thread may own only TB of the TBS chains)\n";
        std::ostringstream prime;  // ring_prime<GI> bodies
        bool any_ring = false;
        for (size_t bi = 0; bi < bigs.size(); bi++) {
            CDecl& d = decls[u.groups[bigs[bi]].decl];
            if (d.loops.size() != 1 || d.dist == "dmnorm") throw Error("grid engine: big declarations must be univariate with exactly one loop");
            out << "      if constexpr (GI == " << bi << ") {  // " << to_string(d.lhs) << " ~ " << d.dist << ", " << d.ninst << " instances\n";
            Env env;
            env[d.loops[0].var] = "(" + std::to_string(d.loops[0].lo) + " + i0 + r_)";
            std::vector<EP> ips;
            for (auto& a : d.iargs) collect_inprods(a, ips);
            cv_mode = true;
            hoisted.clear();
            int nh = 0;
            for (auto& ip : ips) {
                const bool pa = slice_is_param(ip->a[0]), pb = slice_is_param(ip->a[1]);
                if (pa == pb) continue;  // not (parameter slice x data slice): evaluated per chain in place
                const EP& ps = pa ? ip->a[0] : ip->a[1];
                const EP& ds = pa ? ip->a[1] : ip->a[0];
                int lp_ = 0, ld_ = 0;
                std::string pe = gen_slice_elem(ps, "e_", env, lp_);
                Env env0 = env, env1 = env;
                std::string de = gen_slice_elem(ds, "e_", env, ld_);
                if (lp_ != ld_) throw Error("inprod arguments have different lengths: " + to_string(ip));
                // vector-load eligibility: data slice is X[loopvar, lo:hi] with an even leading dimension
                bool vec = false;
                auto dv = vars.find(ds->name);
                if (dv != vars.end() && ds->a.size() == 2 && ds->a[1]->k == Expr::RANGE && ds->a[0]->k == Expr::VAR && ds->a[0]->a.empty() &&
                    ds->a[0]->name == d.loops[0].var && dv->second.dims[0] % 2 == 0 && d.loops[0].lo % 2 == 1)
                    vec = true;
                std::string nm = "ip" + std::to_string(nh++) + "_";
                const int BW = 5, RD = 5, RS = 6;
                const bool ring = allow_ring && vec && !any_ring && bigs.size() == 1 && lp_ % BW == 0 && lp_ / BW >= RD && ring_ok;
                if (allow_ring && vec && !any_ring && bigs.size() == 1 && lp_ >= 12 && ring_ok && tile_want && tile_mma && ips.size() == 1) {
                    // MMA tile mode (8 or 16 rungs per pass).  A warp owns 32 consecutive rows; the product of its 32 x lp_ block of
                    // the design matrix with the lp_ x TBS parameter block runs on the FP64 tensor-core path (mma.sync.m8n8k4.f64):
                    // per block of 4 columns, 4 row blocks x TBS/8 chain blocks, A fragments (one data value per lane and row
                    // block) from the warp's cp.async ring, B fragments (one parameter per lane and chain block) from the staged
                    // parameters.  DMMA shares the FP64 pipe with DFMA on sm_100a (tools/ub_dmma.cu), so the arithmetic floor is
                    // unchanged; the product costs 104 issue slots + 78 LDS.64 per tile instead of 800 DFMA + 200 LDS.128
                    // (tools/ub_dmma_stream.cu: streaming 127 -> 107 us per pass).  The accumulator fragment gives lane (rq, kq)
                    // rows rq + 8 r and chain positions kq * TBT + {0 .. TBT-1}: the same 4 x TBT thread tile as before, so the
                    // Bernoulli-logit terms and the reductions are unchanged.  Columns beyond lp_ (last block) enter as zeros.
                    any_ring = true;
                    unit_tile = true;
                    const int KB = 4, XS = 36;  // columns per block; padded column stride (doubles): LDS.64 of a fragment hits 16 distinct bank pairs
                    const int NBR = (lp_ + KB - 1) / KB, NT = tile_tb / 8;
                    int TRD = std::min(NBR, 6);  // blocks in flight per warp (measured at 16 rungs, ms per pass: 3: 0.1331, 4: 0.1309, 5: 0.1301, 6: 0.1299, 7: 0.1297)
                    if (const char* e = getenv("APTB200_TILE_DEPTH")) TRD = std::max(1, std::min(NBR, atoi(e)));
                    const int TRS = TRD + 1, BLK = KB * XS;
                    auto xaddr = [&](const std::string& row, const std::string& col) {
                        Env e2 = env;
                        e2[d.loops[0].var] = "(" + std::to_string(d.loops[0].lo) + " + " + row + ")";
                        int tmp;
                        return gen_slice_elem(ds, col, e2, tmp);
                    };
                    auto fill = [&](std::ostream& o, const std::string& ind, const std::string& slot, const std::string& row0, const std::string& blk) {
                        o << ind << "_Pragma(\"unroll\") for (int j_ = 0; j_ < 2; ++j_) { const int col_ = (lane_ >> 4) + 2 * j_; if (" << blk << " * " << KB << " + col_ < " << lp_
                          << ") apt::cp_async16(&xw_[" << slot << " * " << BLK << " + col_ * " << XS << " + 2 * (lane_ & 15)], &" << xaddr(row0 + " + 2 * (lane_ & 15)", "(" + blk + " * " + std::to_string(KB) + " + col_)") << "); }\n";
                    };
                    out << "        static_assert(R == 4 && TB == " << 2 * NT << ", \"MMA tile: 4 rows x TBS/4 chains per thread\");\n";
                    out << "        const int lane_ = tid_ & 31, kq_ = lane_ & 3, rq_ = lane_ >> 2;\n";
                    out << "        double* xw_ = xbuf + (tid_ >> 5) * " << TRS * BLK << ";\n";
                    out << "        const double* thF_ = xbuf - Model::FRAG_DOUBLES + lane_ * " << NT << ";  // B fragments of this lane, see mma_stage\n";
                    {   // the responses of the thread's rows are on their way while the product runs
                        Env ey = env;
                        ey[d.loops[0].var] = "(" + std::to_string(d.loops[0].lo) + " + i0 + r_ * 8 + rq_)";
                        out << "        double yh_[R];\n        _Pragma(\"unroll\") for (int r_ = 0; r_ < R; ++r_) yh_[r_] = " << gen_dbl(d.lhs, ey) << ";\n";
                        mma_y_hoisted = gen_dbl(d.lhs, ey);
                    }
                    out << "        double " << nm << "[R][TB];\n";
                    out << "        _Pragma(\"unroll\") for (int r_ = 0; r_ < R; ++r_) { _Pragma(\"unroll\") for (int c_ = 0; c_ < TB; ++c_) " << nm << "[r_][c_] = 0.0; }\n";
                    out << "        _Pragma(\"unroll\") for (int b_ = 0; b_ < " << NBR << "; ++b_) {\n";
                    out << "          apt::cp_async_wait<" << TRD - 1 << ">();\n          __syncwarp();\n";
                    out << "          { const int nb_ = b_ + " << TRD << "; int slot_ = ring_pos + " << TRD << "; if (slot_ >= " << TRS << ") slot_ -= " << TRS << ";\n";
                    out << "            if (nb_ < " << NBR << ") {\n"; fill(out, "              ", "slot_", "i0", "nb_"); out << "            }\n";
                    out << "            else if (i_next >= 0) {\n"; fill(out, "              ", "slot_", "i_next", "(nb_ - " + std::to_string(NBR) + ")"); out << "            }\n";
                    out << "            apt::cp_async_commit(); }\n";
                    {
                        Env e3 = env;
                        int tmp;
                        cv_mode = true;
                        std::string pe2 = gen_slice_elem(ps, "(b_ * " + std::to_string(KB) + " + kq_)", e3, tmp);
                        // A last block of only 1 or 2 columns would spend a whole DMMA step (16 pipe cycles per instruction) on
                        // mostly zeros: those columns go through 4 x TBT plain FMAs per column instead (half the pipe time for 2
                        // columns, a quarter for 1), on the same accumulators.
                        const int remc = lp_ % KB;
                        const bool rem_fma = (remc == 1 || remc == 2) && !getenv("APTB200_MMA_PAD");
                        if (rem_fma) {
                            std::string per = gen_slice_elem(ps, "(" + std::to_string((NBR - 1) * KB) + " + e_)", e3, tmp);
                            out << "          if (b_ == " << NBR - 1 << ") {\n";
                            out << "            _Pragma(\"unroll\") for (int e_ = 0; e_ < " << remc << "; ++e_) {\n";
                            out << "              double xr_[R];\n";
                            out << "              _Pragma(\"unroll\") for (int r_ = 0; r_ < R; ++r_) xr_[r_] = xw_[ring_pos * " << BLK << " + e_ * " << XS << " + rq_ + 8 * r_];\n";
                            out << "              _Pragma(\"unroll\") for (int c_ = 0; c_ < TB; c_ += 2) { const double2 b2_ = *reinterpret_cast<const double2*>(&" << per << ");\n";
                            out << "                _Pragma(\"unroll\") for (int r_ = 0; r_ < R; ++r_) { " << nm << "[r_][c_] = fma(xr_[r_], b2_.x, " << nm << "[r_][c_]); " << nm << "[r_][c_ + 1] = fma(xr_[r_], b2_.y, " << nm << "[r_][c_ + 1]); } }\n";
                            out << "            }\n          } else {\n";
                        } else out << "          {\n";
                        out << "          const bool live_ = b_ * " << KB << " + kq_ < " << lp_ << ";\n";
                        out << "          double bf_[" << NT << "];\n";
                        if (NT == 2) out << "          { const double2 f2_ = *reinterpret_cast<const double2*>(thF_ + b_ * 64); bf_[0] = f2_.x; bf_[1] = f2_.y; }\n";
                        else out << "          bf_[0] = thF_[b_ * 32];\n";
                        // the staged parameters in fragment order: [block][lane][chain block], zeros beyond the last column (one
                        // conflict-free LDS per block; read straight from the [parameter][position] staging, the four lanes of a
                        // row group collide on the same banks)
                        std::ostringstream st;
                        st << "    __device__ static __forceinline__ void mma_stage(const double* thT, double* thF, int tid_) {\n";
                        st << "      constexpr int TBS = Model::TB, TB = " << 2 * NT << ";\n";
                        st << "      for (int i_ = tid_; i_ < " << NBR * 32 * NT << "; i_ += Model::GRID_THREADS) {\n";
                        st << "        const int t_ = i_ % " << NT << ", lane_ = (i_ / " << NT << ") % 32, b_ = i_ / " << 32 * NT << ", kq_ = lane_ & 3, rq_ = lane_ >> 2;\n";
                        st << "        const int c_ = (rq_ >> 1) * TB + (rq_ & 1) + 2 * t_;\n";
                        st << "        thF[i_] = (b_ * " << KB << " + kq_ < " << lp_ << ") ? " << pe2 << " : 0.0;\n      }\n    }\n";
                        mma_stage_fn = st.str();
                        frag_doubles = std::max(frag_doubles, NBR * 32 * NT);
                        out << "          const double* xs_ = xw_ + ring_pos * " << BLK << " + kq_ * " << XS << " + rq_;\n";
                        out << "          _Pragma(\"unroll\") for (int r_ = 0; r_ < R; ++r_) {\n";
                        out << "            double a_ = xs_[8 * r_];\n";
                        if (lp_ % KB != 0) out << "            if (b_ == " << NBR - 1 << ") a_ = live_ ? a_ : 0.0;  // never multiply stale shared memory\n";
                        out << "            _Pragma(\"unroll\") for (int t_ = 0; t_ < " << NT << "; ++t_) apt::dmma884(" << nm << "[r_][2 * t_], " << nm << "[r_][2 * t_ + 1], a_, bf_[t_]);\n          }\n";
                        out << "          }\n";
                    }
                    out << "          ring_pos = (ring_pos + 1 == " << TRS << ") ? 0 : ring_pos + 1;\n        }\n";
                    prime << "      if constexpr (GI == " << bi << ") {\n        const int lane_ = tid_ & 31;\n        double* xw_ = xbuf + (tid_ >> 5) * " << TRS * BLK << ";\n"
                          << "        _Pragma(\"unroll\") for (int b_ = 0; b_ < " << TRD << "; ++b_) {\n";
                    fill(prime, "          ", "b_", "i0", "b_");
                    prime << "          apt::cp_async_commit();\n        }\n      }\n";
                    ring_slots = std::max(ring_slots, (TRS * BLK + 31) / 32);
                    hoisted[ip.get()] = nm + "[r_][c_]";
                    env[d.loops[0].var] = "(" + std::to_string(d.loops[0].lo) + " + i0 + r_ * 8 + rq_)";
                    continue;
                }
                if (allow_ring && vec && !any_ring && bigs.size() == 1 && lp_ % BW == 0 && lp_ / BW >= 3 && ring_ok && tile_want && ips.size() == 1) {
                    // Tile mode.  A warp owns TR consecutive rows; lane = (row group rg, chain group cg): the thread
                    // multiplies RT rows (RT/2 adjacent pairs, one LDS.128 each) by TBT chains, so that one broadcast
                    // load of two parameters feeds 2 RT FMAs and one load of two data values 2 TBT FMAs.  (With one
                    // row x 16 chains per thread the pass was bound by the shared-memory pipe, not by the FP64 pipe:
                    // 9 LDS per 16 DFMA; see profiles/.)  The warp streams its rows through a private cp.async ring
                    // of RS slots of BW columns, 16 bytes per lane and column.
                    any_ring = true;
                    unit_tile = true;
                    const int NBR = lp_ / BW, TR = tile_rows, RG = 32 / tile_cs;
                    int TRD = 3;  // blocks in flight; ring slots = TRD + 1 (deeper rings cost occupancy)
                    if (const char* e = getenv("APTB200_TILE_DEPTH")) TRD = std::max(1, std::min(NBR, atoi(e)));
                    const int TRS = TRD + 1;
                    int unr = (NBR % 2 == 0 ? 2 : 1);
                    if (const char* e = getenv("APTB200_TILE_UNROLL")) unr = std::max(1, atoi(e));
                    auto xaddr = [&](const std::string& row, const std::string& col) {
                        Env e2 = env;
                        e2[d.loops[0].var] = "(" + std::to_string(d.loops[0].lo) + " + " + row + ")";
                        int tmp;
                        return gen_slice_elem(ds, col, e2, tmp);
                    };
                    out << "        const int lane_ = tid_ & 31, rg2_ = ((tid_ & 31) / " << tile_cs << ") * 2;\n";
                    out << "        double* xw_ = xbuf + (tid_ >> 5) * " << TRS * BW * TR << ";\n";
                    out << "        double " << nm << "[R][TB];\n";
                    out << "        _Pragma(\"unroll\") for (int r_ = 0; r_ < R; ++r_) { _Pragma(\"unroll\") for (int c_ = 0; c_ < TB; ++c_) " << nm << "[r_][c_] = 0.0; }\n";
                    out << "        _Pragma(\"unroll " << unr << "\") for (int b_ = 0; b_ < " << NBR << "; ++b_) {  // not fully unrolled: the loop body must stay inside the instruction cache\n";
                    out << "          apt::cp_async_wait<" << TRD - 1 << ">();\n          __syncwarp();\n";
                    out << "          { const int nb_ = b_ + " << TRD << "; int slot_ = ring_pos + " << TRD << "; if (slot_ >= " << TRS << ") slot_ -= " << TRS << ";\n";
                    out << "            if (nb_ < " << NBR << ") { _Pragma(\"unroll\") for (int e_ = 0; e_ < " << BW << "; ++e_) for (int q_ = lane_; q_ < " << TR / 2
                        << "; q_ += 32) apt::cp_async16(&xw_[(slot_ * " << BW << " + e_) * " << TR << " + 2 * q_], &" << xaddr("i0 + 2 * q_", "(nb_ * " + std::to_string(BW) + " + e_)") << "); }\n";
                    out << "            else if (i_next >= 0) { _Pragma(\"unroll\") for (int e_ = 0; e_ < " << BW << "; ++e_) for (int q_ = lane_; q_ < " << TR / 2
                        << "; q_ += 32) apt::cp_async16(&xw_[(slot_ * " << BW << " + e_) * " << TR << " + 2 * q_], &" << xaddr("i_next + 2 * q_", "((nb_ - " + std::to_string(NBR) + ") * " + std::to_string(BW) + " + e_)") << "); }\n";
                    out << "            apt::cp_async_commit(); }\n";
                    {
                        Env e3 = env;
                        int tmp;
                        cv_mode = true;
                        std::string pe2 = gen_slice_elem(ps, "(b_ * " + std::to_string(BW) + " + e_)", e3, tmp);
                        out << "          _Pragma(\"unroll\") for (int e_ = 0; e_ < " << BW << "; ++e_) {\n";
                        out << "            double xr_[R];\n";
                        out << "            _Pragma(\"unroll\") for (int h_ = 0; h_ < R / 2; ++h_) { const double2 x2_ = *reinterpret_cast<const double2*>(&xw_[(ring_pos * " << BW << " + e_) * " << TR
                            << " + h_ * " << 2 * RG << " + rg2_]); xr_[2 * h_] = x2_.x; xr_[2 * h_ + 1] = x2_.y; }\n";
                        out << "            _Pragma(\"unroll\") for (int c_ = 0; c_ < TB; c_ += 2) { const double2 b2_ = *reinterpret_cast<const double2*>(&" << pe2 << ");\n";
                        out << "              _Pragma(\"unroll\") for (int r_ = 0; r_ < R; ++r_) { " << nm << "[r_][c_] = fma(xr_[r_], b2_.x, " << nm << "[r_][c_]); " << nm << "[r_][c_ + 1] = fma(xr_[r_], b2_.y, " << nm << "[r_][c_ + 1]); } }\n          }\n";
                    }
                    out << "          ring_pos = (ring_pos + 1 == " << TRS << ") ? 0 : ring_pos + 1;\n        }\n";
                    prime << "      if constexpr (GI == " << bi << ") {\n        const int lane_ = tid_ & 31;\n        double* xw_ = xbuf + (tid_ >> 5) * " << TRS * BW * TR << ";\n"
                          << "        _Pragma(\"unroll\") for (int b_ = 0; b_ < " << TRD << "; ++b_) {\n"
                          << "          _Pragma(\"unroll\") for (int e_ = 0; e_ < " << BW << "; ++e_) for (int q_ = lane_; q_ < " << TR / 2 << "; q_ += 32) apt::cp_async16(&xw_[(b_ * " << BW << " + e_) * " << TR
                          << " + 2 * q_], &" << xaddr("i0 + 2 * q_", "(b_ * " + std::to_string(BW) + " + e_)") << ");\n          apt::cp_async_commit();\n        }\n      }\n";
                    ring_slots = std::max(ring_slots, TRS * BW * TR / 32);
                    hoisted[ip.get()] = nm + "[r_][c_]";
                    // rows of this thread: pair h_ = r_ / 2 of row group rg
                    env[d.loops[0].var] = "(" + std::to_string(d.loops[0].lo) + " + i0 + (r_ >> 1) * " + std::to_string(2 * RG) + " + rg2_ + (r_ & 1))";
                    continue;
                }
                if (ring) {
                    // cp.async ring: every thread streams its own rows through a private shared-memory ring, RD blocks
                    // of BW columns ahead (half a row), so the HBM latency never reaches the FMA chain
                    any_ring = true;
                    const int NBR = lp_ / BW;
                    auto xaddr = [&](const std::string& row, const std::string& col) {
                        Env e2 = env;
                        e2[d.loops[0].var] = "(" + std::to_string(d.loops[0].lo) + " + " + row + ")";
                        int tmp;
                        return gen_slice_elem(ds, col, e2, tmp);
                    };
                    out << "        double " << nm << "[R][TB];\n";
                    out << "        _Pragma(\"unroll\") for (int c_ = 0; c_ < TB; ++c_) " << nm << "[0][c_] = 0.0;\n";
                    out << "        _Pragma(\"unroll\") for (int b_ = 0; b_ < " << NBR << "; ++b_) {\n";
                    out << "          apt::cp_async_wait<" << RD - 1 << ">();\n          double xr_[" << BW << "];\n";
                    out << "          _Pragma(\"unroll\") for (int e_ = 0; e_ < " << BW << "; ++e_) xr_[e_] = xbuf[(ring_pos * " << BW << " + e_) * GRID_THREADS + tid_];\n";
                    out << "          { const int nb_ = b_ + " << RD << "; int slot_ = ring_pos + " << RD << "; if (slot_ >= " << RS << ") slot_ -= " << RS << ";\n";
                    out << "            if (nb_ < " << NBR << ") { _Pragma(\"unroll\") for (int e_ = 0; e_ < " << BW << "; ++e_) apt::cp_async8(&xbuf[(slot_ * " << BW
                        << " + e_) * GRID_THREADS + tid_], &" << xaddr("i0", "(nb_ * " + std::to_string(BW) + " + e_)") << "); }\n";
                    out << "            else if (i_next >= 0) { _Pragma(\"unroll\") for (int e_ = 0; e_ < " << BW << "; ++e_) apt::cp_async8(&xbuf[(slot_ * " << BW
                        << " + e_) * GRID_THREADS + tid_], &" << xaddr("i_next", "((nb_ - " + std::to_string(NBR) + ") * " + std::to_string(BW) + " + e_)") << "); }\n";
                    out << "            apt::cp_async_commit(); }\n";
                    {
                        Env e3 = env;
                        int tmp;
                        cv_mode = true;
                        std::string pe2 = gen_slice_elem(ps, "(b_ * " + std::to_string(BW) + " + e_)", e3, tmp);
                        out << "          _Pragma(\"unroll\") for (int e_ = 0; e_ < " << BW << "; ++e_) {\n";
                        out << "            _Pragma(\"unroll\") for (int c_ = 0; c_ < TB; c_ += 2) { const double2 b2_ = *reinterpret_cast<const double2*>(&" << pe2 << "); "
                            << nm << "[0][c_] = fma(xr_[e_], b2_.x, " << nm << "[0][c_]); " << nm << "[0][c_ + 1] = fma(xr_[e_], b2_.y, " << nm << "[0][c_ + 1]); }\n          }\n";
                    }
                    out << "          ring_pos = (ring_pos + 1 == " << RS << ") ? 0 : ring_pos + 1;\n        }\n";
                    prime << "      if constexpr (GI == " << bi << ") {\n        _Pragma(\"unroll\") for (int b_ = 0; b_ < " << RD << "; ++b_) {\n"
                          << "          _Pragma(\"unroll\") for (int e_ = 0; e_ < " << BW << "; ++e_) apt::cp_async8(&xbuf[(b_ * " << BW << " + e_) * GRID_THREADS + tid_], &"
                          << xaddr("i0", "(b_ * " + std::to_string(BW) + " + e_)") << ");\n          apt::cp_async_commit();\n        }\n      }\n";
                    ring_slots = RS * BW;
                    hoisted[ip.get()] = nm + "[r_][c_]";
                    continue;
                }
                {   // the rows this thread evaluates next are pulled towards the SM while the current ones are computed
                    Env envn = env;
                    envn[d.loops[0].var] = "(" + std::to_string(d.loops[0].lo) + " + i_next)";
                    int tmp;
                    std::string dn = gen_slice_elem(ds, "e_", envn, tmp);
                    out << "        if (i_next >= 0) { _Pragma(\"unroll\") for (int e_ = 0; e_ < " << lp_ << "; ++e_) apt::prefetch_data(&" << dn << "); }\n";
                }
                out << "        double " << nm << "[R][TB];\n";
                out << "        _Pragma(\"unroll\") for (int r_ = 0; r_ < R; ++r_) { _Pragma(\"unroll\") for (int c_ = 0; c_ < TB; ++c_) " << nm << "[r_][c_] = 0.0; }\n";
                out << "        _Pragma(\"unroll 10\") for (int e_ = 0; e_ < " << lp_ << "; ++e_) {\n";
                out << "          double xr_[R];\n";
                if (vec) {
                    Env envv = env;
                    envv[d.loops[0].var] = "(" + std::to_string(d.loops[0].lo) + " + i0)";
                    int tmp;
                    std::string de0 = gen_slice_elem(ds, "e_", envv, tmp);
                    out << "          if constexpr (R == 2) { const double2 v2_ = __ldcs(reinterpret_cast<const double2*>(&" << de0 << ")); xr_[0] = v2_.x; xr_[1] = v2_.y; }\n";
                    out << "          else { _Pragma(\"unroll\") for (int r_ = 0; r_ < R; ++r_) xr_[r_] = __ldcs(&" << de << "); }\n";
                } else {
                    out << "          _Pragma(\"unroll\") for (int r_ = 0; r_ < R; ++r_) xr_[r_] = __ldcs(&" << de << ");\n";
                }
                // parameters of two chains per shared-memory load (thT rows are 16-byte aligned when TB is even)
                out << "          if constexpr (TB % 2 == 0) {\n            _Pragma(\"unroll\") for (int c_ = 0; c_ < TB; c_ += 2) { const double2 b2_ = *reinterpret_cast<const double2*>(&" << pe
                    << "); _Pragma(\"unroll\") for (int r_ = 0; r_ < R; ++r_) { " << nm << "[r_][c_] = fma(xr_[r_], b2_.x, " << nm << "[r_][c_]); " << nm
                    << "[r_][c_ + 1] = fma(xr_[r_], b2_.y, " << nm << "[r_][c_ + 1]); } }\n          } else {\n";
                out << "            _Pragma(\"unroll\") for (int c_ = 0; c_ < TB; ++c_) { const double b_ = " << pe << "; _Pragma(\"unroll\") for (int r_ = 0; r_ < R; ++r_) "
                    << nm << "[r_][c_] = fma(xr_[r_], b_, " << nm << "[r_][c_]); }\n          }\n";
                out << "        }\n";
                hoisted[ip.get()] = nm + "[r_][c_]";
            }
            double sz_;
            if (unit_tile && primary && d.dist == "dbinom" && opts.fuse_links && (is_call(d.iargs[0], "ilogit") || is_call(d.iargs[0], "expit")) && try_ceval(d.iargs[1], {}, sz_) && sz_ == 1.0) {
                // fused Bernoulli-logit terms of the whole thread tile (pairs of rows in lock step)
                out << "        double y_[R], eta_[R][TB];\n";
                out << "        _Pragma(\"unroll\") for (int r_ = 0; r_ < R; ++r_) { y_[r_] = " << ((!mma_y_hoisted.empty() && mma_y_hoisted == gen_dbl(d.lhs, env)) ? std::string("yh_[r_]") : gen_dbl(d.lhs, env)) << "; _Pragma(\"unroll\") for (int c_ = 0; c_ < TB; ++c_) eta_[r_][c_] = "
                    << gen_dbl(d.iargs[0]->a[0], env) << "; }\n";
                out << "        apt::bern_logit_tile<R, TB>(y_, eta_, acc);\n      }\n";
                cv_mode = false;
                hoisted.clear();
                continue;
            }
            out << "        _Pragma(\"unroll\") for (int r_ = 0; r_ < R; ++r_) {\n";
            out << "          const double x_ = " << gen_dbl(d.lhs, env) << ";\n";
            if (d.dist == "dbinom" && opts.fuse_links && (is_call(d.iargs[0], "ilogit") || is_call(d.iargs[0], "expit")) && try_ceval(d.iargs[1], {}, sz_) && sz_ == 1.0) {
                // fused Bernoulli-logit terms of all chains evaluated in lock step (explicit ILP across chains)
                out << "          double eta_[TB];\n          _Pragma(\"unroll\") for (int c_ = 0; c_ < TB; ++c_) eta_[c_] = " << gen_dbl(d.iargs[0]->a[0], env) << ";\n";
                out << "          apt::bern_logit_acc<TB>(x_, eta_, acc);\n";
            } else {
                out << "          _Pragma(\"unroll\") for (int c_ = 0; c_ < TB; ++c_) acc[c_] += " << density_expr(d, "x_", env) << ";\n";
            }
            out << "        }\n      }\n";
            cv_mode = false;
            hoisted.clear();
        }
        out << "    }\n";
        if (primary) {
            out << "    template <int GI> __device__ static __forceinline__ void ring_prime(const Data& D, double* xbuf, int tid_, int i0) {\n" << prime.str() << "    }\n";
            out << "    static constexpr bool RING = " << (any_ring ? "true" : "false") << ";\n";
        }
    }
    int ring_slots = 0;
    bool ring_ok = true;

    void emit_lp_functions() {
        for (int di : stoch_decls) {
            CDecl& d = decls[di];
            Env env = decl_env(d);
            out << "  // line " << d.line << ": " << to_string(d.lhs) << " ~ " << d.dist << "(...)";
            for (auto& l : d.loops) out << "  for " << l.var << " in " << l.lo << ":" << l.hi;
            out << "\n";
            out << "  __device__ static __forceinline__ double lp_d" << di << "(const Data& D, const double* th, int L0, int L1, int L2) {\n";
            if (d.dist == "dmnorm") {
                const int n = d.mvn_n;
                int rdim = -1;
                for (size_t k = 0; k < d.lhs->a.size(); k++)
                    if (d.lhs->a[k]->k == Expr::RANGE) rdim = (int)k;
                long lo = (long)std::llround(ceval(d.lhs->a[rdim]->a[0], {}));
                out << "    const double x_[" << n << "] = {";
                for (int e = 0; e < n; e++) {
                    auto ref = std::make_shared<Expr>(*d.lhs);
                    ref->a[rdim] = mk_num((double)(lo + e));
                    out << (e ? ", " : "") << gen_dbl(ref, env);
                }
                out << "};\n    const double mu_[" << n << "] = {";
                for (int e = 0; e < n; e++) out << (e ? ", " : "") << gen_dbl(d.mean_elems[e], env);
                out << "};\n    const double U_[" << n * n << "] = {";
                for (int e = 0; e < n * n; e++) out << (e ? ", " : "") << gen_dbl(d.chol_elems[e], env);
                out << "};\n    return apt::dmnorm_chol_log<" << n << ">(x_, mu_, U_, (" << gen_dbl(d.iargs[0], env) << ") != 0.0);\n";
            } else if (d.dist == "dmulti") {
                const int n = d.mvn_n;
                int rdim = -1;
                for (size_t k = 0; k < d.lhs->a.size(); k++)
                    if (d.lhs->a[k]->k == Expr::RANGE) rdim = (int)k;
                long lo = (long)std::llround(ceval(d.lhs->a[rdim]->a[0], {}));
                out << "    const double x_[" << n << "] = {";
                for (int e = 0; e < n; e++) {
                    auto ref = std::make_shared<Expr>(*d.lhs);
                    ref->a[rdim] = mk_num((double)(lo + e));
                    out << (e ? ", " : "") << gen_dbl(ref, env);
                }
                out << "};\n    const double p_[" << n << "] = {";
                for (int e = 0; e < n; e++) out << (e ? ", " : "") << gen_dbl(d.mean_elems[e], env);
                out << "};\n    return apt::dmulti_log<" << n << ">(x_, p_, " << gen_dbl(d.iargs[0], env) << ");\n";
            } else if (d.dist == "dcat") {
                out << "    const double p_[" << d.mvn_n << "] = {";
                for (int e = 0; e < d.mvn_n; e++) out << (e ? ", " : "") << gen_dbl(d.mean_elems[e], env);
                out << "};\n    return apt::dcat_log<" << d.mvn_n << ">(" << gen_dbl(d.lhs, env) << ", p_);\n";
            } else {
                out << "    return " << density_expr(d, gen_dbl(d.lhs, env), env, lin_of_decl(d), di) << ";\n";
                if (is_fused_bern(d)) {
                    out << "  }\n  __device__ static __forceinline__ double y_d" << di << "(const Data& D, const double* th, int L0, int L1, int L2) {\n    return " << gen_dbl(d.lhs, env) << ";\n";
                    out << "  }\n  __device__ static __forceinline__ double eta_d" << di << "(const Data& D, const double* th, int L0, int L1, int L2) {\n    return " << gen_dbl(d.iargs[0]->a[0], env) << ";\n";
                }
            }
            out << "  }\n";
        }
    }

    bool is_fused_bern(const CDecl& d) {
        double sz;
        return d.dist == "dbinom" && opts.fuse_links && (is_call(d.iargs[0], "ilogit") || is_call(d.iargs[0], "expit")) && try_ceval(d.iargs[1], {}, sz) && sz == 1.0;
    }
    void emit_inst_loop(int di, const std::string& acc) {
        const CDecl& d = decls[di];
        if (d.ninst == 1) out << "    if (c.rank == 0) " << acc << " += " << lp_call(di, "0") << ";\n";
        else if (is_fused_bern(d) && d.ninst >= 64 && !getenv("APTB200_NO_BATCH")) {
            // fused Bernoulli-logit terms eight rows at a time (explicit ILP: one term is a chain of ~30 dependent FP64
            // operations); the terms are added in the same order as the one-by-one loop
            std::string ycall = lp_call(di, "q + u_ * C::W"), ecall = ycall;
            ycall.replace(0, 4, "y_d");
            ecall.replace(0, 4, "eta_d");
            out << "    { int q = c.rank;\n";
            out << "      for (; q + 7 * C::W < " << d.ninst << "; q += 8 * C::W) {\n        double y_[8], e_[8][1];\n";
            out << "        _Pragma(\"unroll\") for (int u_ = 0; u_ < 8; ++u_) { y_[u_] = " << ycall << "; e_[u_][0] = " << ecall << "; }\n";
            out << "        double a_[1] = {" << acc << "}; apt::bern_logit_tile<8, 1>(y_, e_, a_); " << acc << " = a_[0];\n      }\n";
            out << "      for (; q < " << d.ninst << "; q += C::W) " << acc << " += " << lp_call(di, "q") << ";\n    }\n";
        } else
            out << "    _Pragma(\"unroll 4\") for (int q = c.rank; q < " << d.ninst << "; q += C::W) " << acc << " += " << lp_call(di, "q") << ";\n";
    }

    void emit_unit(size_t ui) {
        Unit& u = units[ui];
        const apt_sampler_control& ctl = u.conf->ctl;
        const bool has_partial = std::any_of(u.groups.begin(), u.groups.end(), [](const Group& g) { return !g.full; });
        out << "  // sampler unit(s) " << u.unit0 << ".." << (u.unit0 + u.count - 1) << ": " << u.conf->type << " on " << u.conf->target
            << (u.count > 1 ? " ... (family of " + std::to_string(u.count) + ")" : "") << "\n";
        out << "  struct S" << ui << " {\n";
        out << "    static constexpr int KIND = " << u.kind << ", D = " << u.d << ", NG = " << u.groups.size() << ", UNIT0 = " << u.unit0
            << ", COUNT = " << u.count << ", AI = " << ctl.adaptInterval << ", BLKOFF = " << u.blkoff << ", EMPOFF = " << u.empoff << ";\n";
        out << "    static constexpr bool LOG = " << (ctl.log ? "true" : "false") << ", REFL = " << (ctl.reflective ? "true" : "false")
            << ", ADAPT = " << (ctl.adaptive ? "true" : "false") << ", SCALE_ONLY = " << (ctl.adaptScaleOnly ? "true" : "false")
            << ", TEMPER_PRIORS = " << ((u.kind == 3 ? ctl.useTempering : ctl.temperPriors) ? "true" : "false") << ", HAS_PARTIAL = " << (has_partial ? "true" : "false")
            << ", DISCRETE = " << (u.discrete ? "true" : "false") << ";\n";
        out << "    static constexpr int MAXSTEPS = " << ctl.sliceMaxSteps << ";\n";
        out << "    static constexpr double NTOTAL = " << fmt_d(u.ntotal) << ";\n";
        out << "    static constexpr double AEXP = " << fmt_d(ctl.adaptFactorExponent) << ";\n";
        auto sw = [&](const char* name, const char* type, std::function<std::string(const Group&)> f) {
            out << "    __device__ static constexpr " << type << " " << name << "(int g) { return ";
            for (size_t g = 0; g + 1 < u.groups.size(); g++) out << "g == " << g << " ? " << f(u.groups[g]) << " : ";
            out << f(u.groups.back()) << "; }\n";
        };
        sw("gdecl", "int", [&](const Group& g) { return std::to_string(decls[g.decl].lpidx); });
        sw("gfull", "bool", [&](const Group& g) { return std::string(g.full ? "true" : "false"); });
        sw("gprior", "bool", [&](const Group& g) { return std::string(g.prior ? "true" : "false"); });
        // targets
        out << "    __device__ static __forceinline__ int tgt(int m, int j) { ";
        bool contiguous = true;
        for (size_t j = 1; j < u.tgt.size(); j++) contiguous &= (u.tgt[j] == u.tgt[0] + (int)j);
        if (contiguous) out << "return " << u.tgt[0] << " + m + j; }\n";
        else {
            out << "switch (j) {";
            for (size_t j = 0; j < u.tgt.size(); j++) out << " case " << j << ": return " << u.tgt[j] << ";";
            out << " } return 0; }\n";
        }
        // bounds
        {
            Env env;
            EP lo = u.lower, hi = u.upper;
            if (lo) {
                std::map<std::string, EP> b;
                for (auto& kv : u.bound_env) b[kv.first] = mk_num(kv.second);
                lo = substitute(lo, b); hi = substitute(hi, b);
            }
            out << "    __device__ static __forceinline__ double lower(const Data& D, const double* th, int m) { return "
                << (lo ? gen_dbl(lo, env) : std::string("(-CUDART_INF)")) << "; }\n";
            out << "    __device__ static __forceinline__ double upper(const Data& D, const double* th, int m) { return "
                << (hi ? gen_dbl(hi, env) : std::string("CUDART_INF")) << "; }\n";
        }
        // eval
        out << "    template <class C> __device__ static __forceinline__ void eval(const Data& D, const double* th, int m, const C& c, double* acc, bool newpass, bool skipbig) {\n";
        for (size_t g = 0; g < u.groups.size(); g++) {
            const Group& gr = u.groups[g];
            const CDecl& d = decls[gr.decl];
            std::string acc = "acc[" + std::to_string(g) + "]";
            out << "      // group " << g << ": declaration " << gr.decl << " (" << to_string(d.lhs) << " ~ " << d.dist << "), "
                << (gr.full ? "FULL" : "PARTIAL") << (gr.prior ? ", target's own node(s)" : "") << "\n";
            if (u.count > 1) {
                out << "      { const int* ptr_ = D.ia[" << u.ia_ptr[g] << "]; const int* ix_ = D.ia[" << u.ia_idx[g] << "];\n";
                out << "        for (int q = ptr_[m] + c.rank; q < ptr_[m + 1]; q += C::W) " << acc << " += " << lp_call(gr.decl, "ix_[q]") << "; }\n";
            } else if (gr.full) {
                out << "      if (newpass" << (gr.big ? " && !skipbig" : "") << ") {\n  ";
                emit_inst_loop(gr.decl, acc);
                out << "      }\n";
            } else if (gr.inst.size() == 1) {
                out << "      if (c.rank == 0) " << acc << " += " << lp_call(gr.decl, std::to_string(gr.inst[0])) << ";\n";
            } else if (u.ia_list[g] >= 0) {
                out << "      { const int* ix_ = D.ia[" << u.ia_list[g] << "];\n";
                out << "        for (int q = c.rank; q < " << gr.inst.size() << "; q += C::W) " << acc << " += " << lp_call(gr.decl, "ix_[q]") << "; }\n";
            } else {
                int step = gr.inst[1] - gr.inst[0];
                out << "      for (int q = c.rank; q < " << gr.inst.size() << "; q += C::W) " << acc << " += "
                    << lp_call(gr.decl, std::to_string(gr.inst[0]) + " + q * " + std::to_string(step)) << ";\n";
            }
        }
        out << "    }\n";
        if (multi > 1 && unit_is_multi_capable(u)) {
            // the same sums for CH chains at once: the data a lane loads for a row is used by CH dot products; per chain the
            // terms are added in the order of eval()
            out << "    template <int CH, class C> __device__ static __forceinline__ void eval_multi(const Data& D, const double* const (&thv)[CH], const C& c, double (&accv)[CH][NG]) {\n";
            for (size_t g = 0; g < u.groups.size(); g++) {
                const Group& gr = u.groups[g];
                const CDecl& d = decls[gr.decl];
                const std::string gs = std::to_string(g);
                if (is_fused_bern(d) && d.ninst >= 512) {
                    auto call = [&](const char* fn, const std::string& thname, const std::string& idx) {
                        std::string cs = lp_call(gr.decl, idx);  // "lp_d<di>(D, th, ...)"
                        cs.replace(0, 4, fn);
                        const size_t pos = cs.find("(D, th,");
                        if (pos == std::string::npos) throw Error("internal: unexpected lp_call form");
                        cs.replace(pos, 7, "(D, " + thname + ",");
                        return cs;
                    };
                    out << "      { int q = c.rank;\n";
                    out << "        for (; q + 3 * C::W < " << d.ninst << "; q += 4 * C::W) {\n          double y_[4], e_[4][CH], a_[CH];\n";
                    out << "          _Pragma(\"unroll\") for (int u_ = 0; u_ < 4; ++u_) {\n            y_[u_] = " << call("y_d", "thv[0]", "q + u_ * C::W") << ";\n";
                    out << "            _Pragma(\"unroll\") for (int ch_ = 0; ch_ < CH; ++ch_) e_[u_][ch_] = " << call("eta_d", "thv[ch_]", "q + u_ * C::W") << ";\n          }\n";
                    out << "          _Pragma(\"unroll\") for (int ch_ = 0; ch_ < CH; ++ch_) a_[ch_] = accv[ch_][" << gs << "];\n";
                    out << "          apt::bern_logit_tile<4, CH>(y_, e_, a_);\n";
                    out << "          _Pragma(\"unroll\") for (int ch_ = 0; ch_ < CH; ++ch_) accv[ch_][" << gs << "] = a_[ch_];\n        }\n";
                    out << "        for (; q < " << d.ninst << "; q += C::W)\n          _Pragma(\"unroll\") for (int ch_ = 0; ch_ < CH; ++ch_) accv[ch_][" << gs << "] += "
                        << call("lp_d", "thv[ch_]", "q") << ";\n      }\n";
                } else {
                    out << "      _Pragma(\"unroll\") for (int ch_ = 0; ch_ < CH; ++ch_) {\n        const double* th = thv[ch_];\n    ";
                    emit_inst_loop(gr.decl, "accv[ch_][" + gs + "]");
                    out << "      }\n";
                }
            }
            out << "    }\n";
        }
        emit_big_eval(u);
        out << "  };\n";
    }

    // wide update (apt_engine.cuh::rw_update_wide): only in cluster mode, for single samplers whose dependants include
    // a FULL declaration with many more instances than one tile could stride through quickly
    bool unit_is_wide(const Unit& u, int cluster, int tiles, int W) const {
        if (cluster <= 1 || u.count != 1 || u.kind >= 2 || getenv("APTB200_NO_WIDE")) return false;
        for (auto& g : u.groups) {
            if (!g.full) return false;  // old values of PARTIAL groups would need a second wide pass
        }
        long mx = 0;
        for (auto& g : u.groups) mx = std::max(mx, decls[g.decl].ninst);
        return mx >= 32L * W;
    }
    // Chains per tile (G::MULTI).  A ladder of many rungs whose only sampler has a large fused Bernoulli-logit likelihood
    // over data shared by all rungs (config 5: 64 rungs x 1024 rows) lets each tile advance MULTI chains together: the
    // rows of X a lane loads feed MULTI dot products (apt_engine.cuh::rw_update_multi), as in the grid engine's tiles.
    int multi = 1;
    bool unit_is_multi_capable(const Unit& u) {
        if (u.count != 1 || u.kind > 1) return false;
        int nbern = 0;
        for (auto& g : u.groups) {
            if (!g.full) return false;
            if (is_fused_bern(decls[g.decl]) && decls[g.decl].ninst >= 512) nbern++;
        }
        return nbern == 1;
    }
    void choose_shape(int& W, int& tiles, int& teams, bool& smem_state) {
        long ndep = 1, items = K;
        for (auto& u : units) {
            items = std::max<long>(items, (long)K * u.count);
            for (auto& g : u.groups) {
                long n = g.full ? decls[g.decl].ninst : (long)g.inst.size();
                ndep = std::max(ndep, n);
            }
        }
        if (opts.lanes_per_chain > 0) W = opts.lanes_per_chain;
        else {
            W = 1;
            while (W < 32 && (long)W * 32 < ndep) W *= 2;
            int kp = 1;
            while (kp < K) kp *= 2;
            while (W < 32 && (long)kp * W < 32) W *= 2;
        }
        if (W < 1 || W > 32 || (W & (W - 1))) throw Error("lanes_per_chain must be a power of two <= 32");
        // several CTAs per SM overlap one ladder's swap phase with another's sweep; a ladder with thousands of
        // concurrent sampler-family members (config 4: 24 rungs x 1000 random effects) gets the widest CTA instead
        int maxthreads = opts.cta_threads > 0 ? opts.cta_threads : (items >= 4096 ? 1024 : 256);
        multi = 1;
        {
            const int want = getenv("APTB200_MULTI") ? atoi(getenv("APTB200_MULTI")) : -1;  // -1: automatic, 0 / 1: off
            const bool capable = units.size() == 1 && unit_is_multi_capable(units[0]) && opts.cluster_ctas <= 1 && opts.engine != 2;
            if (capable && want != 0 && want != 1) {
                const int ch = want > 1 ? want : 2;
                // automatic: only when the rungs do not all fit into one CTA at the chosen width anyway
                if ((ch == 2 || ch == 4) && K % ch == 0 && (want > 1 || (long)K * W > maxthreads)) multi = ch;
            }
            if (multi > 1) {
                // measured on config 5 (64 rungs x 1024 rows, 2048 ladders; tools/tune_c5_multi.sh): 2 chains per tile,
                // 4 lanes, 128-thread CTAs (all 64 rungs in flight), 4 CTAs per SM (128 registers): 1.72e8 updates/s
                // against 8.8e7 for one chain per 32-lane tile in 256-thread CTAs and 1.0e8 for the best single-chain
                // shape; 4 chains per tile lose more in occupancy than they save in loads (1.4e8 at best)
                items = K / multi;
                if (opts.cta_threads <= 0) maxthreads = 128;
                if (opts.lanes_per_chain <= 0)
                    while (W > 4 && items * W > maxthreads) W /= 2;
            }
        }
        tiles = (int)std::min<long>(items, std::max(1, maxthreads / W));
        int team_threads = ((tiles * W + 31) / 32) * 32;
        teams = team_threads == 32 ? 4 : 1;
        size_t bytes = (size_t)K * (P + ND) * 8 * teams;
        // resident state: up to ~96 KB per CTA when several CTAs share an SM, up to 190 KB when a ladder needs the whole SM anyway
        smem_state = bytes <= 96 * 1024 || (teams == 1 && bytes <= 200 * 1024 && tiles * W >= 256);
    }

    // buildAPT with enableWAIC (APT:322-333): there must be data nodes, and every parameter of a data node must be monitored
    // or be downstream of monitored nodes (mcmc_checkWAICmonitors, APT:1313-1341; variable-level like the reference).
    // Deterministic nodes are inlined here, so "stochastic dependents of a variable" are the declarations that read it.
    void check_waic_monitors(const std::vector<int>& mon) {
        bool any_data = false;
        for (int di : stoch_decls) any_data |= vars.at(decls[di].lhs->name).role == Var::DATA;
        if (!any_data) throw Error("WAIC cannot be calculated, as no data nodes were detected in the model.");  // APT:331
        std::set<std::string> monv, top, visited;
        for (int c : mon)
            if (c >= 0)
                for (auto& n : param_order) {
                    const Var& v = vars.at(n);
                    if (c >= v.theta_off && c < v.theta_off + v.size) monv.insert(n);
                }
        // parents of each stochastic declaration: PARAM variables it reads (patterns that are not its own left-hand side)
        auto parents = [&](int di) {
            std::set<std::string> ps;
            for (auto& p : patterns[di])
                if (!p.own) ps.insert(p.v->name);
            return ps;
        };
        for (auto& n : param_order) top.insert(n);
        for (int di : stoch_decls)
            if (decls[di].lhs_param && !parents(di).empty()) top.erase(decls[di].lhs->name);
        std::set<std::string> cur;
        for (auto& n : top)
            if (!monv.count(n)) cur.insert(n);
        while (!cur.empty()) {
            visited.insert(cur.begin(), cur.end());
            std::set<std::string> next, bad;
            for (int di : stoch_decls) {
                bool dep = false;
                for (auto& pn : parents(di)) dep |= cur.count(pn) != 0;
                if (!dep) continue;
                const Var& lv = vars.at(decls[di].lhs->name);
                if (lv.role == Var::DATA) bad.insert(lv.name);
                else if (!monv.count(lv.name) && !visited.count(lv.name)) next.insert(lv.name);
            }
            if (!bad.empty()) {
                std::string l;
                for (auto& b : bad) l += (l.empty() ? "" : ", ") + b;
                throw Error("In order for a valid WAIC calculation, all parameters of data nodes in the model must be monitored, or be"
                            " downstream from monitored nodes. See help(buildMCMC) for more information on valid sets of monitored nodes"
                            " for WAIC calculations.\n Currently, the following data nodes have un-monitored upstream parameters:\n " + l);
            }
            cur = next;
        }
        // model$simulate(paramDeps) (APT:327,618) would draw every unmonitored stochastic node downstream of a monitored one
        // from its conditional prior; the device engine does not simulate, so such a configuration is refused
        std::set<std::string> down;
        for (bool grew = true; grew;) {
            grew = false;
            for (int di : stoch_decls) {
                const std::string& ln = decls[di].lhs->name;
                if (!decls[di].lhs_param || monv.count(ln) || down.count(ln)) continue;
                for (auto& pn : parents(di))
                    if (monv.count(pn) || down.count(pn)) { down.insert(ln); grew = true; break; }
            }
        }
        if (!down.empty())
            throw Error("calculateWAIC: the unmonitored node '" + *down.begin() + "' is downstream of monitored nodes; the reference would"
                        " simulate it from its prior for every sample (APT:618), which the device engine does not do. Monitor it.");
    }

    Lowered run() {
        canonicalise();
        build_vars();
        inline_all();
        build_patterns();
        build_units();
        // monitors
        std::vector<int> mon, mon2;
        auto names = theta_names();
        std::vector<std::string> mon_names, mon2_names;
        // node ids of the logProb monitors: stochastic declarations in model order, instances in loop order
        std::vector<long> lp_node_off;
        long n_lp_nodes = 0;
        for (int di : stoch_decls) { lp_node_off.push_back(n_lp_nodes); n_lp_nodes += decls[di].ninst; }
        auto expand_logprob = [&](const std::string& m, std::vector<int>& dst, std::vector<std::string>& dn) {
            // "logProb_y", "logProb_y[3]", "logProb_y[2:5]" (nimble's logProb_<var> model variables)
            EP e = parse_expr_text(m.substr(8));
            if (e->k != Expr::VAR) throw Error("cannot parse monitor '" + m + "'");
            auto it = vars.find(e->name);
            if (it == vars.end() || (it->second.role != Var::PARAM && it->second.role != Var::DATA))
                throw Error("logProb monitor of '" + e->name + "', which is not a stochastic node");
            const Var& v = it->second;
            std::vector<int64_t> want, els;
            ref_elements(v, e, {}, want);
            std::set<int64_t> ws(want.begin(), want.end());
            std::map<std::string, double> env;
            int found = 0;
            for (size_t q = 0; q < stoch_decls.size(); q++) {
                const CDecl& d = decls[stoch_decls[q]];
                if (d.lhs->name != e->name) continue;
                for (long inst = 0; inst < d.ninst; inst++) {
                    inst_env(d, inst, env);
                    ref_elements(v, d.lhs, env, els);
                    if (els.empty() || !ws.count(els[0])) continue;  // a multivariate node is named by its first element
                    dst.push_back(-1 - (int)(lp_node_off[q] + inst));
                    std::string nm = "logProb_" + v.name;
                    if (!v.dims.empty()) {
                        nm += "[" + std::to_string(els[0] % v.dims[0] + 1);
                        if (v.dims.size() > 1) nm += ", " + std::to_string(els[0] / v.dims[0] + 1);
                        nm += "]";
                    }
                    dn.push_back(nm);
                    found++;
                }
            }
            if (!found) throw Error("monitor '" + m + "' matches no stochastic node");
        };
        auto expand_mon = [&](const std::vector<std::string>& src, std::vector<int>& dst, std::vector<std::string>& dn) {
            for (auto& m : src) {
                if (m.rfind("logProb_", 0) == 0) { expand_logprob(m, dst, dn); continue; }
                for (int t : expand_target(m)) { dst.push_back(t); dn.push_back(names[(size_t)t]); }
            }
        };
        expand_mon(spec.monitors, mon, mon_names);
        expand_mon(spec.monitors2, mon2, mon2_names);
        if (opts.enable_waic) check_waic_monitors(mon);
        int DMAX = 1;
        bool has_family = false;
        for (auto& u : units) {
            DMAX = std::max(DMAX, u.d);
            has_family |= u.count > 1;
            if (u.count > 1 && u.groups.size() > 4) throw Error("sampler families with more than 4 dependent declarations are not supported");
        }
        int W, tiles, teams;
        bool smem_state;
        choose_shape(W, tiles, teams, smem_state);
        // Cluster mode: with few ladders and a CTA-wide team whose work items (sampler-family members x rungs) exceed
        // one CTA's tiles many times over, a ladder spans a thread-block cluster of up to 8 CTAs (one per SM); chain
        // state then lives in global memory and the stages are separated by cluster barriers.
        int cluster = 1;
        {
            long items = K;
            for (auto& u : units) items = std::max<long>(items, (long)K * u.count);
            const int nl = std::max(1, opts.n_ladders);
            if (teams == 1 && items >= 8L * tiles && !getenv("APTB200_NO_CLUSTER")) {
                while (cluster < 8 && (long)nl * cluster * 2 <= 148 && items >= 4L * tiles * cluster * 2) cluster *= 2;
            }
            if (opts.cluster_ctas > 0) {
                if (opts.cluster_ctas > 8 || (opts.cluster_ctas & (opts.cluster_ctas - 1))) throw Error("cluster_ctas must be 1, 2, 4 or 8");
                if (teams != 1 && opts.cluster_ctas > 1) throw Error("cluster_ctas > 1 needs a CTA-wide team (lanes_per_chain * chains in flight > 32)");
                cluster = opts.cluster_ctas;
            }
            if (multi > 1) cluster = 1;
            if (cluster > 1) smem_state = false;
        }
        // ---- engine selection.  Grid engine: few chains, at least one declaration too large for one CTA per
        // chain (config 3: N = 1e6 rows): every tempered update of all chains is served by ONE pass over the data.
        int engine = 1, TB = 1, R = 1, csplit = 1;
        {
            long maxfull = 0;
            for (auto& u : units)
                for (auto& g : u.groups)
                    if (g.full) maxfull = std::max(maxfull, decls[g.decl].ninst);
            const long nchains = (long)std::max(1, opts.n_ladders) * K;
            if (opts.engine == 2 || (opts.engine == 0 && !has_family && nchains <= 256 && maxfull >= 65536)) engine = 2;
            for (auto& u : units) {
                if (u.kind == 3) {
                    if (opts.engine == 2) throw Error("grid engine: sampler_RW_multinomial_tempered is not supported (use engine = 1)");
                    engine = 1;
                }
                if (u.kind == 2) {
                    if (opts.engine == 2) throw Error("grid engine: sampler_slice_tempered is not supported (use engine = 1)");
                    engine = 1;
                }
            }
            if (engine == 2) {
                if (has_family) throw Error("grid engine: sampler families are not supported yet (use engine = 1)");
                const long big_min = opts.engine == 2 ? 256 : 8192;
                bool any = false, all = true;
                for (auto& u : units) {
                    int nb = 0;
                    for (auto& g : u.groups)
                        if (g.full && decls[g.decl].ninst >= big_min && decls[g.decl].loops.size() == 1 && decls[g.decl].dist != "dmnorm" && decls[g.decl].dist != "dcat") { g.big = true; nb++; any = true; }
                    if (nb > 2) throw Error("grid engine: at most 2 large declarations per sampler are supported");
                    if (nb == 0) all = false;
                }
                if (!all && opts.engine != 2) {  // auto: a sampler without a large declaration belongs to the ladder-resident engine
                    engine = 1;
                    for (auto& u : units)
                        for (auto& g : u.groups) g.big = false;
                } else {
                if (!any) throw Error("grid engine requested but no declaration is large enough to be evaluated grid-wide");
                if (!all) throw Error("grid engine: every sampler needs a large declaration among its dependents (use engine = 1)");
                TB = opts.rungs_per_pass > 0 ? opts.rungs_per_pass : 16;
                while (TB > 1 && TB / 2 >= nchains) TB /= 2;
                if (TB < 1 || TB > 32 || (TB & (TB - 1))) throw Error("rungs_per_pass must be a power of two <= 32");
                // scalar streaming path (used when tile mode does not apply): one row per thread, all TB chains.  Two rows
                // per thread and chains split over two threads were measured slower and survive only as tuning switches.
                R = getenv("APTB200_GRID_R2") ? 2 : 1;
                csplit = (getenv("APTB200_GRID_CSPLIT2") && TB >= 2) ? 2 : 1;
                W = std::min(32, 256 / TB);  // control lanes per chain inside the finishing CTA
                ring_ok = (R == 1) && !getenv("APTB200_NO_RING");
                // tile mode: TBT chains x RT rows per thread, CS = TB / TBT lanes share a row group
                tile_want = ring_ok && csplit == 1 && TB >= 2 && !getenv("APTB200_NO_TILE");
                {
                    int tbt = TB >= 16 ? 4 : 2;
                    if (const char* e = getenv("APTB200_TILE_TBT")) tbt = std::max(2, std::min(TB, atoi(e)));  // tuning knobs (scratch/)
                    tile_cs = std::max(1, TB / tbt);
                    tile_rt = std::min(4, 2 * tile_cs);  // 4 rows x 4 chains measured best at 16 rungs (8 x 4 spills at 128 registers)
                    if (const char* e = getenv("APTB200_TILE_RT")) tile_rt = std::max(2, atoi(e) & ~1);
                    tile_rows = (32 / tile_cs) * tile_rt;
                    // the X beta product of a tile on the FP64 tensor-core path: 8 or 16 rungs per pass (1 or 2 chain blocks of 8)
                    tile_mma = (TB == 8 || TB == 16) && !getenv("APTB200_NO_MMA") && !getenv("APTB200_TILE_TBT") && !getenv("APTB200_TILE_RT");
                    if (tile_mma) { tile_cs = 4; tile_rt = 4; tile_rows = 32; tile_tb = TB; }
                }
                }
            }
        }

        if (engine == 2) multi = 1;  // chains per tile belong to the ladder-resident engine
        std::ostringstream body;
        std::swap(out, body);  // emit model body first to learn which arrays are used
        emit_lp_functions();
        // log-probability of one stochastic node by id (logProb monitors; evaluated only when a row is written)
        out << "  __device__ static double lp_node(const Data& D, const double* th, int id) {\n";
        {
            long off = 0;
            for (int di : stoch_decls) {
                out << "    if (id < " << off + decls[di].ninst << ") return " << lp_call(di, "(id - " + std::to_string(off) + ")") << ";\n";
                off += decls[di].ninst;
            }
        }
        out << "    return CUDART_NAN;\n  }\n";
        // data nodes (model$getNodeNames(dataOnly = TRUE), APT:323) for calculateWAIC: node j of the walk -> its log-density
        {
            long ndata = 0;
            std::ostringstream fn;
            for (int di : stoch_decls) {
                if (vars.at(decls[di].lhs->name).role != Var::DATA) continue;
                fn << "    if (j < " << ndata + decls[di].ninst << "L) return " << lp_call(di, "(int)(j - " + std::to_string(ndata) + "L)") << ";\n";
                ndata += decls[di].ninst;
            }
            out << "  static constexpr long NDATA = " << ndata << ";\n";
            out << "  __device__ static __forceinline__ double lp_data(const Data& D, const double* th, long j) {\n" << fn.str() << "    return CUDART_NAN;\n  }\n";
        }
        out << "  template <class C> __device__ static __forceinline__ void logprob_all(const Data& D, const double* th, const C& c, double* acc) {\n";
        for (int di : stoch_decls) {
            out << "  ";
            emit_inst_loop(di, "acc[" + std::to_string(decls[di].lpidx) + "]");
        }
        out << "  }\n";
        for (size_t ui = 0; ui < units.size(); ui++) emit_unit(ui);
        if (engine == 2) {
            // dispatchers used by the grid kernels: propose-ahead for an arbitrary stage, stage launch on the host
            out << "  template <class C> __device__ static __forceinline__ void propose_stage(int s, C& c, const Data& D, int k, double* pend, const double* delta) {\n    switch (s) {\n";
            for (size_t i = 0; i < units.size(); i++)
                out << "      case " << i << ": { apt::Pend<S" << i << "::NG> p_; apt::rw_propose<Model, S" << i << ">(c, D, k, 0, p_, " << (units[i].kind == 1 ? "delta" : "nullptr")
                    << "); apt::store_pend<S" << i << "::NG>(pend, p_, c.rank); break; }\n";
            out << "    }\n  }\n";
            out << "  __host__ __device__ static constexpr bool stage_is_block(int s) { return ";
            for (size_t i = 0; i < units.size(); i++) out << "s == " << i << " ? " << (units[i].kind == 1 ? "true" : "false") << " : ";
            out << "false; }\n";
            out << "  __host__ __device__ static constexpr int stage_unit0(int s) { return ";
            for (size_t i = 0; i < units.size(); i++) out << "s == " << i << " ? " << units[i].unit0 << " : ";
            out << "0; }\n";
            out << "  template <class C> __device__ static __forceinline__ void delta_stage(int s, C& c, int k, double* z, double* outp) {\n    switch (s) {\n";
            for (size_t i = 0; i < units.size(); i++)
                if (units[i].kind == 1) out << "      case " << i << ": apt::block_increment<Model, S" << i << ">(c, k, 0, z, outp); break;\n";
            out << "      default: break;\n    }\n  }\n";
            out << "  template <class F> __device__ static __forceinline__ void with_stage(int s, F&& f) {\n    switch (s) {\n";
            for (size_t i = 0; i < units.size(); i++) out << "      case " << i << ": f(apt::StageTag<S" << i << ">{}); break;\n";
            out << "    }\n  }\n";
            // one cooperative launch per chunk of iterations: every CTA must be resident (they wait for each other)
            out << "  static cudaError_t launch_run(const apt::GArgs& a, int grid, size_t smem, cudaStream_t st) {\n";
            out << "    void* args_[] = {(void*)&a};\n";
            out << "    return cudaLaunchCooperativeKernel((void*)apt::grid_run_kernel<Model>, dim3((unsigned)grid), dim3(GRID_THREADS), args_, smem, st);\n  }\n";
            out << "  static void set_stage_smem(size_t smem) { cudaFuncSetAttribute(apt::grid_run_kernel<Model>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }\n";
            out << "  static int stage_occupancy(int, size_t smem) { int n = 1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, apt::grid_run_kernel<Model>, GRID_THREADS, smem); return n; }\n";
        }
        // sweep: stages
        if (engine != 2) {
            out << "  template <class C> __device__ static __forceinline__ void propose_stage(int, C&, const Data&, int, double*, const double*) {}\n";
            out << "  __host__ __device__ static constexpr bool stage_is_block(int) { return false; }\n  __host__ __device__ static constexpr int stage_unit0(int) { return 0; }\n";
            out << "  template <class C> __device__ static __forceinline__ void delta_stage(int, C&, int, double*, double*) {}\n";
            out << "  static cudaError_t launch_run(const apt::GArgs&, int, size_t, cudaStream_t) { return cudaSuccess; }\n";
            out << "  static int stage_occupancy(int, size_t) { return 1; }\n  static void set_stage_smem(size_t) {}\n";
        }
        out << "  template <class C> __device__ static __forceinline__ void sweep(C& c, const Data& D, bool active) {\n";
        size_t ui = 0;
        bool first = true;
        while (ui < units.size()) {
            if (!first) out << "    c.team_sync();\n";
            first = false;
            if (units[ui].count > 1) {
                out << "    if (active) for (int it_ = c.tile_id; it_ < K * " << units[ui].count << "; it_ += c.ntiles) apt::rw_update<Model, S" << ui
                    << ">(c, D, it_ / " << units[ui].count << ", it_ % " << units[ui].count << ");\n";
                out << "    c.team_sync();\n    if (active) apt::family_commit<Model, S" << ui << ">(c, D);\n";
                ui++;
            } else if (unit_is_wide(units[ui], cluster, tiles, W)) {
                // cluster mode, a sampler whose dependants are a large declaration: the whole cluster evaluates ONE
                // chain at a time (K x W lanes could not cover the latency of thousands of terms per lane)
                out << "    for (int k = 0; k < K; k++) apt::rw_update_wide<Model, S" << ui << ">(c, D, k, active);\n";
                ui++;
            } else if (multi > 1 && unit_is_multi_capable(units[ui])) {
                out << "    if (active) for (int k = c.tile_id * MULTI; k < K; k += c.ntiles * MULTI) apt::rw_update_multi<Model, S" << ui << ", MULTI>(c, D, k);\n";
                ui++;
            } else {
                out << "    if (active) for (int k = c.tile_id; k < K; k += c.ntiles) {\n";
                while (ui < units.size() && units[ui].count == 1 && !unit_is_wide(units[ui], cluster, tiles, W)) {
                    out << "      apt::" << (units[ui].kind == 2 ? "slice_update" : units[ui].kind == 3 ? "multinomial_update" : "rw_update") << "<Model, S" << ui << ">(c, D, k, 0);\n";
                    ui++;
                }
                out << "    }\n";
            }
        }
        out << "  }\n";
        std::swap(out, body);  // `body` now holds the model body; `out` is the file

        out << "// Generated by the aptb200 lowering step (nimbleapt_b200/csrc/lower.cpp). Do not edit.\n";
        if (engine == 2) out << "#define APT_SP_EXP_POLY 1  // see apt_softplus.cuh::apt_softplus_neg_batch\n";
        out << "#include \"apt_grid.cuh\"\n#include \"apt_host.cuh\"\n";
        out << "__device__ __forceinline__ double apt_sq(double x) { return x * x; }\n";
        out << "// unnamed namespace: every lowered model defines a type called aptgen::Model; internal linkage keeps the\n// template instantiations of different modules loaded into one process apart.\nnamespace {\nnamespace aptgen {\nusing apt::Data;\n";
        auto emit_tab = [&](const char* name, const std::vector<int>& v) {
            out << "__device__ const int " << name << "[" << std::max<size_t>(v.size(), 1) << "] = {";
            for (size_t i = 0; i < v.size(); i++) out << (i ? ", " : "") << v[i];
            if (v.empty()) out << "0";
            out << "};\n";
        };
        emit_tab("MON_TAB", mon);
        emit_tab("MON2_TAB", mon2);
        for (size_t i = 0; i < units.size(); i++)
            if (units[i].kind == 1 && !units[i].conf->propCov.empty()) {
                out << "static const double PROPCOV" << i << "[" << units[i].conf->propCov.size() << "] = {";
                for (size_t q = 0; q < units[i].conf->propCov.size(); q++) out << (q ? ", " : "") << fmt_d(units[i].conf->propCov[q]);
                out << "};\n";
            }
        out << "struct Model {\n";
        out << "  static constexpr int P = " << P << ", ND = " << ND << ", NU = " << NU << ", NSAMP = " << units.size() << ", K = " << K
            << ", NMON = " << mon.size() << ", NMON2 = " << mon2.size() << ", DMAX = " << DMAX << ", W = " << W << ";\n";
        {
            bool lpmon = false;
            for (int c : mon) lpmon |= c < 0;
            for (int c : mon2) lpmon |= c < 0;
            out << "  static constexpr bool HAS_LP_MON = " << (lpmon ? "true" : "false") << ";\n";
        }
        out << "  static constexpr int NARR = " << used_arrays.size() << ", NIARR = " << iarrays.size() << ", BLKSZ = " << BLKSZ << ", EMPSZ = " << EMPSZ << ";\n";
        out << "  static constexpr int TILES_PER_LADDER = " << tiles << ", TEAMS_PER_CTA = " << teams << ", CLUSTER = " << (engine == 1 ? cluster : 1) << ", MULTI = " << multi << ";\n";
        {
            // resident CTAs per SM the ladder kernel is compiled for (register cap = 64K / (LADDER_MINB * CTA)): as many
            // as shared memory admits, but never fewer than 64 registers per thread
            const int team_threads = ((tiles * W + 31) / 32) * 32, cta = team_threads * teams;
            const bool logtab = team_threads != 32 && K <= 32;
            const long psize = logtab ? 2L * K * K : ((long)K * (K - 1) / 2 + 1) / 2 * 2;
            const long team_bytes = ((7L * K + psize + (long)tiles * multi * 2 * DMAX + (smem_state && engine == 1 ? (long)K * (P + ND) : 0)) * 8 + 2L * K * 4 + 15) / 16 * 16;
            long minb = 232448 / (team_bytes * teams + 1024);
            minb = std::max(1L, std::min(minb, (long)(1024 / cta)));
            if (multi > 1) minb = std::min(minb, (long)std::max(1, 512 / cta));  // two chains per tile: 128 registers
            if (const char* e = getenv("APTB200_LADDER_MINB")) minb = std::max(0, atoi(e));  // 0 = no register cap
            out << "  static constexpr int LADDER_MINB = " << minb << ";\n";
        }
        out << "  static constexpr bool STATE_IN_SMEM = " << (smem_state && engine == 1 ? "true" : "false") << ", HAS_FAMILY = " << (has_family ? "true" : "false") << ";\n";
        out << "  static constexpr int ENGINE = " << engine << ", TB = " << TB << ", R = " << R << ", GRID_THREADS = " << (getenv("APTB200_GRID_THREADS") ? atoi(getenv("APTB200_GRID_THREADS")) : 256) << ", RING_SLOTS = " << ring_slots << ", FRAG_DOUBLES = " << frag_doubles << ", CSPLIT = " << csplit << ", GRID_MINB = " << (getenv("APTB200_GRID_MINB") ? std::max(1, atoi(getenv("APTB200_GRID_MINB"))) : (csplit == 2 ? 3 : 2)) << ", NSTAGE = " << units.size() << ";\n";
        // derived arrays (data-only terms, filled by derived_kernel when data is uploaded)
        out << "  static constexpr int NDER = " << derived.size() << ";\n";
        out << "  static long der_inst(int k) { return ";
        for (size_t k = 0; k < derived.size(); k++) out << "k == " << k << " ? " << decls[derived[k].decl].ninst << "L : ";
        out << "0L; }\n";
        out << "  static int der_width(int k) { return 2; }\n";
        out << "  __device__ static __forceinline__ void der_fill(const Data& D, int k, long i, double* out) {\n    switch (k) {\n";
        for (size_t k = 0; k < derived.size(); k++) {
            const CDecl& d = decls[derived[k].decl];
            out << "      case " << k << ": {";
            long mul = 1;
            for (size_t l = 0; l < 3; l++) {
                if (l < d.loops.size()) {
                    long n = d.loops[l].hi - d.loops[l].lo + 1;
                    out << " const int L" << l << " = " << d.loops[l].lo << " + (int)((i / " << mul << ") % " << n << ");";
                    mul *= n;
                } else out << " const int L" << l << " = 0;";
            }
            out << " (void)L0; (void)L1; (void)L2; apt::dpois_pre(" << gen_dbl(d.lhs, decl_env(d)) << ", out + 2 * i); break; }\n";
        }
        out << "      default: break;\n    }\n  }\n";
        // entries of the per-rung sampler block state that sampler$reset() leaves alone (the multinomial sampler's u, APT:1133-1142)
        out << "  __host__ __device__ static constexpr bool blk_keep(int i) { return ";
        for (auto& u : units)
            if (u.kind == 3) out << "i == " << u.blkoff << " || ";
        out << "false; }\n";
        out << "  __device__ static __forceinline__ int mon(int c) { return MON_TAB[c]; }\n";
        out << "  __device__ static __forceinline__ int mon2(int c) { return MON2_TAB[c]; }\n";
        out << body.str();
        out << "  static const apt::HostSamplerInfo& sampler_info(int s) {\n    static const apt::HostSamplerInfo info[" << units.size() << "] = {\n";
        for (size_t i = 0; i < units.size(); i++) {
            const Unit& u = units[i];
            out << "      {" << u.kind << ", " << u.d << ", " << u.unit0 << ", " << u.count << ", " << u.blkoff << ", " << u.empoff << ", "
                << u.conf->ctl.adaptInterval << ", " << fmt_d(u.kind == 2 ? u.conf->ctl.width : u.conf->ctl.scale) << ", "
                << ((u.kind == 1 && !u.conf->propCov.empty()) ? "PROPCOV" + std::to_string(i) : std::string("nullptr")) << "},\n";
        }
        out << "    };\n    return info[s];\n  }\n};\n}  // namespace aptgen\n}  // namespace\nAPT_DEFINE_MODULE(aptgen::Model)\n";

        Lowered L;
        L.source = out.str();
        L.P = P; L.K = K; L.NU = NU; L.ND = ND; L.NMON = (int)mon.size(); L.NMON2 = (int)mon2.size(); L.DMAX = DMAX; L.W = W; L.engine = engine;
        for (auto* a : used_arrays) { L.arrays.push_back(a); L.array_names.push_back(a->name); }
        L.iarrays = iarrays;
        L.theta0.assign((size_t)P, 0.0);
        for (auto& n : param_order) {
            const Var& v = vars.at(n);
            for (int64_t e = 0; e < v.size; e++) L.theta0[(size_t)(v.theta_off + e)] = v.arr->v[(size_t)e];
        }
        L.param_names = names; L.mon_names = mon_names; L.mon2_names = mon2_names;
        return L;
    }
};

}  // namespace

Lowered lower_model(const ModelSpec& spec, int n_temps, const apt_build_opts& opts) {
    if (n_temps < 2) throw Error("need at least 2 temperatures");
    if (n_temps > 1024) throw Error("at most 1024 temperatures are supported");
    Lowerer lw(spec, n_temps, opts);
    return lw.run();
}

}  // namespace aptf
