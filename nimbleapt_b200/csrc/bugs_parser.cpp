// bugs_parser.cpp — tokenizer + recursive-descent parser for the BUGS subset accepted by the lowering step.
// Grammar (R operator precedence for the operators supported):
//   program := '{'? stmt* '}'?
//   stmt    := 'for' '(' IDENT 'in' expr ':' expr ')' '{' stmt* '}'  |  lhs ('~' | '<-') expr
//   expr    := term (('+'|'-') term)* ; term := unary (('*'|'/') unary)* ; unary := ('-'|'+') unary | power
//   power   := primary ('^' unary)? ; primary := NUM | IDENT | IDENT '(' args ')' | IDENT '[' idx,* ']' | '(' expr ')'
//   idx     := expr (':' expr)?         (':' binds loosest inside an index / for header; write 1:(n-1))
// Comparisons and logical operators (R precedence: arithmetic > comparison > '!' > '&' > '|') are accepted wherever an
// expression is, as calls of reserved names (__lt __gt __le __ge __eq __ne __not __and __or); they are what the bodies of
// user-supplied deterministic functions need (parse_user_function below, demo/APT_demo.R:42-71).
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include "front.h"

namespace aptf {

EP mk_num(double v) { auto e = std::make_shared<Expr>(); e->k = Expr::NUM; e->num = v; return e; }
EP mk_var(const std::string& n, std::vector<EP> idx) { auto e = std::make_shared<Expr>(); e->k = Expr::VAR; e->name = n; e->a = std::move(idx); return e; }
EP mk_bin(char op, EP l, EP r) { auto e = std::make_shared<Expr>(); e->k = Expr::BIN; e->op = op; e->a = {std::move(l), std::move(r)}; return e; }
EP mk_call(const std::string& f, std::vector<EP> args) { auto e = std::make_shared<Expr>(); e->k = Expr::CALL; e->name = f; e->a = std::move(args); return e; }

std::string to_string(const EP& e) {
    std::ostringstream o;
    switch (e->k) {
        case Expr::NUM: o << e->num; break;
        case Expr::VAR:
            o << e->name;
            if (!e->a.empty()) {
                o << "[";
                for (size_t i = 0; i < e->a.size(); i++) o << (i ? ", " : "") << to_string(e->a[i]);
                o << "]";
            }
            break;
        case Expr::NEG: o << "-(" << to_string(e->a[0]) << ")"; break;
        case Expr::BIN: o << "(" << to_string(e->a[0]) << " " << e->op << " " << to_string(e->a[1]) << ")"; break;
        case Expr::RANGE: o << to_string(e->a[0]) << ":" << to_string(e->a[1]); break;
        case Expr::CALL: {
            o << e->name << "(";
            bool first = true;
            for (auto& x : e->a) { o << (first ? "" : ", ") << to_string(x); first = false; }
            for (auto& x : e->kw) { o << (first ? "" : ", ") << x.first << "=" << to_string(x.second); first = false; }
            o << ")";
            break;
        }
    }
    return o.str();
}

namespace {
struct Tok {
    enum T { NUM, ID, SYM, NL, END } t;
    std::string s;
    double v = 0;
    int line = 0;
};

std::vector<Tok> tokenize(const std::string& src) {
    std::vector<Tok> out;
    size_t i = 0, n = src.size();
    int line = 1, depth = 0;
    while (i < n) {
        char c = src[i];
        if (c == '#') { while (i < n && src[i] != '\n') i++; continue; }
        if (c == '\n') { if (depth == 0) out.push_back({Tok::NL, "\n", 0, line}); line++; i++; continue; }
        if (c == ';') { out.push_back({Tok::NL, ";", 0, line}); i++; continue; }
        if (isspace((unsigned char)c)) { i++; continue; }
        if (isdigit((unsigned char)c) || (c == '.' && i + 1 < n && isdigit((unsigned char)src[i + 1]))) {
            char* end = nullptr;
            double v = strtod(src.c_str() + i, &end);
            size_t j = (size_t)(end - src.c_str());
            if (j < n && src[j] == 'L') j++;  // R integer literal suffix
            out.push_back({Tok::NUM, src.substr(i, j - i), v, line});
            i = j;
            continue;
        }
        if (isalpha((unsigned char)c) || c == '_' || c == '.') {
            size_t j = i;
            while (j < n && (isalnum((unsigned char)src[j]) || src[j] == '_' || src[j] == '.')) j++;
            out.push_back({Tok::ID, src.substr(i, j - i), 0, line});
            i = j;
            continue;
        }
        if (c == '<' && i + 1 < n && src[i + 1] == '-') { out.push_back({Tok::SYM, "<-", 0, line}); i += 2; continue; }
        if (i + 1 < n) {
            const std::string two = src.substr(i, 2);
            if (two == "<=" || two == ">=" || two == "==" || two == "!=" || two == "||" || two == "&&") { out.push_back({Tok::SYM, two, 0, line}); i += 2; continue; }
        }
        if (c == '(' || c == '[') depth++;
        if (c == ')' || c == ']') depth--;
        out.push_back({Tok::SYM, std::string(1, c), 0, line});
        i++;
    }
    out.push_back({Tok::END, "", 0, line});
    return out;
}

struct Parser {
    std::vector<Tok> t;
    size_t p = 0;
    const Tok& cur() const { return t[p]; }
    bool is(const char* s) const { return (cur().t == Tok::SYM || cur().t == Tok::ID) && cur().s == s; }
    [[noreturn]] void fail(const std::string& m) const {
        throw Error("BUGS parse error (line " + std::to_string(cur().line) + "): " + m + " near '" + cur().s + "'");
    }
    void expect(const char* s) { if (!is(s)) fail(std::string("expected '") + s + "'"); p++; }
    void skip_nl() { while (cur().t == Tok::NL) p++; }

    EP expr() {
        EP l = land();
        while (is("|") || is("||")) { p++; skip_nl(); l = mk_call("__or", {l, land()}); }
        return l;
    }
    EP land() {
        EP l = lnot();
        while (is("&") || is("&&")) { p++; skip_nl(); l = mk_call("__and", {l, lnot()}); }
        return l;
    }
    EP lnot() {
        if (is("!")) { p++; return mk_call("__not", {lnot()}); }
        return cmp();
    }
    EP cmp() {
        EP l = arith();
        static const char* ops[6][2] = {{"<", "__lt"}, {">", "__gt"}, {"<=", "__le"}, {">=", "__ge"}, {"==", "__eq"}, {"!=", "__ne"}};
        for (auto& o : ops)
            if (is(o[0])) { p++; skip_nl(); return mk_call(o[1], {l, arith()}); }
        return l;
    }
    EP arith() {
        EP l = term();
        while (is("+") || is("-")) { char op = cur().s[0]; p++; skip_nl(); l = mk_bin(op, l, term()); }
        return l;
    }
    EP term() {
        EP l = unary();
        while (is("*") || is("/")) { char op = cur().s[0]; p++; skip_nl(); l = mk_bin(op, l, unary()); }
        return l;
    }
    EP unary() {
        if (is("-")) { p++; auto e = std::make_shared<Expr>(); e->k = Expr::NEG; e->a = {unary()}; return e; }
        if (is("+")) { p++; return unary(); }
        return power();
    }
    EP power() {
        EP b = primary();
        if (is("^")) { p++; skip_nl(); return mk_bin('^', b, unary()); }
        return b;
    }
    EP index() {
        EP lo = expr();
        if (is(":")) { p++; auto r = std::make_shared<Expr>(); r->k = Expr::RANGE; r->a = {lo, expr()}; return r; }
        return lo;
    }
    EP primary() {
        if (cur().t == Tok::NUM) { double v = cur().v; p++; return mk_num(v); }
        if (is("(")) { p++; EP e = expr(); expect(")"); return e; }
        if (cur().t == Tok::ID) {
            std::string name = cur().s;
            p++;
            if (name == "TRUE") return mk_num(1);
            if (name == "FALSE") return mk_num(0);
            if (name == "Inf") return mk_num(INFINITY);
            if (is("(")) {
                p++;
                auto e = std::make_shared<Expr>();
                e->k = Expr::CALL; e->name = name;
                while (!is(")")) {
                    if (cur().t == Tok::ID && t[p + 1].t == Tok::SYM && t[p + 1].s == "=") {
                        std::string kwn = cur().s; p += 2;
                        e->kw.push_back({kwn, expr()});
                    } else {
                        e->a.push_back(expr());
                    }
                    if (is(",")) p++; else if (!is(")")) fail("expected ',' or ')'");
                }
                p++;
                return e;
            }
            if (is("[")) {
                p++;
                std::vector<EP> idx;
                while (true) {
                    if (is(",") || is("]")) fail("empty index (write explicit ranges such as 1:n)");
                    idx.push_back(index());
                    if (is(",")) { p++; continue; }
                    break;
                }
                expect("]");
                return mk_var(name, idx);
            }
            return mk_var(name);
        }
        fail("unexpected token");
    }

    void stmts(std::vector<RawDecl>& out, std::vector<Loop>& loops, bool in_block, bool single = false) {
        while (true) {
            skip_nl();
            if (cur().t == Tok::END) { if (in_block) fail("missing '}'"); return; }
            if (is("}")) { if (!in_block) fail("unbalanced '}'"); return; }
            if (is("for")) {
                p++; expect("(");
                if (cur().t != Tok::ID) fail("expected loop variable");
                Loop lp; lp.var = cur().s; p++;
                if (!is("in")) fail("expected 'in'");
                p++;
                lp.lo = arith(); expect(":"); lp.hi = arith(); expect(")");
                skip_nl();
                loops.push_back(lp);
                if (is("{")) { p++; stmts(out, loops, true); expect("}"); }
                else stmts(out, loops, in_block, true);  // a loop body of one statement without braces (demo/APT_demo.R:90-91)
                loops.pop_back();
                if (single) return;
                continue;
            }
            RawDecl d;
            d.line = cur().line;
            d.loops = loops;
            EP lhs = expr();
            if (lhs->k == Expr::CALL) {
                if (lhs->a.size() != 1 || lhs->a[0]->k != Expr::VAR) fail("left-hand side must be a node or link(node)");
                d.link = lhs->name;
                lhs = lhs->a[0];
            }
            if (lhs->k != Expr::VAR) fail("left-hand side must be a node");
            d.lhs = lhs;
            if (is("~")) d.stoch = true;
            else if (is("<-")) d.stoch = false;
            else fail("expected '~' or '<-'");
            p++; skip_nl();
            d.rhs = expr();
            if (d.stoch && d.rhs->k != Expr::CALL) fail("right-hand side of '~' must be a distribution");
            if (cur().t != Tok::NL && cur().t != Tok::END && !is("}")) fail("unexpected trailing token");
            out.push_back(d);
            if (single) return;
        }
    }
};
// ---- user-supplied deterministic functions: nimbleFunction(run = function(k = double(), v = double(1), ...) { ... })
struct FStmt {
    enum Kind { ASSIGN, RETURN, IF } k = ASSIGN;
    std::string name;           // ASSIGN
    EP e;                       // ASSIGN value / RETURN value / IF condition
    std::vector<FStmt> a, b;    // IF: then / else
};
struct FParser : Parser {
    std::vector<FStmt> block() {
        std::vector<FStmt> out;
        skip_nl();
        if (is("{")) {
            p++;
            while (true) {
                skip_nl();
                if (is("}")) { p++; break; }
                if (cur().t == Tok::END) fail("missing '}' in function body");
                stmt(out);
            }
        } else stmt(out);
        return out;
    }
    void skip_call() {  // IDENT '(' balanced ')'
        p++; expect("(");
        int depth = 1;
        while (depth > 0) {
            if (cur().t == Tok::END) fail("unbalanced '('");
            if (is("(")) depth++;
            if (is(")")) depth--;
            p++;
        }
    }
    void stmt(std::vector<FStmt>& out) {
        if (is("if")) {
            p++; expect("(");
            FStmt s; s.k = FStmt::IF; s.e = expr(); expect(")");
            s.a = block();
            const size_t save = p;
            skip_nl();
            if (is("else")) { p++; s.b = block(); } else p = save;
            out.push_back(std::move(s));
            return;
        }
        if (is("return")) {
            p++; expect("(");
            FStmt s; s.k = FStmt::RETURN; s.e = expr(); expect(")");
            out.push_back(std::move(s));
            return;
        }
        if (is("returnType")) { skip_call(); return; }
        if (cur().t != Tok::ID) fail("unsupported statement in a user-supplied function (assignments, if / else, return)");
        FStmt s; s.k = FStmt::ASSIGN; s.name = cur().s; p++;
        if (is("[")) fail("assignment to an element of a vector is not supported in user-supplied functions");
        if (!is("<-") && !is("=")) fail("expected '<-'");
        p++; skip_nl();
        s.e = expr();
        out.push_back(std::move(s));
    }
};

using FEnv = std::map<std::string, EP>;
// slice argument bound to a vector parameter, indexed by `idx` inside the body: rfx[1:K][k] -> rfx[1 + k - 1]
EP index_into(const EP& arg, const EP& idx) {
    if (arg->k != Expr::VAR) throw Error("vector argument of a user-supplied function must be a variable or a slice of one: " + to_string(arg));
    if (arg->a.empty()) return mk_var(arg->name, {idx});
    std::vector<EP> ind = arg->a;
    int nr = 0;
    for (auto& x : ind)
        if (x->k == Expr::RANGE) {
            nr++;
            const EP lo = x->a[0];
            x = (lo->k == Expr::NUM && lo->num == 1.0) ? idx : mk_bin('-', mk_bin('+', lo, idx), mk_num(1));
        }
    if (nr != 1) throw Error("vector argument of a user-supplied function must have exactly one range: " + to_string(arg));
    return mk_var(arg->name, ind);
}
EP fsubst(const EP& e, const FEnv& env) {
    if (e->k == Expr::VAR) {
        auto it = env.find(e->name);
        if (it != env.end()) {
            if (e->a.empty()) return it->second;
            if (e->a.size() != 1 || e->a[0]->k == Expr::RANGE) throw Error("user-supplied functions may index their vector arguments with one scalar index: " + to_string(e));
            return index_into(it->second, fsubst(e->a[0], env));
        }
    }
    auto r = std::make_shared<Expr>(*e);
    for (auto& x : r->a) x = fsubst(x, env);
    for (auto& x : r->kw) x.second = fsubst(x.second, env);
    return r;
}
// symbolic execution of a statement list: the value returned as one expression (nullptr: falls off the end)
EP fexec(const std::vector<FStmt>& st, size_t i, FEnv& env) {
    for (; i < st.size(); i++) {
        const FStmt& s = st[i];
        if (s.k == FStmt::ASSIGN) { env[s.name] = fsubst(s.e, env); continue; }
        if (s.k == FStmt::RETURN) return fsubst(s.e, env);
        const EP c = fsubst(s.e, env);
        FEnv ea = env, eb = env;
        const EP ra = fexec(s.a, 0, ea), rb = fexec(s.b, 0, eb);
        if (ra && rb) return mk_call("__if", {c, ra, rb});
        if (ra) { const EP rest = fexec(st, i + 1, eb); if (!rest) throw Error("a user-supplied function must return a value on every path"); return mk_call("__if", {c, ra, rest}); }
        if (rb) { const EP rest = fexec(st, i + 1, ea); if (!rest) throw Error("a user-supplied function must return a value on every path"); return mk_call("__if", {c, rest, rb}); }
        // neither branch returns: merge what they assigned
        for (auto& kv : ea) {
            auto ib = eb.find(kv.first);
            const EP vb = ib != eb.end() ? ib->second : nullptr;
            if (vb && to_string(vb) == to_string(kv.second)) env[kv.first] = kv.second;
            else if (vb) env[kv.first] = mk_call("__if", {c, kv.second, vb});
            else env[kv.first] = kv.second;  // defined on one path only: R would fail on the other if it were read
        }
        for (auto& kv : eb)
            if (!ea.count(kv.first)) env[kv.first] = kv.second;
    }
    return nullptr;
}
}  // namespace

UserFunction parse_user_function(const std::string& name, const std::string& src) {
    FParser ps;
    ps.t = tokenize(src);
    ps.skip_nl();
    // accept  nimbleFunction(run = function(...) {...})  as well as the bare  function(...) {...}
    bool wrapped = false;
    if (ps.is("nimbleFunction")) {
        ps.p++; ps.expect("(");
        ps.skip_nl();
        if (!ps.is("run")) ps.fail("expected run = function(...)");
        ps.p++; ps.expect("=");
        wrapped = true;
    }
    ps.skip_nl();
    if (!ps.is("function")) ps.fail("expected 'function'");
    ps.p++; ps.expect("(");
    UserFunction f;
    f.name = name;
    while (!ps.is(")")) {
        ps.skip_nl();
        if (ps.cur().t != Tok::ID) ps.fail("expected a parameter name");
        UserFunction::Param pr;
        pr.name = ps.cur().s; ps.p++;
        if (ps.is("=")) {
            ps.p++;
            if (!(ps.is("double") || ps.is("integer") || ps.is("logical"))) ps.fail("parameter types must be double(), double(1), integer() ...");
            ps.p++; ps.expect("(");
            if (ps.cur().t == Tok::NUM) { pr.ndim = (int)ps.cur().v; ps.p++; while (!ps.is(")")) ps.p++; }
            ps.expect(")");
        }
        if (pr.ndim > 1) ps.fail("matrix parameters are not supported in user-supplied functions");
        f.params.push_back(pr);
        ps.skip_nl();
        if (ps.is(",")) ps.p++;
    }
    ps.p++;
    auto body = std::make_shared<std::vector<FStmt>>(ps.block());
    ps.skip_nl();
    if (wrapped) { if (ps.is(")")) ps.p++; ps.skip_nl(); }
    if (ps.cur().t != Tok::END) ps.fail("text after the end of the function");
    f.call = [body, name](const std::vector<EP>& args, const std::vector<std::string>& pnames) {
        FEnv env;
        if (args.size() != pnames.size()) throw Error("user-supplied function " + name + "() takes " + std::to_string(pnames.size()) + " arguments");
        for (size_t i = 0; i < args.size(); i++) env[pnames[i]] = args[i];
        const EP r = fexec(*body, 0, env);
        if (!r) throw Error("user-supplied function " + name + "() does not return a value");
        return r;
    };
    return f;
}

EP expand_user_calls(const EP& e, const std::map<std::string, UserFunction>& fns, int depth) {
    if (depth > 16) throw Error("user-supplied functions nest too deeply (recursion is not supported)");
    auto r = std::make_shared<Expr>(*e);
    for (auto& x : r->a) x = expand_user_calls(x, fns, depth);
    for (auto& x : r->kw) x.second = expand_user_calls(x.second, fns, depth);
    if (r->k == Expr::CALL) {
        auto it = fns.find(r->name);
        if (it != fns.end()) {
            std::vector<EP> args = r->a;
            std::vector<std::string> pn;
            for (auto& p_ : it->second.params) pn.push_back(p_.name);
            // keyword arguments in any order
            if (!r->kw.empty()) {
                args.resize(pn.size());
                for (auto& kv : r->kw) {
                    size_t j = 0;
                    while (j < pn.size() && pn[j] != kv.first) j++;
                    if (j == pn.size()) throw Error("user-supplied function " + r->name + "() has no parameter '" + kv.first + "'");
                    args[j] = kv.second;
                }
                for (auto& a_ : args)
                    if (!a_) throw Error("user-supplied function " + r->name + "(): missing argument");
            }
            return expand_user_calls(it->second.call(args, pn), fns, depth + 1);
        }
    }
    return r;
}

std::vector<RawDecl> parse_bugs(const std::string& code) {
    Parser ps;
    ps.t = tokenize(code);
    std::vector<RawDecl> out;
    std::vector<Loop> loops;
    ps.skip_nl();
    bool braces = ps.is("{");
    if (braces) ps.p++;
    ps.stmts(out, loops, braces);
    if (braces) ps.expect("}");
    ps.skip_nl();
    if (ps.cur().t != Tok::END) ps.fail("text after the end of the model");
    return out;
}

EP parse_expr_text(const std::string& s) {
    Parser ps;
    ps.t = tokenize(s);
    ps.skip_nl();
    EP e = ps.expr();
    ps.skip_nl();
    if (ps.cur().t != Tok::END) ps.fail("unexpected text");
    return e;
}

}  // namespace aptf
