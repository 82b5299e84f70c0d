// bugs_parser.cpp — tokenizer + recursive-descent parser for the BUGS subset accepted by the lowering step.
// Grammar (R operator precedence for the operators supported):
//   program := '{'? stmt* '}'?
//   stmt    := 'for' '(' IDENT 'in' expr ':' expr ')' '{' stmt* '}'  |  lhs ('~' | '<-') expr
//   expr    := term (('+'|'-') term)* ; term := unary (('*'|'/') unary)* ; unary := ('-'|'+') unary | power
//   power   := primary ('^' unary)? ; primary := NUM | IDENT | IDENT '(' args ')' | IDENT '[' idx,* ']' | '(' expr ')'
//   idx     := expr (':' expr)?         (':' binds loosest inside an index / for header; write 1:(n-1))
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

    void stmts(std::vector<RawDecl>& out, std::vector<Loop>& loops, bool in_block) {
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
                lp.lo = expr(); expect(":"); lp.hi = expr(); expect(")");
                skip_nl(); expect("{");
                loops.push_back(lp);
                stmts(out, loops, true);
                expect("}");
                loops.pop_back();
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
        }
    }
};
}  // namespace

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
