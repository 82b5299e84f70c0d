"""oracle/bugs_front.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

An independent (Python `ast`-based) front end for the BUGS subset, used ONLY to feed the CPU oracle
(oracle/apt_oracle.c).  It shares no code with the product's C++ parser / lowering
(nimbleapt_b200/csrc/), so oracle-vs-engine parity also cross-checks the lowering.

It restates, for the supported subset, what nimble does between `nimbleModel()` and the sampler `setup`
functions of the reference (all [NIMBLE-UNVERIFIED]):
  * declarations -> nodes; BUGS parameterisations -> canonical ones
    (dnorm(mean, tau | sd= | var=), dgamma(shape, rate | scale=), dbinom(prob, size), dexp(rate | scale=),
     dmnorm(mean, cholesky=, prec_param=) as used by nimbleAPT/vignettes/nimbleAPT-vignette.Rmd:56-64);
  * link functions on the left-hand side -> inverse link on the right;
  * `model$getDependencies(target)` (APT_functions.R:671,793): target nodes + deterministic descendants +
    first stochastic descendants, topologically sorted;
  * `model$expandNodeNames`, `getBound(target, 'lower'|'upper')` (APT_functions.R:702-703).
"""
import ast
import re
import numpy as np

OPS = dict(END=0, CONST=1, LOOP=2, LOAD0=3, LOAD1=4, LOAD2=5, ADD=6, SUB=7, MUL=8, DIV=9, POW=10, NEG=11,
           EXP=12, LOG=13, ABS=14, SQRT=15, ILOGIT=16, LOGIT=17, PHI=18, ICLOGLOG=19, CLOGLOG=20, LOG1P=21,
           STEP=22, MIN=23, MAX=24, INPROD=25, SUM=26, LGAMMA=27,
           CEIL=28, FLOOR=29, ROUND=30, TRUNC=31, LT=32, GT=33, LE=34, GE=35, EQ=36, NE=37, OR=38, AND=39, NOT=40, SEL=41)
DISTS = dict(dnorm=0, dunif=1, dgamma=2, dbeta=3, dbinom=4, dpois=5, dmnorm=6, dexp=7, dcat=8, dmulti=9)
DISCRETE_DISTS = ("dcat", "dbinom", "dpois", "dmulti")  # model$isDiscrete(target), APT_functions.R:926
DREC, MAXARGS, ARGREC = 80, 6, 8
R_KIND, R_DIST, R_NLOOP, R_LO0, R_HI0, R_VAR, R_LNDIM, R_LIDX0, R_LRDIM, R_LRLO, R_LRHI, R_NARGS, R_ARGS = \
    0, 1, 2, 3, 6, 9, 10, 11, 13, 14, 15, 16, 17
FUNC1 = dict(exp="EXP", log="LOG", abs="ABS", sqrt="SQRT", ilogit="ILOGIT", expit="ILOGIT", logit="LOGIT",
             phi="PHI", iprobit="PHI", icloglog="ICLOGLOG", cloglog="CLOGLOG", log1p="LOG1P", step="STEP",
             lgamma="LGAMMA", loggam="LGAMMA")
INVLINK = dict(log="exp", logit="ilogit", probit="iprobit", cloglog="icloglog")
# what the bodies of user-supplied deterministic functions add (demo/APT_demo.R:42-71)
FUNC1.update(ceiling="CEIL", floor="FLOOR", round="ROUND", trunc="TRUNC", __not="NOT")
FUNC2 = dict(__lt="LT", __gt="GT", __le="LE", __ge="GE", __eq="EQ", __ne="NE", __or="OR", __and="AND")


# --------------------------------------------------------------------------- parsing
class Decl:
    def __init__(self, loops, lhs, stoch, rhs, link=None):
        self.loops, self.lhs, self.stoch, self.rhs, self.link = loops, lhs, stoch, rhs, link


def _split_statements(text):
    """Yield ('for', header, body_text) | ('stmt', text)."""
    i, n = 0, len(text)
    while i < n:
        while i < n and text[i] in " \t\r\n;":
            i += 1
        if i >= n:
            break
        m = re.match(r"for\s*\(", text[i:])
        if m:
            j = i + m.end()
            depth = 1
            while depth:
                depth += {"(": 1, ")": -1}.get(text[j], 0)
                j += 1
            header = text[i + m.end():j - 1]
            while text[j] in " \t\r\n":
                j += 1
            if text[j] != "{":  # a body of one statement (demo/APT_demo.R:90-91): up to the end of that statement
                k, depth = j, 0
                while k < n and not (text[k] in "\n;" and depth == 0 and not re.search(r"(<-|~|[-+*/^,])\s*$", text[j:k])):
                    depth += {"(": 1, "[": 1, ")": -1, "]": -1}.get(text[k], 0)
                    k += 1
                yield ("for", header, text[j:k])
                i = k
                continue
            k = j + 1
            depth = 1
            while depth:
                depth += {"{": 1, "}": -1}.get(text[k], 0)
                k += 1
            yield ("for", header, text[j + 1:k - 1])
            i = k
        else:
            j, depth = i, 0
            while j < n:
                c = text[j]
                if c in "([":
                    depth += 1
                elif c in ")]":
                    depth -= 1
                elif c in "\n;" and depth == 0:
                    # a trailing operator continues the statement on the next line
                    if c == "\n" and re.search(r"(<-|~|[-+*/^,])\s*$", text[i:j]):
                        j += 1
                        continue
                    break
                j += 1
            s = text[i:j].strip()
            if s:
                yield ("stmt", s)
            i = j


def _to_tree(node):
    if isinstance(node, ast.Expression):
        return _to_tree(node.body)
    if isinstance(node, ast.Constant):
        return ("num", float(node.value))
    if isinstance(node, ast.Name):
        if node.id == "Inf":
            return ("num", float("inf"))
        return ("var", node.id, [])
    if isinstance(node, ast.UnaryOp):
        if isinstance(node.op, ast.USub):
            return ("neg", _to_tree(node.operand))
        if isinstance(node.op, ast.UAdd):
            return _to_tree(node.operand)
        if isinstance(node.op, ast.Not):
            return ("call", "__not", [_to_tree(node.operand)], {})
    if isinstance(node, ast.BoolOp):
        f = "__or" if isinstance(node.op, ast.Or) else "__and"
        t = _to_tree(node.values[0])
        for v in node.values[1:]:
            t = ("call", f, [t, _to_tree(v)], {})
        return t
    if isinstance(node, ast.Compare) and len(node.ops) == 1:
        f = {ast.Lt: "__lt", ast.Gt: "__gt", ast.LtE: "__le", ast.GtE: "__ge", ast.Eq: "__eq", ast.NotEq: "__ne"}[type(node.ops[0])]
        return ("call", f, [_to_tree(node.left), _to_tree(node.comparators[0])], {})
    if isinstance(node, ast.BinOp):
        op = {ast.Add: "+", ast.Sub: "-", ast.Mult: "*", ast.Div: "/", ast.Pow: "^"}[type(node.op)]
        return ("bin", op, _to_tree(node.left), _to_tree(node.right))
    if isinstance(node, ast.Subscript):
        sl = node.slice
        parts = list(sl.elts) if isinstance(sl, ast.Tuple) else [sl]
        idx = []
        for p in parts:
            if isinstance(p, ast.Slice):
                idx.append(("range", _to_tree(p.lower), _to_tree(p.upper)))
            else:
                idx.append(_to_tree(p))
        return ("var", node.value.id, idx)
    if isinstance(node, ast.Call):
        args = [_to_tree(a) for a in node.args]
        kw = {k.arg: _to_tree(k.value) for k in node.keywords}
        return ("call", node.func.id, args, kw)
    raise ValueError("unsupported BUGS syntax: " + ast.dump(node))


def _parse_expr(s):
    s = s.strip().replace("^", "**")
    s = re.sub(r"\bTRUE\b", "1", s)
    s = re.sub(r"\bFALSE\b", "0", s)
    # R's logical operators bind looser than comparisons (Python's | and & bind tighter): spell them as or / and / not
    s = re.sub(r"\|\|?", " or ", s)
    s = re.sub(r"&&?", " and ", s)
    s = re.sub(r"!(?!=)", " not ", s)
    s = re.sub(r"\bin\b", " in_ ", s) if False else s
    return _to_tree(ast.parse(s, mode="eval"))


def parse_bugs(code):
    code = re.sub(r"#[^\n]*", "", code).strip()
    if code.startswith("{"):
        code = code[1:code.rindex("}")]
    decls = []

    def walk(text, loops):
        for item in _split_statements(text):
            if item[0] == "for":
                m = re.match(r"\s*(\w+)\s+in\s+(.*)$", item[1], re.S)
                var, rng = m.group(1), m.group(2)
                depth = 0
                for pos, ch in enumerate(rng):
                    depth += {"(": 1, "[": 1, ")": -1, "]": -1}.get(ch, 0)
                    if ch == ":" and depth == 0:
                        break
                walk(item[2], loops + [(var, _parse_expr(rng[:pos]), _parse_expr(rng[pos + 1:]))])
            else:
                s = item[1]
                depth, pos, kind = 0, -1, None
                for p, ch in enumerate(s):
                    depth += {"(": 1, "[": 1, ")": -1, "]": -1}.get(ch, 0)
                    if depth == 0 and ch == "~":
                        pos, kind = p, "~"
                        break
                    if depth == 0 and s[p:p + 2] == "<-":
                        pos, kind = p, "<-"
                        break
                assert kind, "cannot parse statement: " + s
                lhs = _parse_expr(s[:pos])
                rhs = _parse_expr(s[pos + len(kind):])
                link = None
                if lhs[0] == "call":
                    link = lhs[1]
                    lhs = lhs[2][0]
                decls.append(Decl(list(loops), lhs, kind == "~", rhs, link))

    walk(code, [])
    return decls


# --------------------------------------------------------------------------- user-supplied deterministic functions
def _balanced(text, i, open_, close):
    """index just behind the bracket that closes text[i] == open_."""
    depth, j = 0, i
    while True:
        depth += {open_: 1, close: -1}.get(text[j], 0)
        j += 1
        if depth == 0:
            return j


def _fn_block(text, i):
    """statements of the block (or single statement) that starts at text[i:]; returns (list, index behind it)."""
    while text[i] in " \t\r\n;":
        i += 1
    if text[i] == "{":
        j = _balanced(text, i, "{", "}")
        return _fn_statements(text[i + 1:j - 1]), j
    st, j = _fn_one(text, i)
    return ([st] if st else []), j


def _fn_one(text, i):
    n = len(text)
    m = re.match(r"if\s*\(", text[i:])
    if m:
        j = _balanced(text, i + m.end() - 1, "(", ")")
        cond = _parse_expr(text[i + m.end():j - 1])
        then, j = _fn_block(text, j)
        k = j
        while k < n and text[k] in " \t\r\n":
            k += 1
        els = []
        if text[k:k + 4] == "else" and not (text[k + 4:k + 5].isalnum() or text[k + 4:k + 5] == "_"):
            els, j = _fn_block(text, k + 4)
        return ("if", cond, then, els), j
    j, depth = i, 0
    while j < n and not (text[j] in "\n;" and depth == 0):
        depth += {"(": 1, "[": 1, ")": -1, "]": -1}.get(text[j], 0)
        j += 1
    st = text[i:j].strip()
    if not st or st.startswith("returnType"):
        return None, j
    m = re.match(r"return\s*\((.*)\)$", st, re.S)
    if m:
        return ("return", _parse_expr(m.group(1))), j
    m = re.match(r"(\w+)\s*(<-|=)(?!=)(.*)$", st, re.S)
    assert m, "unsupported statement in a user-supplied function: " + st
    return ("assign", m.group(1), _parse_expr(m.group(3))), j


def _fn_statements(text):
    out, i, n = [], 0, len(text)
    while True:
        while i < n and text[i] in " \t\r\n;":
            i += 1
        if i >= n:
            return out
        st, i = _fn_one(text, i)
        if st:
            out.append(st)


class UserFunction:
    """nimbleFunction(run = function(a = double(), v = double(1), ...) {...}) as a tree transformer: call(args) executes
    the body symbolically on the argument trees and returns ONE expression tree (conditionals become ("call", "__if", ..))."""

    def __init__(self, name, src):
        src = re.sub(r"#[^\n]*", "", src)
        i = src.index("function")
        i = src.index("(", i)
        j = _balanced(src, i, "(", ")")
        self.name, self.params = name, []
        depth, cur = 0, ""
        for ch in src[i + 1:j - 1] + ",":
            depth += {"(": 1, ")": -1}.get(ch, 0)
            if ch == "," and depth == 0:
                if cur.strip():
                    self.params.append(cur.split("=")[0].strip())
                cur = ""
            else:
                cur += ch
        k = src.index("{", j)
        self.body = _fn_statements(src[k + 1:_balanced(src, k, "{", "}") - 1])

    @staticmethod
    def _index(arg, idx):
        assert arg[0] == "var", "vector argument must be a variable or a slice of one"
        if not arg[2]:
            return ("var", arg[1], [idx])
        out, nr = [], 0
        for i in arg[2]:
            if i[0] == "range":
                nr += 1
                out.append(idx if i[1] == ("num", 1.0) else ("bin", "-", ("bin", "+", i[1], idx), ("num", 1.0)))
            else:
                out.append(i)
        assert nr == 1
        return ("var", arg[1], out)

    def _subst(self, t, env):
        k = t[0]
        if k == "var":
            if t[1] in env:
                if not t[2]:
                    return env[t[1]]
                assert len(t[2]) == 1 and t[2][0][0] != "range"
                return self._index(env[t[1]], self._subst(t[2][0], env))
            return ("var", t[1], [("range", self._subst(i[1], env), self._subst(i[2], env)) if i[0] == "range" else self._subst(i, env) for i in t[2]])
        if k == "neg":
            return ("neg", self._subst(t[1], env))
        if k == "bin":
            return ("bin", t[1], self._subst(t[2], env), self._subst(t[3], env))
        if k == "call":
            return ("call", t[1], [self._subst(a, env) for a in t[2]], {kk: self._subst(v, env) for kk, v in t[3].items()})
        return t

    def _exec(self, st, env):
        for i, s in enumerate(st):
            if s[0] == "assign":
                env[s[1]] = self._subst(s[2], env)
            elif s[0] == "return":
                return self._subst(s[1], env)
            else:
                c = self._subst(s[1], env)
                ea, eb = dict(env), dict(env)
                ra, rb = self._exec(s[2], ea), self._exec(s[3], eb)
                if ra is not None and rb is not None:
                    return ("call", "__if", [c, ra, rb], {})
                if ra is not None:
                    return ("call", "__if", [c, ra, self._exec(st[i + 1:], eb)], {})
                if rb is not None:
                    return ("call", "__if", [c, self._exec(st[i + 1:], ea), rb], {})
                for name in set(ea) | set(eb):
                    va, vb = ea.get(name), eb.get(name)
                    env[name] = va if vb is None or va == vb else vb if va is None else ("call", "__if", [c, va, vb], {})
        return None

    def call(self, args, kw):
        vals = list(args) + [None] * (len(self.params) - len(args))
        for k, v in kw.items():
            vals[self.params.index(k)] = v
        assert all(v is not None for v in vals), "missing argument of " + self.name
        r = self._exec(self.body, dict(zip(self.params, vals)))
        assert r is not None, self.name + "() does not return a value"
        return r


def expand_user_calls(t, fns, depth=0):
    assert depth < 16, "user-supplied functions nest too deeply"
    k = t[0]
    if k == "var":
        return ("var", t[1], [("range", expand_user_calls(i[1], fns, depth), expand_user_calls(i[2], fns, depth)) if i[0] == "range"
                              else expand_user_calls(i, fns, depth) for i in t[2]])
    if k == "neg":
        return ("neg", expand_user_calls(t[1], fns, depth))
    if k == "bin":
        return ("bin", t[1], expand_user_calls(t[2], fns, depth), expand_user_calls(t[3], fns, depth))
    if k == "call":
        args = [expand_user_calls(a, fns, depth) for a in t[2]]
        kw = {kk: expand_user_calls(v, fns, depth) for kk, v in t[3].items()}
        if t[1] in fns:
            return expand_user_calls(fns[t[1]].call(args, kw), fns, depth + 1)
        return ("call", t[1], args, kw)
    return t


# --------------------------------------------------------------------------- model
class Var:
    def __init__(self, name, role, dims):
        self.name, self.role, self.dims = name, role, tuple(int(d) for d in dims)
        self.off = -1

    @property
    def size(self):
        return int(np.prod(self.dims)) if self.dims else 1

    @property
    def nrow(self):
        return self.dims[0] if self.dims else 1


class OracleModel:
    """Concrete model: variables, canonical declarations in topological order, bytecode."""

    def __init__(self, code, constants=None, data=None, inits=None, functions=None):
        self.constants = {k: np.asarray(v, dtype=float) for k, v in (constants or {}).items()}
        self.data = {k: np.asarray(v, dtype=float) for k, v in (data or {}).items()}
        self.inits = {k: np.asarray(v, dtype=float) for k, v in (inits or {}).items()}
        self.decls = parse_bugs(code)
        if functions:
            fns = {k: UserFunction(k, v) for k, v in functions.items()}
            for d in self.decls:
                d.rhs = expand_user_calls(d.rhs, fns)
        self.code, self.consts = [], []
        self._canonicalise()
        self._build_vars()
        self._toposort()
        self._compile()

    # ---- helpers on trees
    def const_eval(self, t, env=None):
        """Evaluate a tree built from numbers, constants and loop variables (scalars or numpy arrays)."""
        env = env or {}
        k = t[0]
        if k == "num":
            return t[1]
        if k == "neg":
            return -self.const_eval(t[1], env)
        if k == "bin":
            a, b = self.const_eval(t[2], env), self.const_eval(t[3], env)
            return {"+": lambda: a + b, "-": lambda: a - b, "*": lambda: a * b, "/": lambda: a / b,
                    "^": lambda: a ** b}[t[1]]()
        if k == "var":
            if t[1] in env and not t[2]:
                return env[t[1]]
            if t[1] in self.constants:
                arr = self.constants[t[1]]
                if not t[2]:
                    return float(arr) if arr.ndim == 0 or arr.size == 1 else arr
                idx = tuple(np.asarray(self.const_eval(i, env)).astype(int) - 1 for i in t[2])
                return arr[idx]
            raise KeyError("not a constant: " + t[1])
        if k == "call" and t[1] in ("length",):
            return float(np.size(self.const_eval(t[2][0], env)))
        raise ValueError("not constant-evaluable: %r" % (t,))

    def _canonicalise(self):
        out = []
        for d in self.decls:
            loops = [(v, int(self.const_eval(lo)), int(self.const_eval(hi))) for v, lo, hi in d.loops]
            lhs = d.lhs
            if d.stoch:
                _, name, args, kw = d.rhs
                if name == "dbern":
                    name, args, kw = "dbinom", [kw.get("prob", args[0] if args else None), ("num", 1.0)], {}
                dist = name
                one = ("num", 1.0)
                if dist == "dnorm":
                    mean = kw.get("mean", args[0] if args else None)
                    if "sd" in kw:
                        sd = kw["sd"]
                    elif "var" in kw:
                        sd = ("call", "sqrt", [kw["var"]], {})
                    else:
                        tau = kw.get("tau", args[1] if len(args) > 1 else None)
                        sd = ("bin", "/", one, ("call", "sqrt", [tau], {}))
                    cargs = [mean, sd]
                elif dist == "dunif":
                    cargs = [kw.get("min", args[0] if args else None), kw.get("max", args[1] if len(args) > 1 else None)]
                elif dist == "dgamma":
                    shape = kw.get("shape", args[0] if args else None)
                    if "scale" in kw:
                        scale = kw["scale"]
                    else:
                        scale = ("bin", "/", one, kw.get("rate", args[1] if len(args) > 1 else None))
                    cargs = [shape, scale]
                elif dist == "dbeta":
                    cargs = [kw.get("shape1", args[0] if args else None), kw.get("shape2", args[1] if len(args) > 1 else None)]
                elif dist == "dbinom":
                    cargs = [kw.get("prob", args[0] if args else None), kw.get("size", args[1] if len(args) > 1 else None)]
                    if "prob" in kw and "size" not in kw and args:
                        cargs[1] = args[0]
                    if "size" in kw and "prob" not in kw and args:
                        cargs[0] = args[0]
                elif dist == "dpois":
                    cargs = [kw.get("lambda", args[0] if args else None)]
                elif dist == "dexp":
                    cargs = [kw["scale"]] if "scale" in kw else [("bin", "/", one, kw.get("rate", args[0] if args else None))]
                elif dist == "dcat":  # nimbleAPT/demo/APT_demo.R:84
                    cargs = [kw.get("prob", args[0] if args else None)]
                elif dist == "dmulti":  # target of sampler_RW_multinomial_tempered, APT_functions.R:1046
                    cargs = [kw.get("prob", args[0] if args else None), kw.get("size", args[1] if len(args) > 1 else None)]
                elif dist == "dmnorm":
                    mean = kw.get("mean", args[0] if args else None)
                    if "cholesky" in kw:
                        cargs = [mean, kw["cholesky"], kw.get("prec_param", one)]
                    else:
                        # dmnorm(mean, prec) / prec = / cov = with a constant matrix: chol() once, like nimble; the factor becomes
                        # a constant array of its own
                        cov = "cov" in kw
                        mt = kw["cov"] if cov else kw.get("prec", args[1] if len(args) > 1 else None)
                        assert mt is not None and mt[0] == "var", "dmnorm needs cholesky=, prec= or cov= (a constant matrix)"
                        M = np.asarray(self.constants[mt[1]], dtype=float)
                        if mt[2]:
                            r0, r1 = int(self.const_eval(mt[2][0][1])), int(self.const_eval(mt[2][0][2]))
                            c0, c1 = int(self.const_eval(mt[2][1][1])), int(self.const_eval(mt[2][1][2]))
                            M = M[r0 - 1:r1, c0 - 1:c1]
                        name = "chol__%d" % len(out)
                        self.constants[name] = np.linalg.cholesky(M).T.copy()  # upper factor, U'U = M
                        cargs = [mean, ("var", name, []), ("num", 0.0) if cov else one]
                else:
                    raise ValueError("unsupported distribution " + dist)
                out.append(dict(loops=loops, lhs=lhs, stoch=True, dist=dist, args=cargs))
            else:
                rhs = d.rhs
                if d.link:
                    rhs = ("call", INVLINK[d.link], [rhs], {})
                # scalarise an element-wise vector declaration  v[a:b] <- f(w[c:d])
                rdims = [k for k, i in enumerate(lhs[2]) if i[0] == "range"]
                if rdims:
                    assert len(rdims) == 1
                    lo = int(self.const_eval(lhs[2][rdims[0]][1])); hi = int(self.const_eval(lhs[2][rdims[0]][2]))
                    ev = "_e%d" % len(out)
                    loops = loops + [(ev, lo, hi)]
                    lidx = list(lhs[2]); lidx[rdims[0]] = ("var", ev, [])
                    lhs = ("var", lhs[1], lidx)

                    def scal(t):
                        if t[0] == "var":
                            idx = []
                            for i in t[2]:
                                if i[0] == "range":
                                    l2 = int(self.const_eval(i[1]))
                                    idx.append(("bin", "+", ("var", ev, []), ("num", float(l2 - lo))))
                                else:
                                    idx.append(scal(i))
                            return ("var", t[1], idx)
                        if t[0] == "bin":
                            return ("bin", t[1], scal(t[2]), scal(t[3]))
                        if t[0] == "neg":
                            return ("neg", scal(t[1]))
                        if t[0] == "call":
                            assert t[1] not in ("inprod", "sum"), "vector reductions inside vector declarations unsupported"
                            return ("call", t[1], [scal(a) for a in t[2]], t[3])
                        return t
                    rhs = scal(rhs)
                out.append(dict(loops=loops, lhs=lhs, stoch=False, dist=None, args=[rhs]))
        self.cdecls = out

    def _lhs_extent(self, d):
        """max index per LHS dimension over the declaration's loop ranges."""
        env = {v: np.array([lo, hi], dtype=float) for v, lo, hi in d["loops"]}
        ext = []
        for i in d["lhs"][2]:
            if i[0] == "range":
                ext.append(int(self.const_eval(i[2])))
            else:
                ext.append(int(np.max(self.const_eval(i, env))))
        return ext

    def _build_vars(self):
        self.vars = {}
        for d in self.cdecls:
            name = d["lhs"][1]
            ext = self._lhs_extent(d)
            if d["stoch"]:
                role = "data" if name in self.data else "param"
            else:
                role = "det"
            src = self.data.get(name) if role == "data" else self.inits.get(name)
            dims = tuple(src.shape) if src is not None and np.ndim(src) > 0 else tuple(ext)
            if name in self.vars:
                old = self.vars[name]
                assert old.role == role, "variable %s declared with mixed roles" % name
                dims = tuple(max(a, b) for a, b in zip(old.dims, dims)) if old.dims else dims
                old.dims = dims
            else:
                self.vars[name] = Var(name, role, dims)
        for name, arr in self.constants.items():
            if name not in self.vars and np.ndim(arr) > 0 and arr.size > 1:
                self.vars[name] = Var(name, "const", arr.shape)
        order = [v for v in self.vars.values() if v.role == "param"] + [v for v in self.vars.values() if v.role == "det"]
        off = 0
        for v in order:
            v.off = off; off += v.size
        self.nmut = off
        self.nparam = sum(v.size for v in self.vars.values() if v.role == "param")
        for v in self.vars.values():
            if v.role in ("data", "const"):
                v.off = off; off += v.size
        self.ntot = off
        self.arena0 = np.zeros(self.ntot)
        for v in self.vars.values():
            src = {"param": self.inits, "data": self.data, "const": self.constants}.get(v.role, {}).get(v.name)
            if src is not None:
                self.arena0[v.off:v.off + v.size] = np.asarray(src, dtype=float).reshape(-1, order="F")
        self.varlist = list(self.vars.values())
        self.varid = {v.name: i for i, v in enumerate(self.varlist)}

    def _reads(self, t, acc):
        if t[0] == "var":
            if t[1] in self.vars:
                acc.append(t)
            for i in t[2]:
                if i[0] == "range":
                    self._reads(i[1], acc); self._reads(i[2], acc)
                else:
                    self._reads(i, acc)
        elif t[0] == "bin":
            self._reads(t[2], acc); self._reads(t[3], acc)
        elif t[0] == "neg":
            self._reads(t[1], acc)
        elif t[0] == "call":
            for a in t[2]:
                self._reads(a, acc)
            for a in t[3].values():
                self._reads(a, acc)
        return acc

    def _toposort(self):
        n = len(self.cdecls)
        writers = {}
        for i, d in enumerate(self.cdecls):
            writers.setdefault(d["lhs"][1], []).append(i)
        deps = [set() for _ in range(n)]
        for i, d in enumerate(self.cdecls):
            for a in d["args"]:
                for r in self._reads(a, []):
                    if self.vars[r[1]].role in ("param", "det"):
                        deps[i].update(w for w in writers.get(r[1], []) if w != i)
        order, done = [], set()
        while len(order) < n:
            progressed = False
            for i in range(n):
                if i not in done and deps[i] <= done:
                    order.append(i); done.add(i); progressed = True
                    break
            assert progressed, "cyclic model"
        self.cdecls = [self.cdecls[i] for i in order]

    # ---- bytecode
    def _emit(self, op, a=0):
        self.code += [OPS[op], int(a)]

    def _const(self, v):
        self.consts.append(float(v)); return len(self.consts) - 1

    def _slice_rec(self, t, loopnames, subst):
        """arg record (kind 1) for a vector slice var[.., lo:hi, ..]; returns (record, other-index tree or None)."""
        v = self.vars[t[1]]
        idx = t[2]
        if not idx:  # whole vector
            return [1, self.varid[t[1]], 1, 0, 1, v.size, -1, 0], None
        rd = [k for k, i in enumerate(idx) if i[0] == "range"]
        assert len(rd) == 1, "exactly one range expected in a vector slice"
        lo, hi = int(self.const_eval(idx[rd[0]][1])), int(self.const_eval(idx[rd[0]][2]))
        other = None
        if len(idx) == 2:
            other = idx[1 - rd[0]]
        return [1, self.varid[t[1]], len(idx), rd[0], lo, hi, 0 if other is not None else -1, 0], other

    def _compile_expr(self, t, loopnames, subst=None):
        """append RPN for tree t (no END)."""
        subst = subst or {}
        k = t[0]
        if k == "num":
            self._emit("CONST", self._const(t[1]))
        elif k == "neg":
            self._compile_expr(t[1], loopnames, subst); self._emit("NEG")
        elif k == "bin":
            self._compile_expr(t[2], loopnames, subst); self._compile_expr(t[3], loopnames, subst)
            self._emit({"+": "ADD", "-": "SUB", "*": "MUL", "/": "DIV", "^": "POW"}[t[1]])
        elif k == "var":
            name = t[1]
            if name in subst and not t[2]:
                self._emit("CONST", self._const(subst[name]))
            elif name in loopnames and not t[2]:
                self._emit("LOOP", loopnames.index(name))
            elif name in self.vars:
                v = self.vars[name]
                assert all(i[0] != "range" for i in t[2]), "vector slice used where a scalar is required: " + name
                for dim, i in enumerate(t[2]):
                    self._compile_expr(i, loopnames, subst)
                    if any(self.vars[r[1]].role in ("param", "det") for r in self._reads(i, [])):
                        # index computed from a sampled node (demo: rfx_theta[k]): clamped into the array so that a
                        # state the prior excludes (k outside 1..K has logProb -Inf) cannot read out of bounds
                        self._emit("CONST", self._const(float(v.dims[dim]))); self._emit("MIN")
                        self._emit("CONST", self._const(1.0)); self._emit("MAX")
                if len(t[2]) == 0:
                    assert v.size == 1, "whole-array use of %s where a scalar is required" % name
                self._emit("LOAD%d" % len(t[2]), self.varid[name])
            elif name in self.constants:
                self._emit("CONST", self._const(float(self.const_eval(t))))
            else:
                raise KeyError("unknown variable " + name)
        elif k == "call":
            f = t[1]
            if f in FUNC1:
                self._compile_expr(t[2][0], loopnames, subst); self._emit(FUNC1[f])
            elif f in FUNC2:
                self._compile_expr(t[2][0], loopnames, subst); self._compile_expr(t[2][1], loopnames, subst); self._emit(FUNC2[f])
            elif f == "__if":  # evaluated eagerly on both sides (indices from sampled nodes are clamped), then selected
                for a in t[2]:
                    self._compile_expr(a, loopnames, subst)
                self._emit("SEL")
            elif f in ("pow",):
                self._compile_expr(t[2][0], loopnames, subst); self._compile_expr(t[2][1], loopnames, subst); self._emit("POW")
            elif f in ("min", "max"):
                self._compile_expr(t[2][0], loopnames, subst); self._compile_expr(t[2][1], loopnames, subst); self._emit(f.upper())
            elif f in ("inprod", "sum"):
                recs, others = [], []
                for a in t[2]:
                    r, o = self._slice_rec(a, loopnames, subst); recs.append(r); others.append(o)
                for o in others:
                    if o is not None:
                        self._compile_expr(o, loopnames, subst)
                # descriptor lives in the code stream behind a jump-free guard: [op, off] then records skipped
                # by storing them at the end of the code array later; here we reserve via a patch list.
                self._patches.append((len(self.code), recs))
                self._emit("INPROD" if f == "inprod" else "SUM", -1)
            else:
                raise ValueError("unsupported function " + f)
        else:
            raise ValueError(t)

    def compile_scalar(self, t, loopnames, subst=None):
        pc = len(self.code)
        self._compile_expr(t, loopnames, subst)
        self._emit("END")
        return pc

    def _compile(self):
        self._patches = []
        recs = []
        for d in self.cdecls:
            r = [0] * DREC
            loopnames = [v for v, _, _ in d["loops"]]
            assert len(loopnames) <= 3
            r[R_KIND] = 1 if d["stoch"] else 0
            r[R_DIST] = DISTS[d["dist"]] if d["stoch"] else 0
            r[R_NLOOP] = len(loopnames)
            for l, (v, lo, hi) in enumerate(d["loops"]):
                r[R_LO0 + l], r[R_HI0 + l] = lo, hi
            lhs = d["lhs"]
            r[R_VAR] = self.varid[lhs[1]]
            r[R_LNDIM] = len(lhs[2])
            r[R_LRDIM] = -1
            r[R_LIDX0] = r[R_LIDX0 + 1] = -1
            for k, i in enumerate(lhs[2]):
                if i[0] == "range":
                    r[R_LRDIM], r[R_LRLO], r[R_LRHI] = k, int(self.const_eval(i[1])), int(self.const_eval(i[2]))
                else:
                    r[R_LIDX0 + k] = self.compile_scalar(i, loopnames)
            r[R_NARGS] = len(d["args"])
            for a, t in enumerate(d["args"]):
                base = R_ARGS + a * ARGREC
                if d["stoch"] and d["dist"] in ("dmnorm", "dcat", "dmulti") and a == 0:
                    rec, other = self._slice_rec(t, loopnames, None)
                    if other is not None:
                        rec[6] = self.compile_scalar(other, loopnames)
                    r[base:base + ARGREC] = rec
                elif d["stoch"] and d["dist"] == "dmnorm" and a == 1:
                    idx = t[2]
                    v = self.vars[t[1]]
                    if idx:
                        lo1, hi1 = int(self.const_eval(idx[0][1])), int(self.const_eval(idx[0][2]))
                        lo2, hi2 = int(self.const_eval(idx[1][1])), int(self.const_eval(idx[1][2]))
                    else:
                        lo1, hi1, lo2, hi2 = 1, v.dims[0], 1, v.dims[1]
                    r[base:base + ARGREC] = [2, self.varid[t[1]], lo1, hi1, lo2, hi2, 0, 0]
                else:
                    r[base] = 0
                    r[base + 1] = self.compile_scalar(t, loopnames)
            recs.append(r)
        self.decl_recs = recs

    def finalize_code(self):
        """Place pending slice descriptors behind the code and patch the INPROD/SUM operands.  Call once, after
        every expression (declarations, sampler bounds) has been compiled and before handing code to C."""
        for pos, rr in self._patches:
            self.code[pos + 1] = len(self.code)
            for rec in rr:
                self.code += rec
            if len(rr) == 1:
                self.code += [0] * ARGREC
        self._patches = []

    # ---- graph queries
    def _inst_env(self, d):
        """dict loopvar -> array over all instances (first loop fastest), and the instance count."""
        sizes = [hi - lo + 1 for _, lo, hi in d["loops"]]
        n = int(np.prod(sizes)) if sizes else 1
        env, rem = {}, np.arange(n)
        for (v, lo, hi), s in zip(d["loops"], sizes):
            env[v] = (lo + rem % s).astype(float)
            rem = rem // s
        return env, n

    def _elem_index(self, t, env, n):
        """flattened 0-based element indices (array over instances) addressed by a non-range var reference;
        for range references returns (lo_index_array, count, stride)."""
        v = self.vars[t[1]]
        idx = t[2]
        if not idx:
            return np.zeros(n, dtype=np.int64), v.size, 1
        base = np.zeros(n, dtype=np.int64)
        cnt, stride_r = 1, 1
        stride = 1
        for k, i in enumerate(idx):
            if i[0] == "range":
                lo, hi = int(self.const_eval(i[1])), int(self.const_eval(i[2]))
                base = base + (lo - 1) * stride
                cnt, stride_r = hi - lo + 1, stride
            else:
                try:
                    val = np.asarray(self.const_eval(i, env))
                except (KeyError, ValueError):
                    # index computed from a sampled node: only constant arrays may be indexed this way, and those
                    # are never looked up here (the sampled node inside the index is itself one of the reads)
                    raise ValueError("element of a non-constant array selected by a sampled node: %r" % (t,))
                base = base + (np.broadcast_to(val, (n,)).astype(np.int64) - 1) * stride
            stride *= v.dims[k]
        return base, cnt, stride_r

    def lhs_elements(self, d, env, n):
        return self._elem_index(d["lhs"], env, n)

    def expand_target(self, target):
        """'beta[1:50]' | 'centroids[1]' | 'theta' | ['a','b[2]'] -> list of arena indices (PARAM elements)."""
        if isinstance(target, (list, tuple)):
            out = []
            for t in target:
                out += self.expand_target(t)
            return out
        t = _parse_expr(target)
        v = self.vars[t[1]]
        base, cnt, stride = self._elem_index(t, {}, 1)
        if not t[2]:
            return [v.off + e for e in range(v.size)]
        return [v.off + int(base[0]) + e * stride for e in range(cnt)]

    def dependencies(self, target_idx):
        """getDependencies(target): (calc_decl, calc_inst, tn_decl, tn_inst) in topological order."""
        frontier = np.zeros(self.nmut, dtype=bool)
        frontier[np.asarray(target_idx, dtype=int)] = True
        tgt = frontier.copy()
        cd, ci, td, ti = [], [], [], []
        for di, d in enumerate(self.cdecls):
            env, n = self._inst_env(d)
            hit = np.zeros(n, dtype=bool)
            v = self.vars[d["lhs"][1]]
            is_target_decl = False
            if d["stoch"] and v.role == "param":
                base, cnt, stride = self.lhs_elements(d, env, n)
                for e in range(cnt):
                    h = tgt[v.off + base + e * stride]
                    hit |= h
                is_target_decl = hit.any()
                own = hit.copy()
            for a in d["args"]:
                for r in self._reads(a, []):
                    rv = self.vars[r[1]]
                    if rv.role not in ("param", "det"):
                        continue
                    base, cnt, stride = self._elem_index(r, env, n)
                    for e in range(cnt):
                        hit |= frontier[rv.off + base + e * stride]
            if not hit.any():
                continue
            insts = np.nonzero(hit)[0]
            cd += [di] * len(insts); ci += insts.tolist()
            if is_target_decl:
                o = np.nonzero(own)[0]
                td += [di] * len(o); ti += o.tolist()
            if not d["stoch"]:
                base, cnt, stride = self.lhs_elements(d, env, n)
                for e in range(cnt):
                    frontier[v.off + base[insts] + e * stride] = True
        return cd, ci, td, ti

    def target_bounds(self, target_idx):
        """getBound(target, 'lower'/'upper') for a scalar target: trees with loop variables substituted."""
        for d in self.cdecls:
            v = self.vars[d["lhs"][1]]
            if not (d["stoch"] and v.role == "param"):
                continue
            env, n = self._inst_env(d)
            base, cnt, stride = self.lhs_elements(d, env, n)
            for e in range(cnt):
                w = np.nonzero(v.off + base + e * stride == target_idx)[0]
                if len(w):
                    subst = {k: float(a[w[0]]) for k, a in env.items()}
                    inf = float("inf")
                    if d["dist"] == "dunif":
                        return d["args"][0], d["args"][1], subst
                    if d["dist"] in ("dgamma", "dexp", "dpois", "dbinom"):
                        return ("num", 0.0), ("num", inf), subst
                    if d["dist"] == "dbeta":
                        return ("num", 0.0), ("num", 1.0), subst
                    return ("num", -inf), ("num", inf), subst
        raise KeyError("target has no stochastic declaration")

    def target_dist(self, target_idx):
        """model$getDistribution(target) of the node that owns this element (APT_functions.R:1046)."""
        for d in self.cdecls:
            v = self.vars[d["lhs"][1]]
            if not (d["stoch"] and v.role == "param"):
                continue
            env, n = self._inst_env(d)
            base, cnt, stride = self.lhs_elements(d, env, n)
            for e in range(cnt):
                if np.any(v.off + base + e * stride == target_idx):
                    return d["dist"]
        raise KeyError("target has no stochastic declaration")

    def target_nodes(self, target_idx):
        """unique(model$expandNodeNames(target)): the (declaration, instance) pairs owning the target elements."""
        nodes = set()
        for di, d in enumerate(self.cdecls):
            v = self.vars[d["lhs"][1]]
            if not (d["stoch"] and v.role == "param"):
                continue
            env, n = self._inst_env(d)
            base, cnt, stride = self.lhs_elements(d, env, n)
            for e in range(cnt):
                for t in target_idx:
                    for w in np.nonzero(v.off + base + e * stride == t)[0]:
                        nodes.add((di, int(w)))
        return nodes

    def is_discrete(self, target_idx):
        """model$isDiscrete(target) for a scalar target (APT_functions.R:926)."""
        for d in self.cdecls:
            v = self.vars[d["lhs"][1]]
            if not (d["stoch"] and v.role == "param"):
                continue
            env, n = self._inst_env(d)
            base, cnt, stride = self.lhs_elements(d, env, n)
            for e in range(cnt):
                if np.any(v.off + base + e * stride == target_idx):
                    return d["dist"] in DISCRETE_DISTS
        raise KeyError("target has no stochastic declaration")

    def data_node_lp_indices(self):
        """dataNodes <- model$getNodeNames(dataOnly = TRUE) (APT_functions.R:323) as indices into the C oracle's
        logProb arena (stochastic declarations in topological order, instances in loop order)."""
        out, off = [], 0
        for d in self.cdecls:
            if not d["stoch"]:
                continue
            env, n = self._inst_env(d)
            if self.vars[d["lhs"][1]].role == "data":
                out += list(range(off, off + n))
            off += n
        return out

    def check_waic_monitors(self, monitors):
        """mcmc_checkWAICmonitors (APT_functions.R:1313-1341): every parameter of a data node must be monitored or
        be downstream of monitored nodes.  Variable-level, like the reference."""
        mon = {_parse_expr(m)[1] for m in monitors if not m.startswith("logProb_")}
        stoch = [d for d in self.cdecls if d["stoch"]]

        def parents(d):  # stochastic parent variables of a declaration, through deterministic nodes
            seen, todo, out = set(), [d], set()
            while todo:
                q = todo.pop()
                for a in q["args"]:
                    for r in self._reads(a, []):
                        rv = self.vars[r[1]]
                        if rv.role == "param":
                            out.add(rv.name)
                        elif rv.role == "det" and rv.name not in seen:
                            seen.add(rv.name)
                            todo += [x for x in self.cdecls if not x["stoch"] and x["lhs"][1] == rv.name]
            return out

        par = [(d, parents(d)) for d in stoch]
        top = {d["lhs"][1] for d, p in par if not p and self.vars[d["lhs"][1]].role == "param"}
        this = {v for v in top if v not in mon}
        visited = set()
        while this:
            visited |= this
            nxt = [d for d, p in par if p & this]
            bad = sorted({d["lhs"][1] for d in nxt if self.vars[d["lhs"][1]].role == "data"})
            if bad:
                raise ValueError("In order for a valid WAIC calculation, all parameters of data nodes in the model must be "
                                 "monitored, or be downstream from monitored nodes. Currently, the following data nodes have "
                                 "un-monitored upstream parameters: " + ", ".join(bad))
            this = {d["lhs"][1] for d in nxt} - mon - visited
        # model$simulate(paramDeps) (APT_functions.R:327,618) would draw unmonitored stochastic nodes downstream of
        # monitored ones from their prior: not restated (the engine refuses the configuration as well)
        down, grew = set(), True
        while grew:
            grew = False
            for d, p in par:
                ln = d["lhs"][1]
                if self.vars[ln].role != "param" or ln in mon or ln in down:
                    continue
                if p & (mon | down):
                    down.add(ln); grew = True
        if down:
            raise NotImplementedError("unmonitored node %s is downstream of monitored nodes" % sorted(down)[0])

    def _logprob_monitor(self, nm):
        """'logProb_y' | 'logProb_y[3]' | 'logProb_y[2:5]' -> codes -(lp index)-1 in the C oracle's logProb arena
        (stochastic declarations in topological order, instances in loop order) and column labels."""
        t = _parse_expr(nm[len("logProb_"):])
        v = self.vars[t[1]]
        base, cnt, stride = self._elem_index(t, {}, 1)
        want = set(range(v.size)) if not t[2] else {int(base[0]) + e * stride for e in range(cnt)}
        out, labels, off = [], [], 0
        for d in self.cdecls:
            if not d["stoch"]:
                continue
            env, n = self._inst_env(d)
            if d["lhs"][1] == t[1]:
                b, c, s = self.lhs_elements(d, env, n)
                for inst in range(n):
                    e0 = int(b[inst])  # a multivariate node is named by its first element
                    if e0 in want:
                        out.append(-(off + inst) - 1)
                        if not v.dims:
                            labels.append("logProb_" + v.name)
                        elif len(v.dims) == 1:
                            labels.append("logProb_%s[%d]" % (v.name, e0 + 1))
                        else:
                            labels.append("logProb_%s[%d, %d]" % (v.name, e0 % v.dims[0] + 1, e0 // v.dims[0] + 1))
            off += n
        assert out, "monitor %s matches no stochastic node" % nm
        return out, labels

    def monitor_indices(self, names):
        out, labels = [], []
        for nm in names:
            if nm.startswith("logProb_"):  # nimble's logProb_<var> model variables (demo/APT_demo.R:180-181)
                o, l = self._logprob_monitor(nm)
                out += o; labels += l
                continue
            t = _parse_expr(nm)
            v = self.vars[t[1]]
            idxs = self.expand_target(nm)
            for a in idxs:
                e = a - v.off
                if len(v.dims) <= 1:
                    labels.append("%s[%d]" % (v.name, e + 1) if v.size > 1 else v.name)
                else:
                    labels.append("%s[%d, %d]" % (v.name, e % v.dims[0] + 1, e // v.dims[0] + 1))
            out += idxs
        return out, labels
