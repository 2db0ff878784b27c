"""A small reader for the Julia `Expr` text NeuralPDE's code generator produces, and the Python twin of the `lower`
pass of ext/NeuralPDEB200Ext.jl (same functions, same order, same IR) -- so that the mapping

    generated loss function (src/discretize.jl:28-152, grammar of src/symbolic_utilities.jl:132-331)  ->  residual IR

is exercised by tests even though no Julia runs in this image.  Test infrastructure, not product code.

AST (mirrors Julia's Expr heads):  Sym(name) | int | float | ("call", f, args) | (".", f, args) broadcast call |
("ref", a, idx) | ("field", a, name) | ("tuple", items) | ("vect", items) | ("=", lhs, rhs) | ("block", stmts) |
("let", binds, body) | ("->", args, body) | ("range", a, b) | Sym(":")
"""
import math
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple


@dataclass(frozen=True)
class Sym:
    name: str


# ---- tokenizer ---------------------------------------------------------------------------------------------------------
_TOK = re.compile(r"""
    (?P<ws>\s+) |
    (?P<num>\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+(?:[eE][-+]?\d+)?) |
    (?P<id>[^\W\d][\w!]*) |
    (?P<op>->|\.\^|\.\*|\./|\.\+|\.-|[-+*/^=,:()\[\].])
""", re.X | re.U)


def tokenize(text: str) -> List[Tuple[str, str]]:
    out, i = [], 0
    while i < len(text):
        m = _TOK.match(text, i)
        if not m:
            raise SyntaxError("cannot tokenize at %r" % text[i:i + 20])
        i = m.end()
        if m.lastgroup == "ws":
            if "\n" in m.group():
                out.append(("nl", "\n"))
            continue
        out.append((m.lastgroup, m.group()))
    return out


_OPS = {"+", "-", "*", "/", "^"}


class Parser:
    def __init__(self, text: str):
        self.t = tokenize(text)
        self.i = 0

    def peek(self, skip_nl=True):
        j = self.i
        while skip_nl and j < len(self.t) and self.t[j][0] == "nl":
            j += 1
        return self.t[j] if j < len(self.t) else ("eof", "")

    def next(self, skip_nl=True):
        while skip_nl and self.i < len(self.t) and self.t[self.i][0] == "nl":
            self.i += 1
        tok = self.t[self.i] if self.i < len(self.t) else ("eof", "")
        self.i += 1
        return tok

    def expect(self, val):
        tok = self.next()
        if tok[1] != val:
            raise SyntaxError("expected %r, got %r" % (val, tok[1]))

    # statement := expr [ '=' expr ] | expr '->' expr
    def statement(self):
        lhs = self.expr()
        k = self.peek(skip_nl=False)
        if k[1] == "=":
            self.next()
            return ("=", lhs, self.expr())
        if k[1] == "->":
            self.next()
            return ("->", lhs, self.statement())
        return lhs

    # expr := unary { ('.-' | '-' | '.+' | '+') unary }      (the generated text only uses `.-` at the top of the loss)
    def expr(self):
        e = self.unary()
        while self.peek(skip_nl=False)[1] in (".-", ".+"):
            op = self.next()[1]
            e = ("call", Sym(op), [e, self.unary()])
        return e

    def unary(self):
        if self.peek()[1] == "-":             # negative literal
            self.next()
            v = self.postfix()
            if isinstance(v, (int, float)):
                return -v
            return ("call", Sym("-"), [v])
        return self.postfix()

    def postfix(self):
        e = self.primary()
        while True:
            k = self.peek(skip_nl=False)
            if k[1] == "(":
                self.next()
                e = ("call", e, self.args(")"))
            elif k[1] == "[":
                self.next()
                e = ("ref", e, self.args("]"))
            elif k[1] == ".":
                self.next()
                if self.peek(skip_nl=False)[1] == "(":
                    self.next()
                    e = (".", e, self.args(")"))
                else:
                    e = ("field", e, self.next()[1])
            elif k[1] == ":" and isinstance(e, int):       # 1:1 inside an index
                self.next()
                e = ("range", e, self.primary())
            else:
                return e

    def args(self, close):
        items = []
        while self.peek()[1] != close:
            if self.peek()[1] == ":" :
                self.next()
                items.append(Sym(":"))
            else:
                items.append(self.statement() if close == ")" else self.expr())
            if self.peek()[1] == ",":
                self.next()
        self.expect(close)
        return items

    def primary(self):
        kind, val = self.next()
        if kind == "num":
            return float(val) if any(c in val for c in ".eE") else int(val)
        if kind == "id":
            if val == "begin":
                return ("block", self.block_until("end"))
            if val == "let":
                binds = self.statement()
                body = ("block", self.block_until("end"))
                return ("let", binds, body)
            return Sym(val)
        if val == "(":
            if self.peek()[1] in _OPS and self.t[self.i + 1][1] == ")":        # (+) (*) ... operator as a value
                op = self.next()[1]
                self.expect(")")
                return Sym(op)
            items = []
            trailing = False
            while self.peek()[1] != ")":
                items.append(self.statement())
                trailing = False
                if self.peek()[1] == ",":
                    self.next()
                    trailing = True
            self.expect(")")
            if len(items) == 1 and not trailing:
                return items[0]
            return ("tuple", items)
        if val == "[":
            return ("vect", self.args("]"))
        raise SyntaxError("unexpected token %r" % val)

    def block_until(self, word):
        stmts = []
        while True:
            k = self.peek()
            if k == ("id", word):
                self.next()
                return stmts
            if k[0] == "eof":
                raise SyntaxError("missing `%s`" % word)
            stmts.append(self.statement())


def parse(text: str):
    return Parser(text).statement()


# ---- lowering (twin of ext/NeuralPDEB200Ext.jl: `lower_loss_function`) ---------------------------------------------------
OP = {"+": "add", "-": "sub", "*": "mul", "/": "div", "^": "pow"}
UNARY = {"sin": "sin", "cos": "cos", "exp": "exp", "log": "log", "tanh": "tanh", "sqrt": "sqrt", "abs": "abs"}
CONSTS = {"π": math.pi, "pi": math.pi, "ℯ": math.e}


@dataclass
class Lowered:
    taps: List[Tuple[int, int, Tuple[int, ...]]] = field(default_factory=list)     # (net, order, dirs)
    prog: List[Tuple[str, int, int, float]] = field(default_factory=list)
    net_rows: Dict[int, List[int]] = field(default_factory=dict)
    dim: int = 0                                   # rows of the point matrix the engine sees
    const_rows: Dict[int, float] = field(default_factory=dict)   # rows the shim appends: constant bc coordinates


class _Builder:
    def __init__(self):
        self.out = Lowered()
        self.memo: Dict[tuple, int] = {}

    def emit(self, op, a=0, b=0, imm=0.0) -> int:
        key = (op, a, b, float(imm))
        if key not in self.memo:
            self.out.prog.append(key)
            self.memo[key] = len(self.out.prog) - 1
        return self.memo[key]

    def tap(self, net, order, dirs) -> int:
        key = (net, order, tuple(dirs))
        if key not in self.out.taps:
            self.out.taps.append(key)
        return self.emit("tap", self.out.taps.index(key))


def _stmts(node):
    """flatten nested blocks into a statement list"""
    if isinstance(node, tuple) and node[0] == "block":
        out = []
        for s in node[1]:
            out += _stmts(s)
        return out
    return [node]


def lower_loss_function(fn, depvars: List[str], eq_params: List[str] = (), param_estim: bool = False,
                        default_p: Optional[List[float]] = None) -> Lowered:
    """`fn`: AST of `(cord, θ, phi, derivative, integral, u, p) -> begin ... end` (build_symbolic_loss_function).
    depvars: dependent-variable names in dict_depvars order (network k serves depvars[k])."""
    assert fn[0] == "->"
    b = _Builder()
    env: Dict[str, tuple] = {}            # symbol -> ("coord", row) | ("const", v) | ("param", i) | ("net", k)
    loss = None
    for st in _stmts(fn[2]):
        if isinstance(st, tuple) and st[0] == "=" and isinstance(st[1], tuple) and st[1][0] == "tuple":
            _bind_tuple(st[1][1], st[2][1], env, depvars, param_estim, default_p)
        elif isinstance(st, tuple) and st[0] == "let":
            binds = st[1]
            _bind_tuple(binds[1][1] if binds[1][0] == "tuple" else [binds[1]],
                        binds[2][1] if binds[2][0] == "tuple" else [binds[2]], env, depvars, param_estim, default_p)
            n_coord = 1 + max([v[1] for v in env.values() if v[0] == "coord"], default=-1)
            for s2 in _stmts(st[2]):
                if isinstance(s2, tuple) and s2[0] == "=" and isinstance(s2[1], Sym) and s2[1].name.startswith("cord"):
                    k = int(s2[1].name[4:]) - 1                                  # cord<k> = vcat(vars...)   discretize.jl:111-116
                    assert s2[2][0] == "call" and s2[2][1] == Sym("vcat")
                    rows = []
                    for v in s2[2][2]:
                        kind, val = env[v.name]
                        if kind == "const":      # fill(value, ...): the network still needs a row; the shim appends it
                            hit = [r for r, c in b.out.const_rows.items() if c == val]
                            if not hit:
                                b.out.const_rows[n_coord + len(b.out.const_rows)] = val
                                hit = [n_coord + len(b.out.const_rows) - 1]
                            rows.append(hit[0])
                        else:
                            assert kind == "coord", "network input %s is not a point row" % v.name
                            rows.append(val)
                    b.out.net_rows[k] = rows
                    env[s2[1].name] = ("net", k)
                else:
                    loss = s2
    assert loss is not None, "no loss expression found"
    b.out.dim = 1 + max([v[1] for v in env.values() if v[0] == "coord"], default=-1) + len(b.out.const_rows)
    _lower(b, loss, env)
    return b.out


def _bind_tuple(lhs, rhs, env, depvars, param_estim, default_p):
    for l, r in zip(lhs, rhs):
        name = l.name
        if isinstance(r, tuple) and r[0] == "ref" and r[1] == Sym("cord"):            # cord[[i], :]      discretize.jl:126
            env[name] = ("coord", int(r[2][0][1][0]) - 1)
        elif isinstance(r, tuple) and r[0] == "call" and r[1] == Sym("fill"):           # fill(v, size(...)): constant bc coordinate
            env[name] = ("const", float(r[2][0]))
        elif isinstance(r, tuple) and r[0] == "field" and isinstance(r[1], tuple) and r[1][0] == "field" \
                and r[1][2] == "depvar":                                                 # θ.depvar.<name>   discretize.jl:58-80
            env[name] = ("net", depvars.index(r[2]))
        elif isinstance(r, tuple) and r[0] == "ref" and r[1] == Sym("phi"):             # phi[i]
            env[name] = ("net", int(r[2][0]) - 1)
        elif isinstance(r, tuple) and r[0] == "ref" and isinstance(r[1], tuple) and r[1][0] == "field" and r[1][2] == "p":
            rng = r[2][0]                                                                # θ.p[i:i]          discretize.jl:83-95
            env[name] = ("param", int(rng[1] if isinstance(rng, tuple) else rng) - 1)
        elif isinstance(r, tuple) and r[0] == "call" and isinstance(r[1], tuple) and r[1][0] == "field" \
                and r[1][2] == "allowed_getindex":                                       # default_p[i]      discretize.jl:97-109
            env[name] = ("const", float(default_p[int(r[2][1]) - 1]))
        else:
            raise ValueError("unrecognised binding %s = %r" % (name, r))


def _net_of(arg, env) -> int:
    """network index from a cord<k> / phi<k> / θ<k> symbol (single-output: cord1 / phi / θ -> 0)"""
    if isinstance(arg, Sym):
        if arg.name in env and env[arg.name][0] == "net":
            return env[arg.name][1]
        m = re.search(r"(\d+)$", arg.name)
        return int(m.group(1)) - 1 if m else 0
    return 0


def _lower(b: _Builder, ex, env) -> int:
    if isinstance(ex, (int, float)):
        return b.emit("const", imm=float(ex))
    if isinstance(ex, Sym):
        if ex.name in CONSTS:
            return b.emit("const", imm=CONSTS[ex.name])
        kind, val = env[ex.name]
        if kind == "coord":
            return b.emit("coord", val)
        if kind == "const":
            return b.emit("const", imm=val)
        if kind == "param":
            return b.emit("param", val)
        raise ValueError("symbol %s cannot appear in an expression" % ex.name)
    head = ex[0]
    if head == "call" and ex[1] == Sym("u"):                       # u(cord_k, θ_k, phi_k)                  symbolic_utilities.jl:150-159
        return b.tap(_net_of(ex[2][0], env), 0, ())
    if head == "call" and ex[1] == Sym("derivative"):              # derivative(phi_k, u, cord_k, εs, order, θ_k)   :189-201
        _phi, _u, cord, eps, order, _th = ex[2]
        dirs = []
        for e in eps[1]:                                            # each ε vector is one-hot: its position is the direction
            nz = [i for i, v in enumerate(e[1]) if float(v) != 0.0]
            assert len(nz) == 1
            dirs.append(nz[0])
        assert len(dirs) == int(order)
        return b.tap(_net_of(cord, env), int(order), tuple(sorted(dirs)))
    if head in (".", "call"):
        f = ex[1].name.lstrip(".") if isinstance(ex[1], Sym) else None
        args = ex[2]
        if f in OP:
            if len(args) == 1 and f == "-":
                return b.emit("neg", _lower(b, args[0], env))
            if f == "^" and isinstance(args[1], int):
                return b.emit("powi", _lower(b, args[0], env), 0, float(args[1]))
            acc = _lower(b, args[0], env)
            for a in args[1:]:                                      # n-ary + and * fold to the left
                acc = b.emit(OP[f], acc, _lower(b, a, env))
            return acc
        if f in UNARY:
            return b.emit(UNARY[f], _lower(b, args[0], env))
    raise ValueError("expression outside the grammar of _transform_expression: %r" % (ex,))
