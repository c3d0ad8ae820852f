"""OpenQASM-2 subset parser + evaluator restating the reference interpreter.  TEST INFRASTRUCTURE.

The reference keeps its parser and evaluator in Haskell (they are out of scope for the
CUDA backend); GHC is absent here, so this module restates
them only far enough to turn .qasm text into the *primitive op stream* that reaches the
hot path, with the reference's observable quirks:

  grammar              src/Qubism/QASM/Parser.hs:184-335
  statement evaluator  src/Qubism/QASM/Simulation.hs:55-227
  register fusion      src/Qubism/QASM/ProgState.hs:137-166
  errors               src/Qubism/QASM/ProgState.hs:97-110 (RuntimeError)

PARITY UNPINNED: see oracle/__init__.py.

The evaluator is parametrised by a *backend* implementing the L2 seam (the functions of
Qubism.StateVec / Qubism.QGate the interpreter calls, ProgState.hs:152,155,225 and
Simulation.hs:81,85,100,121,142,154-155) with value semantics:

    mk(n) -> sv                      mkStateVec'
    tensor(a, b) -> sv               tensor
    dimension(sv) -> n               dimension
    apply_1q(sv, q, m) -> sv         onJust q m #> sv
    apply_range(sv, f, l, m) -> sv   onRange f l m #> sv
    apply_cnot(sv, c, t) -> sv       cnot c t #> sv
    measure_qubit(sv, q, r) -> (bit, sv)
    collapse(sv, q, b) -> sv
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field

from . import dense as _dense

# The IBM standard header every example includes.  The reference ships it as
# examples/qelib1.inc; everything in it expands to U and CX only.  Restated here (not
# copied) so the oracle does not read /root/reference at run time.
QELIB1 = """
gate u3(theta,phi,lambda) q { U(theta,phi,lambda) q; }
gate u2(phi,lambda) q { U(pi/2,phi,lambda) q; }
gate u1(lambda) q { U(0,0,lambda) q; }
gate cx c,t { CX c,t; }
gate id a { U(0,0,0) a; }
gate x a { u3(pi,0,pi) a; }
gate y a { u3(pi,pi/2,pi/2) a; }
gate z a { u1(pi) a; }
gate h a { u2(0,pi) a; }
gate s a { u1(pi/2) a; }
gate sdg a { u1(-pi/2) a; }
gate t a { u1(pi/4) a; }
gate tdg a { u1(-pi/4) a; }
gate rx(theta) a { u3(theta,-pi/2,pi/2) a; }
gate ry(theta) a { u3(theta,0,0) a; }
gate rz(phi) a { u1(phi) a; }
gate cz a,b { h b; cx a,b; h b; }
gate cy a,b { sdg b; cx a,b; s b; }
gate ch a,b { h b; sdg b; cx a,b; h b; t b; cx a,b; t b; h b; s b; x b; s a; }
gate ccx a,b,c { h c; cx b,c; tdg c; cx a,c; t c; cx b,c; tdg c; cx a,c; t b; t c; h c;
                 cx a,b; t a; tdg b; cx a,b; }
gate crz(lambda) a,b { u1(lambda/2) b; cx a,b; u1(-lambda/2) b; cx a,b; }
gate cu1(lambda) a,b { u1(lambda/2) a; cx a,b; u1(-lambda/2) b; cx a,b; u1(lambda/2) b; }
gate cu3(theta,phi,lambda) c,t { u1((lambda-phi)/2) t; cx c,t; u3(-theta/2,0,-(phi+lambda)/2) t;
                                 cx c,t; u3(theta/2,phi,0) t; }
"""

RWS = {"if", "barrier", "gate", "measure", "reset", "creg", "qreg", "pi", "sin", "cos", "tan",
       "exp", "ln", "sqrt", "U", "CX", "include"}  # Parser.hs:133-135


class ParseError(Exception):
    pass


class RuntimeErrorQ(Exception):
    """ProgState.hs:97-105 RuntimeError (line, message)."""

    def __init__(self, line, msg):
        super().__init__(f"ERROR on line {line}\n{msg}")
        self.line, self.msg = line, msg


_TOK = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<hdr>OPENQASM\s+2\.0\s*;)
  | (?P<float>\d+\.\d+(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)
  | (?P<nat>\d+)
  | (?P<id>[A-Za-z][A-Za-z0-9]*)
  | (?P<str>"[A-Za-z0-9./]*")
  | (?P<dump>:dump)
  | (?P<sym>->|==|[;,()\[\]{}+\-*/])
""", re.X | re.S)


def _lex(src: str):
    out, pos, line = [], 0, 1
    while pos < len(src):
        m = _TOK.match(src, pos)
        if not m:
            raise ParseError(f"line {line}: unexpected {src[pos:pos+10]!r}")
        kind = m.lastgroup
        text = m.group()
        if kind not in ("ws", "hdr"):
            out.append((kind, text, line))
        line += text.count("\n")
        pos = m.end()
    out.append(("eof", "", line))
    return out


class _Parser:
    """Recursive-descent version of Parser.hs:187-335.  AST nodes are tuples whose first
    element is the constructor name of QASM/AST.hs:20-67."""

    def __init__(self, src, idtable=None, include_resolver=None):
        self.toks = _lex(src)
        self.i = 0
        self.ids = idtable if idtable is not None else set()
        self.include_resolver = include_resolver

    def peek(self):
        return self.toks[self.i]

    def next(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def accept(self, text):
        if self.peek()[1] == text and self.peek()[0] in ("sym", "id"):
            return self.next()
        return None

    def expect(self, text):
        t = self.next()
        if t[1] != text:
            raise ParseError(f"line {t[2]}: expected {text!r}, got {t[1]!r}")
        return t

    def program(self):  # Parser.hs:187-189: sepEndBy1 stmt (";" | "}")
        stmts = []
        while self.peek()[0] != "eof":
            stmts.append(self.stmt())
            if not (self.accept(";") or self.accept("}")):
                if self.peek()[0] != "eof":
                    t = self.peek()
                    raise ParseError(f"line {t[2]}: expected ';', got {t[1]!r}")
        if not stmts:
            raise ParseError("empty program")
        return stmts

    def stmt(self):  # Parser.hs:191-198
        kind, text, line = self.peek()
        if text == "if" and kind == "id":
            node = self.cond()
        elif text in ("qreg", "creg") and kind == "id":
            self.next()
            name = self.new_ident()
            self.expect("[")
            size = self.natural()
            self.expect("]")
            node = ("QRegDecl" if text == "qreg" else "CRegDecl", name, size)
        elif text == "gate" and kind == "id":
            node = self.gate_decl()
        elif text == "include" and kind == "id":
            self.next()
            fname = self.next()[1].strip('"')
            src = self.include_resolver(fname) if self.include_resolver else None
            if src is None:
                raise ParseError(f"line {line}: could not include {fname}")
            sub = _Parser(src, self.ids, self.include_resolver)
            node = ("StmtList", sub.program())
        elif text in ("measure", "reset") and kind == "id":
            node = ("QOp", self.qop())
        else:
            node = ("UOp", self.uop())
        return ("PosInfo", line, node)

    def identifier(self):
        kind, text, line = self.next()
        if kind != "id" or text in RWS:
            raise ParseError(f"line {line}: expected identifier, got {text!r}")
        return text

    def new_ident(self):  # Parser.hs:148-152, 342-349
        line = self.peek()[2]
        i = self.identifier()
        if i in self.ids:
            raise ParseError(f"line {line}: Redeclaration of {i}")
        self.ids.add(i)
        return i

    def known_ident(self):  # Parser.hs:154-160
        line = self.peek()[2]
        i = self.identifier()
        if i not in self.ids:
            raise ParseError(f"line {line}: Undeclared identifier: {i}")
        return i

    def natural(self):
        kind, text, line = self.next()
        if kind != "nat":
            raise ParseError(f"line {line}: expected natural, got {text!r}")
        return int(text)

    def gate_decl(self):  # Parser.hs:209-223: formals shadow, table restored afterwards
        self.expect("gate")
        name = self.new_ident()
        saved = set(self.ids)
        params = []
        if self.accept("("):
            while not self.accept(")"):
                p = self.identifier()
                self.ids.add(p)
                params.append(p)
                self.accept(",")
        args = [self.identifier()]
        self.ids.add(args[-1])
        while self.accept(","):
            if self.peek()[1] == "{":
                break
            args.append(self.identifier())
            self.ids.add(args[-1])
        self.expect("{")
        body = []
        while self.peek()[1] != "}":
            body.append(self.uop())
            self.expect(";")
        # the closing "}" is consumed by ``program`` as a statement separator
        self.ids.clear()
        self.ids.update(saved)
        return ("GateDecl", name, params, args, body)

    def cond(self):  # Parser.hs:296-304
        self.expect("if")
        self.expect("(")
        i = self.known_ident()
        self.expect("==")
        n = self.natural()
        self.expect(")")
        return ("Cond", i, n, self.qop())

    def qop(self):  # Parser.hs:255-265
        if self.accept("measure"):
            src = self.argument()
            self.expect("->")
            tgt = self.argument()
            return ("Measure", src, tgt)
        if self.accept("reset"):
            return ("Reset", self.argument())
        return ("QUnitary", self.uop())

    def uop(self):  # Parser.hs:267-294
        kind, text, line = self.peek()
        if kind == "dump":
            self.next()
            return ("Dump",)
        if text == "U" and kind == "id":
            self.next()
            self.expect("(")
            p1 = self.expr()
            self.expect(",")
            p2 = self.expr()
            self.expect(",")
            p3 = self.expr()
            self.expect(")")
            return ("U", p1, p2, p3, self.argument())
        if text == "CX" and kind == "id":
            self.next()
            a1 = self.argument()
            self.expect(",")
            return ("CX", a1, self.argument())
        if text == "barrier" and kind == "id":
            self.next()
            return ("Barrier", self.arg_list())
        name = self.known_ident()
        params = []
        if self.accept("("):
            while not self.accept(")"):
                params.append(self.expr())
                self.accept(",")
        return ("Func", name, params, self.arg_list())

    def arg_list(self):  # list argument = sepEndBy argument comma
        args = []
        while self.peek()[0] == "id" and self.peek()[1] not in RWS:
            args.append(self.argument())
            if not self.accept(","):
                break
        return args

    def argument(self):  # Parser.hs:306-312
        name = self.known_ident()
        if self.accept("["):
            i = self.natural()
            self.expect("]")
            return ("ArgBit", name, i)
        return ("ArgReg", name)

    # expression precedence, Parser.hs:321-335: unary minus binds tightest, then the
    # function prefixes, then `pow`, then * /, then + -
    def expr(self):
        e = self.expr_mul()
        while self.peek()[1] in ("+", "-") and self.peek()[0] == "sym":
            op = self.next()[1]
            e = ("Binary", "Add" if op == "+" else "Sub", e, self.expr_mul())
        return e

    def expr_mul(self):
        e = self.expr_pow()
        while self.peek()[1] in ("*", "/") and self.peek()[0] == "sym":
            op = self.next()[1]
            e = ("Binary", "Mul" if op == "*" else "Div", e, self.expr_pow())
        return e

    def expr_pow(self):
        e = self.expr_fn()
        while self.peek()[1] == "pow" and self.peek()[0] == "id":
            self.next()
            e = ("Binary", "Pow", e, self.expr_fn())
        return e

    def expr_fn(self):
        kind, text, _ = self.peek()
        if kind == "id" and text in ("sin", "cos", "tan", "exp", "ln", "sqrt"):
            self.next()
            return ("Unary", text.capitalize(), self.expr_neg())
        return self.expr_neg()

    def expr_neg(self):
        if self.peek() [0] == "sym" and self.peek()[1] == "-":
            self.next()
            return ("Unary", "Neg", self.term())
        return self.term()

    def term(self):
        kind, text, line = self.next()
        if kind == "id" and text == "pi":
            return ("Pi",)
        if kind in ("float", "nat"):
            return ("Real", float(text))
        if kind == "sym" and text == "(":
            e = self.expr()
            self.expect(")")
            return e
        if kind == "id" and text not in RWS:
            if text not in self.ids:
                raise ParseError(f"line {line}: Undeclared identifier: {text}")
            return ("Ident", text)
        raise ParseError(f"line {line}: bad expression term {text!r}")


def default_include_resolver(fname: str):
    return QELIB1 if fname.endswith("qelib1.inc") else None


def parse(src: str, idtable=None, include_resolver=default_include_resolver):
    """parseOpenQASM (Parser.hs:61-79).  ``idtable`` persists across REPL lines."""
    return _Parser(src, idtable, include_resolver).program()


def eval_expr(e) -> float:
    """Simulation.hs:209-227.  NB ``pi`` is the truncated literal 3.14159265358979 (:211)."""
    k = e[0]
    if k == "Pi":
        return 3.14159265358979
    if k == "Real":
        return e[1]
    if k == "Ident":
        raise RuntimeError("undefined")  # Simulation.hs:212
    if k == "Binary":
        a, b = eval_expr(e[2]), eval_expr(e[3])
        return {"Add": lambda: a + b, "Sub": lambda: a - b, "Mul": lambda: a * b,
                "Div": lambda: a / b, "Pow": lambda: a ** b}[e[1]]()
    if k == "Unary":
        a = eval_expr(e[2])
        return {"Neg": lambda: -a, "Sin": lambda: math.sin(a), "Cos": lambda: math.cos(a),
                "Tan": lambda: math.tan(a), "Exp": lambda: math.exp(a), "Ln": lambda: math.log(a),
                "Sqrt": lambda: math.sqrt(a)}[e[1]]()
    raise ValueError(k)


@dataclass
class QReg:  # ProgState.hs:42-46
    target: str
    start: int
    size: int


@dataclass
class ProgState:  # ProgState.hs:63-69
    stVecs: dict = field(default_factory=dict)
    qregs: dict = field(default_factory=dict)
    cregs: dict = field(default_factory=dict)
    funcs: dict = field(default_factory=dict)
    pos: int = 0


class Evaluator:
    """runStmt and friends (Simulation.hs:55-207) over an abstract backend.

    ``draw()`` supplies the uniform variate of each measureQubit (StateVec.hs:123).
    ``ref_faithful`` keeps the reference's ``withIndex`` write-back bug (Simulation.hs:101:
    the result is stored under the QReg name, not the state-vector id -- harmless until two
    qregs have been fused, after which single-qubit gates on ``name[k]`` are silently lost)
    and the ``reset`` offset bugs (Simulation.hs:149-155).  ``trace`` (a list) receives the
    primitive ops that actually changed the live state: the op stream a backend must match.
    """

    def __init__(self, backend, draw, ref_faithful=True, trace=None):
        self.be = backend
        self.draw = draw
        self.ref_faithful = ref_faithful
        self.ps = ProgState()
        self.trace = trace

    # -- helpers, ProgState.hs:107-110,168-258
    def err(self, msg):
        raise RuntimeErrorQ(self.ps.pos, msg)

    def find(self, name, table):
        if name not in table:
            self.err(f"Undeclared identifier: {name}")
        return table[name]

    def _emit(self, *op):
        if self.trace is not None:
            self.trace.append(op)

    def run(self, prog):  # runProgram' (Simulation.hs:47-53)
        for s in prog:
            self.run_stmt(s)
        return self.ps

    def run_stmt(self, s):  # Simulation.hs:55-76
        k = s[0]
        if k == "PosInfo":
            self.ps.pos = s[1]
            self.run_stmt(s[2])
        elif k == "StmtList":
            for x in s[1]:
                self.run_stmt(x)
        elif k == "QRegDecl":  # addQReg, ProgState.hs:174-181,219-228
            name, size = s[1], s[2]
            if name in self.ps.qregs or name in self.ps.stVecs:
                self.err(f"Redeclaration of {name}")
            self.ps.qregs[name] = QReg(name, 0, size)
            self.ps.stVecs[name] = self.be.mk(size)
            self._emit("ALLOC", name, size)
        elif k == "CRegDecl":  # addCReg, ProgState.hs:191-197
            if s[1] in self.ps.cregs:
                self.err(f"Redeclaration of {s[1]}")
            self.ps.cregs[s[1]] = [0] * s[2]
        elif k == "GateDecl":
            self.ps.funcs[s[1]] = (s[2], s[3], s[4])
        elif k == "QOp":
            op = s[1]
            if op[0] == "QUnitary":
                self.run_stmt(("UOp", op[1]))
            elif op[0] == "Measure":
                self.observe(op[1], op[2])
            else:
                self.reset(op[1])
        elif k == "UOp":
            op = s[1]
            if op[0] == "U":
                g = _dense.unitary(eval_expr(op[1]), eval_expr(op[2]), eval_expr(op[3]))
                self.apply1(g, op[4])
            elif op[0] == "CX":
                self.cx(op[1], op[2])
            elif op[0] == "Func":
                self.custom_op(op[1], [eval_expr(e) for e in op[2]], op[3])
            # Barrier: no-op; Dump: trace only
        elif k == "Cond":  # Simulation.hs:73-76
            cr = self.find(s[1], self.ps.cregs)
            if _dense.crToNatural(cr) == s[2]:
                self.run_stmt(("QOp", s[3]))
        else:
            raise ValueError(k)

    def apply1(self, g, arg):  # (##>), Simulation.hs:79-85
        if arg[0] == "ArgBit":
            self.with_index(lambda sv, idx: self.be.apply_1q(sv, idx, g), arg[1], arg[2], ("U", g))
        else:
            qr = self.find(arg[1], self.ps.qregs)
            self.with_index2(lambda sv, i, j: self.be.apply_range(sv, i, j, g),
                             arg[1], 0, arg[1], qr.size - 1, ("URANGE", g))

    def with_index(self, f, qr, i, tag):  # Simulation.hs:88-101
        reg = self.find(qr, self.ps.qregs)
        sv = self.find(reg.target, self.ps.stVecs)
        idx = reg.start + i
        if not 0 <= idx < self.be.dimension(sv):
            raise IndexError("finite: index out of range")  # Haskell `error`, :99
        sv2 = f(sv, idx)
        dest = qr if self.ref_faithful else reg.target  # the write-back bug, :101
        self.ps.stVecs[dest] = sv2
        if dest == reg.target:
            self._emit(tag[0], reg.target, idx, tag[1])

    def with_index2(self, f, qr1, i, qr2, j, tag):  # Simulation.hs:104-122
        idsv = self.fuse(qr1, qr2)
        sv = self.find(idsv, self.ps.stVecs)
        s1 = self.find(qr1, self.ps.qregs).start
        s2 = self.find(qr2, self.ps.qregs).start
        n = self.be.dimension(sv)
        if not (0 <= s1 + i < n and 0 <= s2 + j < n):
            raise IndexError("finite: index out of range")
        self.ps.stVecs[idsv] = f(sv, s1 + i, s2 + j)
        self._emit(tag[0], idsv, s1 + i, s2 + j, *tag[1:])

    def fuse(self, qr1, qr2):  # fuseQRegs, ProgState.hs:137-166
        id1 = self.find(qr1, self.ps.qregs).target
        id2 = self.find(qr2, self.ps.qregs).target
        if id1 == id2:
            return id1
        sv1 = self.find(id1, self.ps.stVecs)
        sv2 = self.find(id2, self.ps.stVecs)
        new = id1 + "(x)" + id2
        self.ps.stVecs[new] = self.be.tensor(sv1, sv2)
        shift = self.be.dimension(sv1)
        for q in self.ps.qregs.values():
            if q.target == id1:
                q.target = new
            elif q.target == id2:
                q.target, q.start = new, q.start + shift
        del self.ps.stVecs[id1]
        del self.ps.stVecs[id2]
        self._emit("TENSOR", new, id1, id2)
        return new

    def observe(self, argq, argc):  # Simulation.hs:124-144
        def go(i):
            reg = self.find(argq[1], self.ps.qregs)
            sv = self.find(reg.target, self.ps.stVecs)
            k = i + reg.start
            r = self.draw()
            bit, sv2 = self.be.measure_qubit(sv, k, r)
            self.ps.stVecs[reg.target] = sv2
            self._emit("MEASURE", reg.target, k, r, bit, getattr(self.be, "last_pone", None))
            return bit

        if argq[0] == "ArgBit":
            bits = [go(argq[2])]
        else:
            s = self.find(argq[1], self.ps.qregs).size
            bits = [go(i) for i in range(s)]
        if argc[0] == "ArgBit":  # writeBit, ProgState.hs:209-217
            cr = self.find(argc[1], self.ps.cregs)
            if argc[2] < len(cr):
                cr[argc[2]] = bits[0]
            else:
                self.err(f"Index out of bounds when writing to {argc[1]}")
        else:  # writeCReg, ProgState.hs:199-207
            cr = self.find(argc[1], self.ps.cregs)
            if len(bits) == len(cr):
                self.ps.cregs[argc[1]] = list(bits)
            else:
                self.err(f"Mismatched size on overwrite of {argc[1]}")

    def reset(self, arg):  # Simulation.hs:146-156 (offset bugs kept when ref_faithful)
        reg = self.find(arg[1], self.ps.qregs)
        sv = self.find(reg.target, self.ps.stVecs)
        if arg[0] == "ArgBit":
            ks = [arg[2] if self.ref_faithful else reg.start + arg[2]]
        elif self.ref_faithful:
            ks = list(reversed(range(reg.start, reg.size)))  # foldr over [i..s-1]
        else:
            ks = list(reversed(range(reg.start, reg.start + reg.size)))
        for k in ks:
            sv = self.be.collapse(sv, k, 0)
            self._emit("COLLAPSE", reg.target, k, 0)
        self.ps.stVecs[reg.target] = sv

    def cx(self, a1, a2):  # Simulation.hs:158-173
        def one(q1, i, q2, j):
            self.with_index2(lambda sv, c, t: self.be.apply_cnot(sv, c, t), q1, i, q2, j, ("CX",))

        size = lambda q: self.find(q, self.ps.qregs).size
        if a1[0] == "ArgBit" and a2[0] == "ArgBit":
            one(a1[1], a1[2], a2[1], a2[2])
        elif a1[0] == "ArgBit":
            for j in range(size(a2[1])):
                one(a1[1], a1[2], a2[1], j)
        elif a2[0] == "ArgBit":
            for i in range(size(a1[1])):
                one(a1[1], i, a2[1], a2[2])
        else:
            s1, s2 = size(a1[1]), size(a2[1])
            if s1 != s2:
                self.err(f"QRegs of different sizes supplied to CX: {a1[1]} {a2[1]}")
            for i in range(s1):
                one(a1[1], i, a2[1], i)

    def custom_op(self, name, params, args):  # Simulation.hs:175-207
        ps, as_, uops = self.find(name, self.ps.funcs)
        arg_binds = dict(zip(as_, args))
        par_binds = dict(zip(ps, params))

        def bind_e(e):
            if e[0] == "Binary":
                return ("Binary", e[1], bind_e(e[2]), bind_e(e[3]))
            if e[0] == "Unary":
                return ("Unary", e[1], bind_e(e[2]))
            if e[0] == "Ident":
                if e[1] not in par_binds:
                    self.err(f"Could not bind {e[1]}")
                return ("Real", par_binds[e[1]])
            return e

        def bind_a(a):
            if a[0] == "ArgBit":
                self.err("Attempted to bind an ArgBit")
            if a[1] not in arg_binds:
                self.err(f"Could not bind {a[1]}")
            return arg_binds[a[1]]

        bound = []
        for op in uops:
            if op[0] == "U":
                bound.append(("U", bind_e(op[1]), bind_e(op[2]), bind_e(op[3]), bind_a(op[4])))
            elif op[0] == "CX":
                bound.append(("CX", bind_a(op[1]), bind_a(op[2])))
            elif op[0] == "Barrier":
                bound.append(("Barrier", [bind_a(a) for a in op[1]]))
            elif op[0] == "Func":
                bound.append(("Func", op[1], [bind_e(e) for e in op[2]], [bind_a(a) for a in op[3]]))
            else:
                bound.append(op)
        for op in bound:
            self.run_stmt(("UOp", op))


class DenseBackend:
    """The reference's own algorithm (oracle.dense) behind the evaluator seam."""

    def __init__(self):
        self.n = {}

    def mk(self, n):
        return (n, _dense.mkStateVec(n))

    def tensor(self, a, b):
        return (a[0] + b[0], _dense.tensor(a[1], b[1]))

    def dimension(self, sv):
        return sv[0]

    def apply_1q(self, sv, q, m):
        return (sv[0], _dense.apply(_dense.onJust(sv[0], q, m), sv[1]))

    def apply_range(self, sv, f, l, m):
        return (sv[0], _dense.apply(_dense.onRange(sv[0], f, l, m), sv[1]))

    def apply_cnot(self, sv, c, t):
        return (sv[0], _dense.apply(_dense.cnot(sv[0], c, t), sv[1]))

    def measure_qubit(self, sv, q, r):
        bit, v, self.last_pone = _dense.measureQubit(sv[0], q, r, sv[1])
        return bit, (sv[0], v)

    def collapse(self, sv, q, b):
        return (sv[0], _dense.collapse(sv[0], q, b, sv[1]))


class StructuredBackend:
    """oracle.structured behind the same seam (for programs too wide for DenseBackend)."""

    def mk(self, n):
        from . import structured as S
        return (n, S.mk_state(n))

    def tensor(self, a, b):
        from . import structured as S
        return (a[0] + b[0], S.tensor(a[1], b[1]))

    def dimension(self, sv):
        return sv[0]

    def apply_1q(self, sv, q, m):
        from . import structured as S
        return (sv[0], S.apply_1q(sv[0], q, m, sv[1]))

    def apply_range(self, sv, f, l, m):
        from . import structured as S
        v = sv[1]
        for i in reversed(range(f, l + 1)):  # mconcat applies the last factor first
            v = S.apply_1q(sv[0], i, m, v)
        return (sv[0], v)

    def apply_cnot(self, sv, c, t):
        from . import structured as S
        return (sv[0], S.apply_cnot(sv[0], c, t, sv[1]))

    def measure_qubit(self, sv, q, r):
        from . import structured as S
        bit, v, self.last_pone = S.measure_qubit(sv[0], q, r, sv[1])
        return bit, (sv[0], v)

    def collapse(self, sv, q, b):
        from . import structured as S
        return (sv[0], S.collapse(sv[0], q, b, sv[1]))


def run_qasm(src: str, backend=None, draws=None, ref_faithful=True, trace=None):
    """Parse + evaluate.  ``draws``: iterable of uniform variates consumed one per
    measureQubit (use -1.0 to force One where possible, 2.0 to force Zero)."""
    backend = backend or DenseBackend()
    it = iter(draws if draws is not None else [])

    def draw():
        try:
            return next(it)
        except StopIteration:
            raise RuntimeError("ran out of measurement draws")

    ev = Evaluator(backend, draw, ref_faithful=ref_faithful, trace=trace)
    ev.run(parse(src))
    return ev.ps
