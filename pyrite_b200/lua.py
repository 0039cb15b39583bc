"""A small Lua interpreter (the subset of Lua 5.3/5.4 that project files use, and a good deal more).

pyrite evaluates `project.lua` with an embedded Lua (mlua 0.8.9; pyrite/src/project/mod.rs:29-93).
No Lua exists in this image, so the project surface is evaluated here: lexer -> recursive-descent
parser -> tree-walking evaluator.  Supported: all expression forms and operators (arithmetic with the
integer/float distinction, `//`, `%`, `^`, `..`, comparisons, `and`/`or`/`not`, `#`, bitwise ops),
table constructors, function definitions / closures / varargs / method calls / call sugar
(`f{...}`, `f"..."`), `local`, multiple assignment, `if`/`while`/`repeat`/numeric and generic `for`,
`break`, `goto`-less control flow, `return`, metatables (`__index`, `__newindex`, `__call`, arithmetic,
`__concat`, `__eq`, `__lt`, `__le`, `__len`, `__unm`, `__tostring`), and a standard-library subset
(`pairs`, `ipairs`, `next`, `type`, `tostring`, `tonumber`, `select`, `rawget`/`rawset`/`rawequal`/`rawlen`,
`setmetatable`/`getmetatable`, `assert`, `error`, `pcall`, `print`, `unpack`, `require`, `math.*`,
`string.*` basics, `table.*` basics).  Not supported: coroutines, `goto`, string patterns, integer
division corner cases of 64-bit wrap-around, the `io`/`os`/`debug` libraries.
"""
from __future__ import annotations

import math
import re
from pathlib import Path
from typing import Any, Callable, Dict, List, Optional


class LuaError(Exception):
    def __init__(self, message, value=None):
        super().__init__(message)
        self.value = message if value is None else value


class LuaTable:
    __slots__ = ("hash", "meta")

    def __init__(self):
        self.hash: Dict[Any, Any] = {}
        self.meta: Optional["LuaTable"] = None

    @staticmethod
    def _key(k):
        if isinstance(k, float) and k.is_integer():
            return int(k)
        return k

    def get(self, k):
        return self.hash.get(self._key(k))

    def set(self, k, v):
        k = self._key(k)
        if k is None:
            raise LuaError("table index is nil")
        if isinstance(k, float) and k != k:
            raise LuaError("table index is NaN")
        if v is None:
            self.hash.pop(k, None)
        else:
            self.hash[k] = v

    def length(self):
        n = 0
        h = self.hash
        while (n + 1) in h:
            n += 1
        return n

    def items(self):
        return list(self.hash.items())


class LuaFunction:
    __slots__ = ("params", "vararg", "body", "env", "name", "interp")

    def __init__(self, params, vararg, body, env, name, interp):
        self.params, self.vararg, self.body, self.env, self.name, self.interp = params, vararg, body, env, name, interp

    def __call__(self, *args):
        return self.interp.call_function(self, list(args))


# --------------------------------------------------------------------------------------- lexer
KEYWORDS = {"and", "break", "do", "else", "elseif", "end", "false", "for", "function", "goto", "if", "in", "local", "nil", "not", "or",
            "repeat", "return", "then", "true", "until", "while"}
_TOKEN = re.compile(r"""
    (?P<ws>[ \t\r]+) | (?P<nl>\n) |
    (?P<longcomment>--\[(?P<lc_eq>=*)\[) | (?P<comment>--[^\n]*) |
    (?P<longstring>\[(?P<ls_eq>=*)\[) |
    (?P<number>0[xX][0-9a-fA-F]+(?:\.[0-9a-fA-F]*)?(?:[pP][+-]?\d+)? | \d+\.?\d*(?:[eE][+-]?\d+)? | \.\d+(?:[eE][+-]?\d+)?) |
    (?P<name>[A-Za-z_][A-Za-z0-9_]*) |
    (?P<string>"(?:\\.|[^"\\\n])*" | '(?:\\.|[^'\\\n])*') |
    (?P<op>\.\.\.|\.\.|==|~=|<=|>=|<<|>>|//|::|[-+*/%^\#&~|<>=(){}\[\];:,.])
""", re.X)
_ESCAPES = {"n": "\n", "t": "\t", "r": "\r", "a": "\a", "b": "\b", "f": "\f", "v": "\v", "\\": "\\", '"': '"', "'": "'", "\n": "\n"}


def _unescape(s: str) -> str:
    out, i = [], 0
    while i < len(s):
        c = s[i]
        if c != "\\":
            out.append(c)
            i += 1
            continue
        i += 1
        c = s[i]
        if c in _ESCAPES:
            out.append(_ESCAPES[c]); i += 1
        elif c == "x":
            out.append(chr(int(s[i + 1:i + 3], 16))); i += 3
        elif c.isdigit():
            j = i
            while j < len(s) and j < i + 3 and s[j].isdigit():
                j += 1
            out.append(chr(int(s[i:j]))); i = j
        elif c == "z":
            i += 1
            while i < len(s) and s[i].isspace():
                i += 1
        elif c == "u":
            j = s.index("}", i)
            out.append(chr(int(s[i + 2:j], 16))); i = j + 1
        else:
            raise LuaError(f"invalid escape sequence '\\{c}'")
    return "".join(out)


def tokenize(src: str, chunk: str):
    toks, pos, line = [], 0, 1
    n = len(src)
    if src.startswith("#"):
        pos = src.find("\n") if "\n" in src else n
    while pos < n:
        m = _TOKEN.match(src, pos)
        if not m:
            raise LuaError(f"{chunk}:{line}: unexpected symbol near '{src[pos:pos + 10]}'")
        kind = m.lastgroup
        text = m.group(0)
        if m.group("longcomment") is not None or m.group("longstring") is not None:
            is_comment = m.group("longcomment") is not None
            eq = m.group("lc_eq") if is_comment else m.group("ls_eq")
            close = "]" + eq + "]"
            end = src.find(close, m.end())
            if end < 0:
                raise LuaError(f"{chunk}:{line}: unfinished long {'comment' if is_comment else 'string'}")
            body = src[m.end():end]
            if not is_comment:
                if body.startswith("\n"):
                    body = body[1:]
                toks.append(("string", body, line))
            line += src.count("\n", pos, end + len(close))
            pos = end + len(close)
            continue
        pos = m.end()
        if kind == "nl":
            line += 1
        elif kind in ("ws", "comment"):
            pass
        elif kind == "number":
            t = text.lower()
            if t.startswith("0x"):
                if "." in t or "p" in t:
                    value = float.fromhex(t if "p" in t else t + "p0")
                else:
                    value = int(t, 16)
            elif any(c in t for c in ".e"):
                value = float(t)
            else:
                value = int(t)
            toks.append(("number", value, line))
        elif kind == "name":
            toks.append(("keyword" if text in KEYWORDS else "name", text, line))
        elif kind == "string":
            toks.append(("string", _unescape(text[1:-1]), line))
            line += text.count("\n")
        else:
            toks.append(("op", text, line))
    toks.append(("eof", None, line))
    return toks


# --------------------------------------------------------------------------------------- parser
BINARY_PRIORITY = {"or": (1, 1), "and": (2, 2), "<": (3, 3), ">": (3, 3), "<=": (3, 3), ">=": (3, 3), "~=": (3, 3), "==": (3, 3),
                   "|": (4, 4), "~": (5, 5), "&": (6, 6), "<<": (7, 7), ">>": (7, 7), "..": (9, 8), "+": (10, 10), "-": (10, 10),
                   "*": (11, 11), "/": (11, 11), "//": (11, 11), "%": (11, 11), "^": (14, 13)}
UNARY_PRIORITY = 12


class Parser:
    def __init__(self, src: str, chunk: str):
        self.chunk = chunk
        self.toks = tokenize(src, chunk)
        self.i = 0

    # -- helpers
    @property
    def tok(self):
        return self.toks[self.i]

    def error(self, msg):
        raise LuaError(f"{self.chunk}:{self.tok[2]}: {msg} near '{self.tok[1] if self.tok[1] is not None else '<eof>'}'")

    def check(self, kind, value=None):
        t = self.tok
        return t[0] == kind and (value is None or t[1] == value)

    def accept(self, kind, value=None):
        if self.check(kind, value):
            self.i += 1
            return True
        return False

    def expect(self, kind, value=None):
        if not self.check(kind, value):
            self.error(f"'{value or kind}' expected")
        t = self.tok
        self.i += 1
        return t

    def block_follow(self):
        t = self.tok
        return t[0] == "eof" or (t[0] == "keyword" and t[1] in ("else", "elseif", "end", "until"))

    # -- statements
    def parse_chunk(self):
        body = self.block()
        if not self.check("eof"):
            self.error("'<eof>' expected")
        return body

    def block(self):
        stmts = []
        while not self.block_follow():
            if self.check("keyword", "return"):
                line = self.tok[2]
                self.i += 1
                exprs = []
                if not self.block_follow() and not self.check("op", ";"):
                    exprs = self.exprlist()
                self.accept("op", ";")
                stmts.append(("return", exprs, line))
                break
            s = self.statement()
            if s is not None:
                stmts.append(s)
        return stmts

    def statement(self):
        t = self.tok
        line = t[2]
        if t[0] == "op" and t[1] == ";":
            self.i += 1
            return None
        if t[0] == "op" and t[1] == "::":
            self.error("labels / goto are not supported")
        if t[0] == "keyword":
            k = t[1]
            if k == "if":
                self.i += 1
                clauses = []
                cond = self.expr()
                self.expect("keyword", "then")
                clauses.append((cond, self.block()))
                orelse = None
                while True:
                    if self.accept("keyword", "elseif"):
                        cond = self.expr()
                        self.expect("keyword", "then")
                        clauses.append((cond, self.block()))
                    elif self.accept("keyword", "else"):
                        orelse = self.block()
                        self.expect("keyword", "end")
                        break
                    else:
                        self.expect("keyword", "end")
                        break
                return ("if", clauses, orelse, line)
            if k == "while":
                self.i += 1
                cond = self.expr()
                self.expect("keyword", "do")
                body = self.block()
                self.expect("keyword", "end")
                return ("while", cond, body, line)
            if k == "do":
                self.i += 1
                body = self.block()
                self.expect("keyword", "end")
                return ("do", body, line)
            if k == "for":
                self.i += 1
                n1 = self.expect("name")[1]
                if self.accept("op", "="):
                    start = self.expr()
                    self.expect("op", ",")
                    stop = self.expr()
                    step = self.expr() if self.accept("op", ",") else None
                    self.expect("keyword", "do")
                    body = self.block()
                    self.expect("keyword", "end")
                    return ("fornum", n1, start, stop, step, body, line)
                names = [n1]
                while self.accept("op", ","):
                    names.append(self.expect("name")[1])
                self.expect("keyword", "in")
                exprs = self.exprlist()
                self.expect("keyword", "do")
                body = self.block()
                self.expect("keyword", "end")
                return ("forin", names, exprs, body, line)
            if k == "repeat":
                self.i += 1
                body = self.block()
                self.expect("keyword", "until")
                return ("repeat", body, self.expr(), line)
            if k == "function":
                self.i += 1
                target = ("name", self.expect("name")[1], line)
                is_method = False
                full = target[1]
                while self.check("op", ".") or self.check("op", ":"):
                    sep = self.tok[1]
                    self.i += 1
                    key = self.expect("name")[1]
                    target = ("index", target, ("const", key), line)
                    full += sep + key
                    if sep == ":":
                        is_method = True
                        break
                func = self.funcbody(full, is_method, line)
                return ("assign", [target], [func], line)
            if k == "local":
                self.i += 1
                if self.accept("keyword", "function"):
                    name = self.expect("name")[1]
                    return ("localfunc", name, self.funcbody(name, False, line), line)
                names = [self.expect("name")[1]]
                self.attrib()
                while self.accept("op", ","):
                    names.append(self.expect("name")[1])
                    self.attrib()
                exprs = self.exprlist() if self.accept("op", "=") else []
                return ("local", names, exprs, line)
            if k == "break":
                self.i += 1
                return ("break", line)
            if k == "goto":
                self.error("goto is not supported")
        # expression statement: call or assignment
        e = self.suffixedexp()
        if self.check("op", "=") or self.check("op", ","):
            targets = [e]
            while self.accept("op", ","):
                targets.append(self.suffixedexp())
            self.expect("op", "=")
            exprs = self.exprlist()
            for tgt in targets:
                if tgt[0] not in ("name", "index"):
                    self.error("syntax error")
            return ("assign", targets, exprs, line)
        if e[0] not in ("call", "method"):
            self.error("syntax error")
        return ("exprstat", e, line)

    def attrib(self):
        if self.accept("op", "<"):
            self.expect("name")
            self.expect("op", ">")

    def funcbody(self, name, is_method, line):
        self.expect("op", "(")
        params, vararg = (["self"] if is_method else []), False
        if not self.check("op", ")"):
            while True:
                if self.accept("op", "..."):
                    vararg = True
                    break
                params.append(self.expect("name")[1])
                if not self.accept("op", ","):
                    break
        self.expect("op", ")")
        body = self.block()
        self.expect("keyword", "end")
        return ("function", params, vararg, body, name, line)

    # -- expressions
    def exprlist(self):
        exprs = [self.expr()]
        while self.accept("op", ","):
            exprs.append(self.expr())
        return exprs

    def primaryexp(self):
        t = self.tok
        if t[0] == "name":
            self.i += 1
            return ("name", t[1], t[2])
        if t[0] == "op" and t[1] == "(":
            self.i += 1
            e = self.expr()
            self.expect("op", ")")
            return ("paren", e)
        self.error("unexpected symbol")

    def suffixedexp(self):
        e = self.primaryexp()
        while True:
            t = self.tok
            if t[0] == "op" and t[1] == ".":
                self.i += 1
                e = ("index", e, ("const", self.expect("name")[1]), t[2])
            elif t[0] == "op" and t[1] == "[":
                self.i += 1
                k = self.expr()
                self.expect("op", "]")
                e = ("index", e, k, t[2])
            elif t[0] == "op" and t[1] == ":":
                self.i += 1
                name = self.expect("name")[1]
                e = ("method", e, name, self.callargs(), t[2])
            elif (t[0] == "op" and t[1] in ("(", "{")) or t[0] == "string":
                e = ("call", e, self.callargs(), t[2])
            else:
                return e

    def callargs(self):
        t = self.tok
        if t[0] == "string":
            self.i += 1
            return [("const", t[1])]
        if t[0] == "op" and t[1] == "{":
            return [self.tablecons()]
        self.expect("op", "(")
        args = []
        if not self.check("op", ")"):
            args = self.exprlist()
        self.expect("op", ")")
        return args

    def tablecons(self):
        line = self.expect("op", "{")[2]
        array, fields = [], []  # entries in order: ("pos", expr) | ("key", kexpr, vexpr)
        entries = []
        while not self.check("op", "}"):
            if self.check("op", "["):
                self.i += 1
                k = self.expr()
                self.expect("op", "]")
                self.expect("op", "=")
                entries.append(("key", k, self.expr()))
            elif self.check("name") and self.toks[self.i + 1][0] == "op" and self.toks[self.i + 1][1] == "=":
                k = self.tok[1]
                self.i += 2
                entries.append(("key", ("const", k), self.expr()))
            else:
                entries.append(("pos", self.expr()))
            if not (self.accept("op", ",") or self.accept("op", ";")):
                break
        self.expect("op", "}")
        del array, fields
        return ("table", entries, line)

    def simpleexp(self):
        t = self.tok
        if t[0] == "number" or t[0] == "string":
            self.i += 1
            return ("const", t[1])
        if t[0] == "keyword":
            if t[1] == "nil":
                self.i += 1
                return ("const", None)
            if t[1] == "true":
                self.i += 1
                return ("const", True)
            if t[1] == "false":
                self.i += 1
                return ("const", False)
            if t[1] == "function":
                self.i += 1
                return self.funcbody("anonymous", False, t[2])
        if t[0] == "op" and t[1] == "...":
            self.i += 1
            return ("vararg", t[2])
        if t[0] == "op" and t[1] == "{":
            return self.tablecons()
        return self.suffixedexp()

    def expr(self, limit=0):
        t = self.tok
        if (t[0] == "keyword" and t[1] == "not") or (t[0] == "op" and t[1] in ("-", "#", "~")):
            self.i += 1
            operand = self.expr(UNARY_PRIORITY)
            left = ("unop", t[1], operand, t[2])
        else:
            left = self.simpleexp()
        while True:
            t = self.tok
            op = t[1] if t[0] in ("op", "keyword") else None
            if op not in BINARY_PRIORITY:
                return left
            lp, rp = BINARY_PRIORITY[op]
            if lp <= limit:
                return left
            self.i += 1
            right = self.expr(rp)
            if op == "and":
                left = ("and", left, right)
            elif op == "or":
                left = ("or", left, right)
            else:
                left = ("binop", op, left, right, t[2])


# --------------------------------------------------------------------------------------- evaluator
class _Break(Exception):
    pass


class _Return(Exception):
    def __init__(self, values):
        self.values = values


class Scope:
    __slots__ = ("vars", "parent")

    def __init__(self, parent=None):
        self.vars: Dict[str, Any] = {}
        self.parent = parent

    def lookup(self, name):
        s = self
        while s is not None:
            if name in s.vars:
                return s
            s = s.parent
        return None


def lua_type(v) -> str:
    if v is None:
        return "nil"
    if isinstance(v, bool):
        return "boolean"
    if isinstance(v, (int, float)):
        return "number"
    if isinstance(v, str):
        return "string"
    if isinstance(v, LuaTable):
        return "table"
    if callable(v):
        return "function"
    return "userdata"


def truthy(v) -> bool:
    return v is not None and v is not False


def tostring(v) -> str:
    if v is None:
        return "nil"
    if v is True:
        return "true"
    if v is False:
        return "false"
    if isinstance(v, int):
        return str(v)
    if isinstance(v, float):
        if v != v:
            return "nan" if math.copysign(1, v) > 0 else "-nan"
        if v in (math.inf, -math.inf):
            return "inf" if v > 0 else "-inf"
        if v.is_integer() and abs(v) < 1e16:
            return f"{v:.1f}"
        return f"{v:.14g}"
    if isinstance(v, str):
        return v
    return f"{lua_type(v)}: 0x{id(v):08x}"


def tonumber(v, base=None):
    if base is not None:
        try:
            return int(str(v).strip(), int(base))
        except ValueError:
            return None
    if isinstance(v, (int, float)) and not isinstance(v, bool):
        return v
    if isinstance(v, str):
        s = v.strip().lower()
        try:
            if s.startswith(("0x", "-0x")):
                return int(s, 16)
            return int(s)
        except ValueError:
            try:
                return float(s)
            except ValueError:
                return None
    return None


class Interpreter:
    def __init__(self, search_dirs: Optional[List[Path]] = None, output: Callable[[str], None] = print):
        self.G = LuaTable()
        self.search_dirs = [Path(d) for d in (search_dirs or [])]
        self.loaded: Dict[str, Any] = {}
        self.output = output
        self.string_meta = LuaTable()
        self.depth = 0
        import sys

        if sys.getrecursionlimit() < 8000:  # a Lua call is about a dozen Python frames; the Lua depth limit is 190 calls
            sys.setrecursionlimit(8000)
        self._install_stdlib()

    # -- running code
    def run(self, src: str, chunk: str = "=chunk", args: Optional[list] = None) -> list:
        body = Parser(src, chunk).parse_chunk()
        fn = LuaFunction([], True, body, Scope(), chunk, self)
        saved = self.current_chunk
        self.current_chunk = chunk.lstrip("=@")
        try:
            return self.call_function(fn, list(args or []))
        finally:
            self.current_chunk = saved

    def run_file(self, path) -> list:
        path = Path(path)
        return self.run(path.read_text(), path.name)

    # -- calls
    def call(self, f, args: list) -> list:
        if isinstance(f, LuaFunction):
            return self.call_function(f, args)
        if callable(f):
            r = f(*args)
            if r is None:
                return []
            if isinstance(r, tuple):
                return list(r)
            if isinstance(r, list):
                return r
            return [r]
        h = self.metamethod(f, "__call")
        if h is not None:
            return self.call(h, [f] + args)
        raise LuaError(f"attempt to call a {lua_type(f)} value")

    def call_function(self, fn: LuaFunction, args: list) -> list:
        scope = Scope(fn.env)
        np_ = len(fn.params)
        for i, p in enumerate(fn.params):
            scope.vars[p] = args[i] if i < len(args) else None
        varargs = args[np_:] if fn.vararg else []
        self.depth += 1
        if self.depth > 190:
            self.depth -= 1
            raise LuaError("stack overflow")
        try:
            self.exec_block(fn.body, scope, varargs)
        except _Return as r:
            return r.values
        finally:
            self.depth -= 1
        return []

    # -- metatables
    def getmetatable(self, v):
        if isinstance(v, LuaTable):
            return v.meta
        if isinstance(v, str):
            return self.string_meta
        return None

    def metamethod(self, v, event):
        mt = self.getmetatable(v)
        return mt.hash.get(event) if mt is not None else None

    def index(self, obj, key, line=None):
        for _ in range(100):
            if isinstance(obj, LuaTable):
                v = obj.get(key)
                if v is not None:
                    return v
                h = obj.meta.hash.get("__index") if obj.meta is not None else None
                if h is None:
                    return None
            else:
                h = self.metamethod(obj, "__index")
                if h is None:
                    raise LuaError(f"attempt to index a {lua_type(obj)} value" + (f" (field '{key}')" if isinstance(key, str) else ""))
            if isinstance(h, LuaTable):
                obj = h
                continue
            r = self.call(h, [obj, key])
            return r[0] if r else None
        raise LuaError("'__index' chain too long; possible loop")

    def setindex(self, obj, key, value):
        for _ in range(100):
            if isinstance(obj, LuaTable):
                h = obj.meta.hash.get("__newindex") if obj.meta is not None else None
                if h is None or obj.get(key) is not None:
                    obj.set(key, value)
                    return
            else:
                h = self.metamethod(obj, "__newindex")
                if h is None:
                    raise LuaError(f"attempt to index a {lua_type(obj)} value")
            if isinstance(h, LuaTable):
                obj = h
                continue
            self.call(h, [obj, key, value])
            return
        raise LuaError("'__newindex' chain too long; possible loop")

    # -- operators
    ARITH_EVENTS = {"+": "__add", "-": "__sub", "*": "__mul", "/": "__div", "%": "__mod", "^": "__pow", "//": "__idiv", "&": "__band",
                    "|": "__bor", "~": "__bxor", "<<": "__shl", ">>": "__shr", "..": "__concat"}

    def arith(self, op, a, b):
        na = a if isinstance(a, (int, float)) and not isinstance(a, bool) else (tonumber(a) if isinstance(a, str) and op != ".." else None)
        nb = b if isinstance(b, (int, float)) and not isinstance(b, bool) else (tonumber(b) if isinstance(b, str) and op != ".." else None)
        if op == "..":
            if isinstance(a, (str, int, float)) and isinstance(b, (str, int, float)) and not isinstance(a, bool) and not isinstance(b, bool):
                return tostring(a) + tostring(b)
        elif na is not None and nb is not None:
            try:
                if op == "+":
                    return na + nb
                if op == "-":
                    return na - nb
                if op == "*":
                    return na * nb
                if op == "/":
                    return float(na) / float(nb) if nb != 0 else (math.nan if na == 0 or na != na else math.copysign(math.inf, float(na)) * (math.copysign(1.0, float(nb))))
                if op == "^":
                    return float(na) ** float(nb)
                if op == "//":
                    if isinstance(na, int) and isinstance(nb, int):
                        if nb == 0:
                            raise LuaError("attempt to perform 'n//0'")
                        return na // nb
                    return math.floor(float(na) / float(nb)) * 1.0 if nb != 0 else math.copysign(math.inf, float(na))
                if op == "%":
                    if isinstance(na, int) and isinstance(nb, int):
                        if nb == 0:
                            raise LuaError("attempt to perform 'n%%0'")
                        return na % nb
                    if nb == 0:
                        return math.nan
                    r = math.fmod(float(na), float(nb))
                    return r + nb if r != 0 and (r < 0) != (nb < 0) else r
                ia, ib = self.toint(na), self.toint(nb)
                if op == "&":
                    return ia & ib
                if op == "|":
                    return ia | ib
                if op == "~":
                    return ia ^ ib
                if op == "<<":
                    return (ia << ib) & 0xFFFFFFFFFFFFFFFF if ib >= 0 else ia >> -ib
                if op == ">>":
                    return (ia & 0xFFFFFFFFFFFFFFFF) >> ib if ib >= 0 else ia << -ib
            except OverflowError:
                return math.inf
        event = self.ARITH_EVENTS[op]
        h = self.metamethod(a, event) or self.metamethod(b, event)
        if h is not None:
            r = self.call(h, [a, b])
            return r[0] if r else None
        bad = b if (na is not None or (op == ".." and isinstance(a, (str, int, float)))) else a
        what = "concatenate" if op == ".." else "perform arithmetic on"
        raise LuaError(f"attempt to {what} a {lua_type(bad)} value")

    @staticmethod
    def toint(v):
        if isinstance(v, int):
            return v
        if isinstance(v, float) and v.is_integer():
            return int(v)
        raise LuaError("number has no integer representation")

    def equals(self, a, b):
        if a is b:
            return True
        if isinstance(a, bool) or isinstance(b, bool):
            return a is b
        if isinstance(a, (int, float)) and isinstance(b, (int, float)):
            return a == b
        if isinstance(a, str) and isinstance(b, str):
            return a == b
        if isinstance(a, LuaTable) and isinstance(b, LuaTable):
            h = self.metamethod(a, "__eq") or self.metamethod(b, "__eq")
            if h is not None:
                r = self.call(h, [a, b])
                return truthy(r[0] if r else None)
        return False

    def less(self, a, b, event="__lt"):
        num = lambda v: isinstance(v, (int, float)) and not isinstance(v, bool)  # noqa: E731
        if (num(a) and num(b)) or (isinstance(a, str) and isinstance(b, str)):
            return a < b if event == "__lt" else a <= b
        h = self.metamethod(a, event) or self.metamethod(b, event)
        if h is not None:
            r = self.call(h, [a, b])
            return truthy(r[0] if r else None)
        raise LuaError(f"attempt to compare {lua_type(a)} with {lua_type(b)}")

    def length(self, v):
        if isinstance(v, str):
            return len(v.encode("utf-8"))
        h = self.metamethod(v, "__len")
        if h is not None:
            r = self.call(h, [v])
            return r[0] if r else None
        if isinstance(v, LuaTable):
            return v.length()
        raise LuaError(f"attempt to get length of a {lua_type(v)} value")

    # -- evaluation
    def eval_multi(self, exprs, scope, varargs) -> list:
        out = []
        for i, e in enumerate(exprs):
            if i == len(exprs) - 1 and e[0] in ("call", "method", "vararg"):
                out.extend(self.eval_many(e, scope, varargs))
            else:
                out.append(self.eval(e, scope, varargs))
        return out

    def eval_many(self, e, scope, varargs) -> list:
        k = e[0]
        if k == "call":
            try:
                f = self.eval(e[1], scope, varargs)
                return self.call(f, self.eval_multi(e[2], scope, varargs))
            except LuaError as err:
                raise self._located(err, e[3]) from None
        if k == "method":
            try:
                obj = self.eval(e[1], scope, varargs)
                f = self.index(obj, e[2])
                if f is None:
                    raise LuaError(f"attempt to call a nil value (method '{e[2]}')")
                return self.call(f, [obj] + self.eval_multi(e[3], scope, varargs))
            except LuaError as err:
                raise self._located(err, e[4]) from None
        if k == "vararg":
            return list(varargs)
        return [self.eval(e, scope, varargs)]

    def _located(self, err: LuaError, line) -> LuaError:
        if getattr(err, "located", False) or not isinstance(err.value, str):
            return err
        new = LuaError(f"{self.current_chunk}:{line}: {err.value}")
        new.located = True
        return new

    current_chunk = "?"

    def eval(self, e, scope, varargs):
        k = e[0]
        if k == "const":
            return e[1]
        if k == "name":
            s = scope.lookup(e[1])
            return s.vars[e[1]] if s is not None else self.G.get(e[1])
        if k == "index":
            try:
                return self.index(self.eval(e[1], scope, varargs), self.eval(e[2], scope, varargs))
            except LuaError as err:
                raise self._located(err, e[3]) from None
        if k in ("call", "method"):
            r = self.eval_many(e, scope, varargs)
            return r[0] if r else None
        if k == "table":
            t = LuaTable()
            n = 1
            entries = e[1]
            for i, ent in enumerate(entries):
                if ent[0] == "key":
                    key = self.eval(ent[1], scope, varargs)
                    if key is None:
                        raise LuaError("table index is nil")
                    t.set(key, self.eval(ent[2], scope, varargs))
                elif i == len(entries) - 1 and ent[1][0] in ("call", "method", "vararg"):
                    for v in self.eval_many(ent[1], scope, varargs):
                        t.set(n, v)
                        n += 1
                else:
                    t.set(n, self.eval(ent[1], scope, varargs))
                    n += 1
            return t
        if k == "function":
            return LuaFunction(e[1], e[2], e[3], scope, e[4], self)
        if k == "binop":
            op = e[1]
            a = self.eval(e[2], scope, varargs)
            b = self.eval(e[3], scope, varargs)
            try:
                if op == "==":
                    return self.equals(a, b)
                if op == "~=":
                    return not self.equals(a, b)
                if op == "<":
                    return self.less(a, b)
                if op == "<=":
                    return self.less(a, b, "__le")
                if op == ">":
                    return self.less(b, a)
                if op == ">=":
                    return self.less(b, a, "__le")
                return self.arith(op, a, b)
            except LuaError as err:
                raise self._located(err, e[4]) from None
        if k == "and":
            a = self.eval(e[1], scope, varargs)
            return self.eval(e[2], scope, varargs) if truthy(a) else a
        if k == "or":
            a = self.eval(e[1], scope, varargs)
            return a if truthy(a) else self.eval(e[2], scope, varargs)
        if k == "unop":
            v = self.eval(e[2], scope, varargs)
            op = e[1]
            try:
                if op == "not":
                    return not truthy(v)
                if op == "#":
                    return self.length(v)
                if op == "-":
                    n = v if isinstance(v, (int, float)) and not isinstance(v, bool) else (tonumber(v) if isinstance(v, str) else None)
                    if n is not None:
                        return -n
                    h = self.metamethod(v, "__unm")
                    if h is None:
                        raise LuaError(f"attempt to perform arithmetic on a {lua_type(v)} value")
                    r = self.call(h, [v, v])
                    return r[0] if r else None
                if op == "~":
                    return ~self.toint(v)
            except LuaError as err:
                raise self._located(err, e[3]) from None
        if k == "paren":
            return self.eval(e[1], scope, varargs)
        if k == "vararg":
            return varargs[0] if varargs else None
        raise LuaError(f"cannot evaluate node {k}")

    def assign(self, target, value, scope, varargs):
        if target[0] == "name":
            s = scope.lookup(target[1])
            if s is not None:
                s.vars[target[1]] = value
            else:
                self.G.set(target[1], value)
        else:
            self.setindex(self.eval(target[1], scope, varargs), self.eval(target[2], scope, varargs), value)

    def exec_block(self, stmts, scope, varargs):
        for s in stmts:
            k = s[0]
            if k == "local":
                values = self.eval_multi(s[2], scope, varargs)
                # a fresh scope per `local` so that closures capture the right binding
                for i, name in enumerate(s[1]):
                    scope.vars[name] = values[i] if i < len(values) else None
            elif k == "assign":
                values = self.eval_multi(s[2], scope, varargs)
                try:
                    for i, tgt in enumerate(s[1]):
                        self.assign(tgt, values[i] if i < len(values) else None, scope, varargs)
                except LuaError as err:
                    raise self._located(err, s[3]) from None
            elif k == "exprstat":
                self.eval_many(s[1], scope, varargs)
            elif k == "return":
                if len(s[1]) == 1 and s[1][0][0] in ("call", "method"):
                    raise _Return(self.eval_many(s[1][0], scope, varargs))
                raise _Return(self.eval_multi(s[1], scope, varargs))
            elif k == "if":
                done = False
                for cond, body in s[1]:
                    if truthy(self.eval(cond, scope, varargs)):
                        self.exec_block(body, Scope(scope), varargs)
                        done = True
                        break
                if not done and s[2] is not None:
                    self.exec_block(s[2], Scope(scope), varargs)
            elif k == "while":
                try:
                    while truthy(self.eval(s[1], scope, varargs)):
                        self.exec_block(s[2], Scope(scope), varargs)
                except _Break:
                    pass
            elif k == "repeat":
                try:
                    while True:
                        inner = Scope(scope)
                        self.exec_block(s[1], inner, varargs)
                        if truthy(self.eval(s[2], inner, varargs)):
                            break
                except _Break:
                    pass
            elif k == "do":
                self.exec_block(s[1], Scope(scope), varargs)
            elif k == "fornum":
                start, stop = self.eval(s[2], scope, varargs), self.eval(s[3], scope, varargs)
                step = self.eval(s[4], scope, varargs) if s[4] is not None else 1
                start, stop, step = (tonumber(start) if isinstance(start, str) else start, tonumber(stop) if isinstance(stop, str) else stop,
                                     tonumber(step) if isinstance(step, str) else step)
                if not all(isinstance(v, (int, float)) and not isinstance(v, bool) for v in (start, stop, step)):
                    raise LuaError(f"{self.current_chunk}:{s[6]}: 'for' initial value must be a number")
                if step == 0:
                    raise LuaError(f"{self.current_chunk}:{s[6]}: 'for' step is zero")
                if not all(isinstance(v, int) for v in (start, stop, step)):
                    start, stop, step = float(start), float(stop), float(step)
                i = start
                try:
                    while (i <= stop) if step > 0 else (i >= stop):
                        inner = Scope(scope)
                        inner.vars[s[1]] = i
                        self.exec_block(s[5], inner, varargs)
                        i += step
                except _Break:
                    pass
            elif k == "forin":
                vals = self.eval_multi(s[2], scope, varargs)
                f, state, ctl = (vals + [None, None, None])[:3]
                try:
                    while True:
                        r = self.call(f, [state, ctl])
                        if not r or r[0] is None:
                            break
                        ctl = r[0]
                        inner = Scope(scope)
                        for i, name in enumerate(s[1]):
                            inner.vars[name] = r[i] if i < len(r) else None
                        self.exec_block(s[3], inner, varargs)
                except _Break:
                    pass
            elif k == "localfunc":
                scope.vars[s[1]] = None
                scope.vars[s[1]] = LuaFunction(s[2][1], s[2][2], s[2][3], scope, s[2][4], self)
            elif k == "break":
                raise _Break()
            else:
                raise LuaError(f"cannot execute node {k}")

    # -- standard library
    def _install_stdlib(self):
        G = self.G
        interp = self

        def lua_next(t, k=None):
            keys = list(t.hash.keys())
            if k is None:
                i = 0
            else:
                k = LuaTable._key(k)
                try:
                    i = keys.index(k) + 1
                except ValueError:
                    raise LuaError("invalid key to 'next'")
            if i >= len(keys):
                return [None]
            return [keys[i], t.hash[keys[i]]]

        def lua_pairs(t):
            h = interp.metamethod(t, "__pairs")
            if h is not None:
                return (interp.call(h, [t]) + [None, None, None])[:3]
            if not isinstance(t, LuaTable):
                raise LuaError(f"bad argument #1 to 'pairs' (table expected, got {lua_type(t)})")
            snapshot = t.items()
            state = {"i": 0}

            def it(_s=None, _c=None):
                while state["i"] < len(snapshot):
                    k, _ = snapshot[state["i"]]
                    state["i"] += 1
                    v = t.hash.get(k)
                    if v is not None:
                        return [k, v]
                return [None]

            return [it, t, None]

        def lua_ipairs(t):
            def it(tt, i):
                i = int(i) + 1
                v = interp.index(tt, i)
                return [None] if v is None else [i, v]

            return [it, t, 0]

        def lua_select(n, *args):
            if n == "#":
                return [len(args)]
            n = int(n)
            if n < 0:
                n = len(args) + n + 1
            return list(args[n - 1:])

        def lua_setmetatable(t, mt=None):
            if not isinstance(t, LuaTable):
                raise LuaError(f"bad argument #1 to 'setmetatable' (table expected, got {lua_type(t)})")
            if mt is not None and not isinstance(mt, LuaTable):
                raise LuaError("bad argument #2 to 'setmetatable' (nil or table expected)")
            t.meta = mt
            return [t]

        def lua_getmetatable(v=None):
            mt = interp.getmetatable(v)
            if mt is not None and mt.hash.get("__metatable") is not None:
                return [mt.hash["__metatable"]]
            return [mt]

        def lua_error(msg=None, level=1):
            raise LuaError(tostring(msg) if isinstance(msg, (str, int, float)) else "error object", msg)

        def lua_assert(*args):
            if not args or not truthy(args[0]):
                raise LuaError(tostring(args[1]) if len(args) > 1 else "assertion failed!", args[1] if len(args) > 1 else None)
            return list(args)

        def lua_pcall(f=None, *args):
            try:
                return [True] + interp.call(f, list(args))
            except LuaError as e:
                return [False, e.value]
            except RecursionError:
                return [False, "stack overflow"]

        def lua_tostring(v=None):
            h = interp.metamethod(v, "__tostring")
            if h is not None:
                r = interp.call(h, [v])
                return [r[0] if r else None]
            return [tostring(v)]

        def lua_print(*args):
            interp.output("\t".join(lua_tostring(a)[0] for a in args))

        def lua_unpack(t, i=1, j=None):
            j = t.length() if j is None else int(j)
            return [t.get(k) for k in range(int(i), j + 1)]

        def lua_require(name):
            if name in interp.loaded:
                return [interp.loaded[name]]
            rel = Path(*str(name).split("."))
            for d in interp.search_dirs:
                for cand in (d / (str(rel) + ".lua"), d / rel / "init.lua"):
                    if cand.exists():
                        saved = interp.current_chunk
                        interp.current_chunk = cand.name
                        try:
                            r = interp.run(cand.read_text(), cand.name, [name, str(cand)])
                        finally:
                            interp.current_chunk = saved
                        v = r[0] if r and r[0] is not None else True
                        interp.loaded[name] = v
                        return [v]
            raise LuaError(f"module '{name}' not found:" + "".join(f"\n\tno file '{d / (str(rel) + '.lua')}'" for d in interp.search_dirs))

        def reg(name, f):
            G.set(name, f)

        reg("next", lua_next); reg("pairs", lua_pairs); reg("ipairs", lua_ipairs); reg("select", lua_select)
        reg("type", lambda v=None: [lua_type(v)]); reg("tostring", lua_tostring); reg("tonumber", lambda v=None, b=None: [tonumber(v, b)])
        reg("rawget", lambda t, k: [t.get(k)]); reg("rawset", lambda t, k, v=None: [t.set(k, v), t][1:]); reg("rawequal", lambda a, b: [a is b or (a == b and type(a) is type(b))])
        reg("rawlen", lambda v: [v.length() if isinstance(v, LuaTable) else len(v)])
        reg("setmetatable", lua_setmetatable); reg("getmetatable", lua_getmetatable); reg("error", lua_error); reg("assert", lua_assert)
        reg("pcall", lua_pcall); reg("print", lua_print); reg("unpack", lua_unpack); reg("require", lua_require)
        G.set("_G", G)
        G.set("_VERSION", "Lua 5.4 (pyrite_b200 subset)")

        m = LuaTable()
        for name in ("sin", "cos", "tan", "asin", "acos", "exp", "sqrt"):
            m.set(name, (lambda fn: lambda x: [fn(float(x))])(getattr(math, name)))
        m.set("atan", lambda y, x=1.0: [math.atan2(float(y), float(x))])
        m.set("log", lambda x, b=None: [math.log(float(x)) if b is None else math.log(float(x), float(b))])
        m.set("floor", lambda x: [x if isinstance(x, int) else int(math.floor(x))])
        m.set("ceil", lambda x: [x if isinstance(x, int) else int(math.ceil(x))])
        m.set("abs", lambda x: [abs(x)])
        m.set("max", lambda *a: [max(a)]); m.set("min", lambda *a: [min(a)])
        m.set("fmod", lambda a, b: [math.fmod(a, b)]); m.set("pow", lambda a, b: [float(a) ** float(b)])
        m.set("tointeger", lambda x: [int(x) if isinstance(x, int) or (isinstance(x, float) and x.is_integer()) else None])
        m.set("type", lambda x=None: ["integer" if isinstance(x, int) and not isinstance(x, bool) else ("float" if isinstance(x, float) else None)])
        m.set("pi", math.pi); m.set("huge", math.inf); m.set("maxinteger", 2 ** 63 - 1); m.set("mininteger", -2 ** 63)
        m.set("rad", lambda x: [math.radians(x)]); m.set("deg", lambda x: [math.degrees(x)])
        G.set("math", m)

        s = LuaTable()
        s.set("len", lambda v: [len(v.encode("utf-8"))]); s.set("upper", lambda v: [v.upper()]); s.set("lower", lambda v: [v.lower()])
        s.set("rep", lambda v, n, sep="": [sep.join([v] * max(int(n), 0))]); s.set("reverse", lambda v: [v[::-1]])

        def lua_sub(v, i=1, j=-1):
            n = len(v)
            i, j = int(i), int(j)
            if i < 0:
                i = max(n + i + 1, 1)
            elif i == 0:
                i = 1
            if j < 0:
                j = n + j + 1
            elif j > n:
                j = n
            return [v[i - 1:j] if i <= j else ""]

        def lua_format(fmt, *args):
            args = list(args)
            out, i = [], 0
            spec = re.compile(r"%([-+ #0]*)(\d+)?(?:\.(\d+))?([cdiouxXeEfgGqsaA%])")
            pos = 0
            for mm in spec.finditer(fmt):
                out.append(fmt[pos:mm.start()])
                pos = mm.end()
                conv = mm.group(4)
                if conv == "%":
                    out.append("%")
                    continue
                if i >= len(args):
                    raise LuaError(f"bad argument #{i + 2} to 'format' (no value)")
                a = args[i]
                i += 1
                pyfmt = "%" + mm.group(1) + (mm.group(2) or "") + ("." + mm.group(3) if mm.group(3) is not None else "")
                if conv in "di":
                    out.append((pyfmt + "d") % interp.toint(a))
                elif conv in "ouxX":
                    out.append((pyfmt + conv) % interp.toint(a))
                elif conv in "eEfgG":
                    out.append((pyfmt + conv) % float(a))
                elif conv == "c":
                    out.append(chr(int(a)))
                elif conv == "q":
                    out.append('"' + str(a).replace("\\", "\\\\").replace('"', '\\"').replace("\n", "\\n") + '"')
                elif conv in "aA":
                    out.append(float(a).hex())
                else:
                    out.append((pyfmt + "s") % lua_tostring(a)[0])
            out.append(fmt[pos:])
            return ["".join(out)]

        s.set("sub", lua_sub); s.set("format", lua_format)
        s.set("byte", lambda v, i=1, j=None: [ord(c) for c in lua_sub(v, i, i if j is None else j)[0]])
        s.set("char", lambda *a: ["".join(chr(int(c)) for c in a)])

        def lua_find(v, pat, init=1, plain=None):
            idx = v.find(pat, max(int(init) - 1, 0))  # plain find only (no Lua patterns)
            return [None] if idx < 0 else [idx + 1, idx + len(pat)]

        s.set("find", lua_find)
        G.set("string", s)
        self.string_meta.set("__index", s)

        t = LuaTable()

        def t_insert(tbl, *a):
            n = tbl.length()
            if len(a) == 1:
                tbl.set(n + 1, a[0])
            else:
                pos = int(a[0])
                for k in range(n, pos - 1, -1):
                    tbl.set(k + 1, tbl.get(k))
                tbl.set(pos, a[1])

        def t_remove(tbl, pos=None):
            n = tbl.length()
            if n == 0 and pos is None:
                return [None]
            pos = n if pos is None else int(pos)
            v = tbl.get(pos)
            for k in range(pos, n):
                tbl.set(k, tbl.get(k + 1))
            tbl.set(n, None)
            return [v]

        def t_concat(tbl, sep="", i=1, j=None):
            j = tbl.length() if j is None else int(j)
            return [sep.join(tostring(tbl.get(k)) for k in range(int(i), j + 1))]

        def t_sort(tbl, comp=None):
            import functools

            vals = [tbl.get(k) for k in range(1, tbl.length() + 1)]
            lt = (lambda a, b: truthy((interp.call(comp, [a, b]) or [None])[0])) if comp is not None else (lambda a, b: interp.less(a, b))
            vals.sort(key=functools.cmp_to_key(lambda a, b: -1 if lt(a, b) else (1 if lt(b, a) else 0)))
            for k, v in enumerate(vals, 1):
                tbl.set(k, v)

        t.set("insert", t_insert); t.set("remove", t_remove); t.set("concat", t_concat); t.set("sort", t_sort); t.set("unpack", lua_unpack)
        t.set("pack", lambda *a: [_packed(a)])
        G.set("table", t)

        pkg = LuaTable()
        pkg.set("loaded", LuaTable())
        pkg.set("path", ";".join(str(d / "?.lua") for d in self.search_dirs))
        G.set("package", pkg)


def _packed(values):
    t = LuaTable()
    for i, v in enumerate(values, 1):
        t.set(i, v)
    t.set("n", len(values))
    return t
