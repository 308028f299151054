"""Host-side mirror of the reference's ``luascad`` crate for the accelerated path.

* ``geom``  — the constructors and methods of ``luascad::geom::LObject``
  (luascad/src/lib.rs:84-287) with the same names, arguments and tree shapes
  (``cube`` = intersection of six axis planes, ``cylinder`` = cone|cylinder between two z
  planes, booleans are binary).  Every call records one node on a scene *backend*; the
  product backend is :class:`Tape` (the C ABI's ``mc_tape_*``), tests plug the oracle in.
* ``eval`` — ``luascad::eval(&str) -> HashMap<String, Output>`` (luascad/src/lib.rs:328-332)
  for the Lua subset the example scenes use.  No Lua interpreter exists in this environment
  (SURVEY.md), so a small recursive-descent evaluator handles assignments, arithmetic,
  ``geom.*`` calls, ``obj:method(...)`` chains and the final ``return { name = ..., }``.
"""
import ctypes as C
import math
import re
import sys

from . import _lib

EPSILON = sys.float_info.epsilon


class Tape:
    """Scene backend on top of the C ABI (one ``mc_tape``)."""

    def __init__(self):
        self._lib = _lib.lib()
        self.handle = self._lib.mc_tape_new()
        if not self.handle:
            raise MemoryError("mc_tape_new failed")

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self._lib.mc_tape_free(h)

    @staticmethod
    def _ids(ids):
        return (C.c_int32 * len(ids))(*ids), len(ids)

    def plane(self, axis, negative, d):
        return _lib.check(self._lib.mc_tape_plane(self.handle, axis, int(negative), d))

    def plane_hesian(self, nx, ny, nz, p):
        return _lib.check(self._lib.mc_tape_plane_hesian(self.handle, nx, ny, nz, p))

    def plane_3_points(self, a, b, c):
        arr = lambda v: (C.c_double * 3)(*v)
        return _lib.check(self._lib.mc_tape_plane_3_points(self.handle, arr(a), arr(b), arr(c)))

    def sphere(self, r):
        return _lib.check(self._lib.mc_tape_sphere(self.handle, r))

    def i_cylinder(self, r):
        return _lib.check(self._lib.mc_tape_i_cylinder(self.handle, r))

    def i_cone(self, slope, offset):
        return _lib.check(self._lib.mc_tape_i_cone(self.handle, slope, offset))

    def unit_sphere_sq(self):
        return _lib.check(self._lib.mc_tape_unit_sphere_sq(self.handle))

    def set_bbox(self, node, bbox):
        _lib.check(self._lib.mc_tape_set_bbox(self.handle, node, (C.c_double * 6)(*bbox)))

    def union(self, ids, r):
        a, n = self._ids(ids)
        return _lib.check(self._lib.mc_tape_union(self.handle, a, n, r))

    def intersection(self, ids, r):
        a, n = self._ids(ids)
        return _lib.check(self._lib.mc_tape_intersection(self.handle, a, n, r))

    def difference(self, ids, r):
        a, n = self._ids(ids)
        return _lib.check(self._lib.mc_tape_difference(self.handle, a, n, r))

    def translate(self, node, x, y, z):
        return _lib.check(self._lib.mc_tape_translate(self.handle, node, x, y, z))

    def rotate(self, node, x, y, z):
        return _lib.check(self._lib.mc_tape_rotate(self.handle, node, x, y, z))

    def scale(self, node, x, y, z):
        return _lib.check(self._lib.mc_tape_scale(self.handle, node, x, y, z))

    def bend(self, node, w):
        return _lib.check(self._lib.mc_tape_bend(self.handle, node, w))

    def twist(self, node, h):
        return _lib.check(self._lib.mc_tape_twist(self.handle, node, h))

    def set_parameters(self, fade_range, r_multiplier):
        _lib.check(self._lib.mc_tape_set_parameters(self.handle, fade_range, r_multiplier))

    def bbox(self, node):
        out = (C.c_double * 6)()
        _lib.check(self._lib.mc_tape_bbox(self.handle, node, out))
        return list(out)

    def program_info(self, node, slack):
        words, flops = C.c_uint64(), C.c_uint64()
        vd, pd = C.c_int32(), C.c_int32()
        _lib.check(self._lib.mc_tape_program_info(self.handle, node, slack, C.byref(words), C.byref(vd),
                                                  C.byref(pd), C.byref(flops)))
        return {"words": words.value, "value_depth": vd.value, "point_depth": pd.value, "flops": flops.value}


    def program_hash(self, node, slack):
        """Content hash of the device program for (node, slack) and the node's bounding box."""
        h = C.c_uint64()
        _lib.check(self._lib.mc_tape_program_hash(self.handle, node, slack, C.byref(h)))
        return h.value


class LObject:
    """``luascad::geom::LObject``: a handle to one node of the scene (luascad/src/lib.rs:84-86)."""

    def __init__(self, backend, node):
        self.backend = backend
        self.node = node

    # --- methods (luascad/src/lib.rs:221-281) ---
    def translate(self, x, y, z):
        return LObject(self.backend, self.backend.translate(self.node, float(x), float(y), float(z)))

    def rotate(self, x, y, z):
        return LObject(self.backend, self.backend.rotate(self.node, float(x), float(y), float(z)))

    def scale(self, x, y, z):
        return LObject(self.backend, self.backend.scale(self.node, float(x), float(y), float(z)))

    def bend(self, width):
        return LObject(self.backend, self.backend.bend(self.node, float(width)))

    def twist(self, height):
        return LObject(self.backend, self.backend.twist(self.node, float(height)))

    def union(self, other, smooth):
        return LObject(self.backend, self.backend.union([self.node, other.node], float(smooth)))

    def difference(self, other, smooth):
        return LObject(self.backend, self.backend.difference([self.node, other.node], float(smooth)))

    def intersection(self, other, smooth):
        return LObject(self.backend, self.backend.intersection([self.node, other.node], float(smooth)))

    def volume(self):
        return Output("Volume", self)

    def boundary(self):
        return Output("Boundary", self)

    def bbox(self):
        return self.backend.bbox(self.node)


class Output:
    """``luascad::geom::Output`` (luascad/src/lib.rs:289-326)."""

    def __init__(self, kind, obj):
        self.kind = kind
        self.obj = obj

    def volume(self):
        return self.obj if self.kind == "Volume" else None

    def boundary(self):
        return self.obj if self.kind == "Boundary" else None

    def object_mut(self):
        return self.obj


class Geom:
    """The ``geom`` table exposed to scripts (constructors of luascad/src/lib.rs:89-219)."""

    def __init__(self, backend=None):
        self.backend = backend if backend is not None else Tape()

    def _o(self, node):
        return LObject(self.backend, node)

    def plane_x(self, x): return self._o(self.backend.plane(0, False, float(x)))
    def plane_y(self, y): return self._o(self.backend.plane(1, False, float(y)))
    def plane_z(self, z): return self._o(self.backend.plane(2, False, float(z)))
    def plane_neg_x(self, x): return self._o(self.backend.plane(0, True, float(x)))
    def plane_neg_y(self, y): return self._o(self.backend.plane(1, True, float(y)))
    def plane_neg_z(self, z): return self._o(self.backend.plane(2, True, float(z)))

    def plane_hesian(self, nx, ny, nz, p):
        return self._o(self.backend.plane_hesian(float(nx), float(ny), float(nz), float(p)))

    def plane_3_points(self, ax, ay, az, bx, by, bz, cx, cy, cz):
        f = float
        return self._o(self.backend.plane_3_points((f(ax), f(ay), f(az)), (f(bx), f(by), f(bz)), (f(cx), f(cy), f(cz))))

    def sphere(self, radius): return self._o(self.backend.sphere(float(radius)))
    def i_cylinder(self, radius): return self._o(self.backend.i_cylinder(float(radius)))
    def i_cone(self, slope): return self._o(self.backend.i_cone(float(slope), 0.0))

    def cube(self, x, y, z, smooth):
        # luascad/src/lib.rs:173-187
        x, y, z, smooth = float(x), float(y), float(z), float(smooth)
        b = self.backend
        planes = [b.plane(0, False, x / 2.0), b.plane(1, False, y / 2.0), b.plane(2, False, z / 2.0),
                  b.plane(0, True, x / 2.0), b.plane(1, True, y / 2.0), b.plane(2, True, z / 2.0)]
        return self._o(b.intersection(planes, smooth))

    def cylinder(self, length, r1, r2, smooth):
        # luascad/src/lib.rs:190-219
        length, r1, r2, smooth = float(length), float(r1), float(r2), float(smooth)
        b = self.backend
        if abs(r1 - r2) < EPSILON:
            cone = b.i_cylinder(r1)
        else:
            slope = abs(r2 - r1) / length
            offset = (-r1 / slope - length * 0.5) if r1 < r2 else (r2 / slope + length * 0.5)
            cone = b.i_cone(slope, offset)
            rmax = max(r1, r2)
            b.set_bbox(cone, [-rmax, -rmax, -math.inf, rmax, rmax, math.inf])
        return self._o(b.intersection([cone, b.plane(2, False, length / 2.0), b.plane(2, True, length / 2.0)], smooth))


# ---------------------------------------------------------------------------------------
# Lua subset evaluator

class LuaError(Exception):
    """Stands in for ``hlua::LuaError`` (luascad/src/lib.rs:328)."""


_TOKEN = re.compile(r"""
    (?P<ws>\s+|--[^\n]*)
  | (?P<num>0[xX][0-9a-fA-F]+|\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)
  | (?P<name>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<str>"[^"\n]*"|'[^'\n]*')
  | (?P<op>==|~=|<=|>=|\.\.|[-+*/^%(){}\[\]=,;:.<>#])
""", re.X)

_MATH = {"pi": math.pi, "huge": math.inf, "abs": abs, "sqrt": math.sqrt, "sin": math.sin, "cos": math.cos,
         "tan": math.tan, "asin": math.asin, "acos": math.acos, "atan": math.atan, "atan2": math.atan2,
         "floor": math.floor, "ceil": math.ceil, "max": max, "min": min, "exp": math.exp, "log": math.log,
         "pow": pow, "rad": math.radians, "deg": math.degrees, "fmod": math.fmod}


class _Parser:
    def __init__(self, src, env):
        self.toks = []
        pos = 0
        while pos < len(src):
            m = _TOKEN.match(src, pos)
            if not m:
                raise LuaError(f"syntax error near {src[pos:pos + 20]!r}")
            pos = m.end()
            if m.lastgroup != "ws":
                self.toks.append((m.lastgroup, m.group()))
        self.i = 0
        self.env = env

    def peek(self, k=0):
        return self.toks[self.i + k] if self.i + k < len(self.toks) else ("eof", "")

    def next(self):
        t = self.peek()
        self.i += 1
        return t

    def accept(self, val):
        if self.peek()[1] == val and self.peek()[0] in ("op", "name"):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise LuaError(f"expected {val!r} near {self.peek()[1]!r}")

    # chunk ::= {stat [';']} ; returns the value of the `return` statement (or None)
    def chunk(self):
        while self.peek()[0] != "eof":
            if self.accept(";"):
                continue
            if self.accept("return"):
                val = self.expr() if self.peek()[0] != "eof" and self.peek()[1] != ";" else None
                self.accept(";")
                if self.peek()[0] != "eof":
                    raise LuaError("'return' must be the last statement")
                return val
            self.accept("local")
            kind, name = self.next()
            if kind != "name":
                raise LuaError(f"unexpected {name!r}")
            if self.peek()[1] == "=":
                self.next()
                self.env[name] = self.expr()
            else:  # expression statement (a call)
                self.i -= 1
                self.expr()
        return None

    def expr(self):
        return self.additive()

    def additive(self):
        v = self.term()
        while self.peek()[1] in ("+", "-") and self.peek()[0] == "op":
            op = self.next()[1]
            r = self.term()
            v = v + r if op == "+" else v - r
        return v

    def term(self):
        v = self.unary()
        while self.peek()[1] in ("*", "/", "%") and self.peek()[0] == "op":
            op = self.next()[1]
            r = self.unary()
            v = v * r if op == "*" else (v / r if op == "/" else math.fmod(v, r))
        return v

    def unary(self):
        if self.peek() == ("op", "-"):
            self.next()
            return -self.unary()
        return self.power()

    def power(self):
        v = self.postfix()
        if self.peek() == ("op", "^"):
            self.next()
            return float(v) ** float(self.unary())
        return v

    def args(self):
        self.expect("(")
        out = []
        if not self.accept(")"):
            out.append(self.expr())
            while self.accept(","):
                out.append(self.expr())
            self.expect(")")
        return out

    def postfix(self):
        kind, tok = self.next()
        if kind == "num":
            v = float(int(tok, 16)) if tok[:2].lower() == "0x" else float(tok)
        elif kind == "str":
            v = tok[1:-1]
        elif kind == "name":
            if tok in ("true", "false"):
                v = tok == "true"
            elif tok == "nil":
                v = None
            elif tok in self.env:
                v = self.env[tok]
            else:
                raise LuaError(f"attempt to use undefined variable {tok!r}")
        elif (kind, tok) == ("op", "("):
            v = self.expr()
            self.expect(")")
        elif (kind, tok) == ("op", "{"):
            v = {}
            n = 1
            while not self.accept("}"):
                if self.peek()[0] == "name" and self.peek(1) == ("op", "="):
                    key = self.next()[1]
                    self.next()
                    v[key] = self.expr()
                else:
                    v[n] = self.expr()
                    n += 1
                if not (self.accept(",") or self.accept(";")):
                    self.expect("}")
                    break
        else:
            raise LuaError(f"unexpected {tok!r}")
        while True:
            if self.peek() == ("op", "."):
                self.next()
                key = self.next()[1]
                v = v[key] if isinstance(v, dict) else getattr(v, key)
            elif self.peek() == ("op", ":"):
                self.next()
                meth = self.next()[1]
                v = getattr(v, meth)(*self.args())
            elif self.peek() == ("op", "("):
                v = v(*self.args())
            else:
                return v


def eval(script, backend=None):  # noqa: A001 - mirrors luascad::eval
    """Evaluate a scene script; returns ``{name: Output}`` like ``luascad::eval``.

    ``backend`` defaults to a fresh :class:`Tape`; all objects of one script share it.
    """
    g = Geom(backend)
    env = {"geom": {n: getattr(g, n) for n in dir(g) if not n.startswith("_") and n != "backend"},
           "math": dict(_MATH), "print": print, "tostring": str, "tonumber": float}
    try:
        result = _Parser(script, env).chunk()
    except LuaError:
        raise
    except _lib.McError:
        raise
    except Exception as e:  # arity / type errors inside geom calls
        raise LuaError(str(e)) from e
    if not isinstance(result, dict) or not all(isinstance(v, Output) for v in result.values()):
        raise LuaError("script must return a table of name = object:volume() / object:boundary()")
    return {str(k): v for k, v in result.items()}
