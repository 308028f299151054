"""Reference interpreter (plain Python, one point at a time) of the flat SDF program of
moosecad_b200/csrc/mc_program.h, as the host lowering (mc_tape.cpp) emits it.  Test infrastructure: it lets the
CPU suite check the LOWERING — opcodes, parameters, per-node slack constants, composed matrices, jump targets —
against the oracle's recursive approx_value without a GPU.  The arithmetic mirrors csrc/mc_vm.cuh in its
non-uniform mode (a failed guard jumps, a passed one runs the subtree), operation for operation; exp / ln are
the portable ones the device uses (oracle libm_mode 1)."""
import math
import struct

OP = dict(END=0, PLANE=1, NPLANE=2, USQ=3, CONE=4, GSPHERE=5, GCYL=6, GUARD=7, ENDG=8, NEG=9, BOOL=10, APUSH=11,
          APOP=12, CHILD=13, BOXDIST=14, SPHERE=15, CYL=16, GBOX=17, BOX=18, GAPUSH=19)
INF = float("inf")


def _f(w):
    return struct.unpack("<d", struct.pack("<Q", int(w)))[0]


def _signbit(x):
    return struct.unpack("<q", struct.pack("<d", x))[0] < 0


class Interp:
    def __init__(self, words, exp, ln):
        self.w = [int(x) for x in words]
        self.exp, self.ln = exp, ln

    def f(self, i):
        return _f(self.w[i])

    def box_distance(self, at, x, y, z):
        xv = max(x - self.f(at + 3), self.f(at + 0) - x)
        yv = max(y - self.f(at + 4), self.f(at + 1) - y)
        zv = max(z - self.f(at + 5), self.f(at + 2) - z)
        return max(xv, max(yv, zv))

    def fold(self, vals, is_union, r, er):
        close, m = False, (INF if is_union else -INF)
        for x in vals:
            if is_union:
                if x < m:
                    close = (m - x) < er
                    m = x
                elif (x - m) < er:
                    close = True
            else:
                if x > m:
                    close = (x - m) < er
                    m = x
                elif (m - x) < er:
                    close = True
        if not close:
            return m
        lim = (m + r) if is_union else (m - r)
        r4 = r / 4.0
        s = 0.0
        for x in vals:
            if is_union:
                if x < lim:
                    s = s + self.exp(-x / r4)
            elif x > lim:
                s = s + self.exp(x / r4)
        return self.ln(s) * -r4 if is_union else self.ln(s) * r4

    @staticmethod
    def plane(q, d, inverted):
        ap = abs(q)
        return (ap - d) if ((not _signbit(q)) != inverted) else (-ap - d)

    def box_fold(self, at, mask, x, y, z):
        vals = []
        for i in range(6):
            if (mask >> i) & 1:
                vals.append(self.plane((x, y, z)[i % 3], self.f(at + i), i >= 3))
        return self.fold(vals, False, self.f(at + 6), self.f(at + 7))

    def __call__(self, x, y, z):
        w, vs, ps, pc = self.w, [], [], 0
        for _ in range(10 * len(w) + 10):
            h = w[pc]
            op, small, idx = h & 0xff, (h >> 8) & 0xff, h >> 40
            if op == OP["END"]:
                assert len(vs) == 1 and not ps
                return vs[0]
            if op == OP["PLANE"]:
                vs.append(self.plane((x, y, z)[small & 3], self.f(pc + 1), bool(small & 4)))
                pc += 2
            elif op == OP["NPLANE"]:
                vs.append(((x * self.f(pc + 1) + y * self.f(pc + 2)) + z * self.f(pc + 3)) - self.f(pc + 4))
                pc += 5
            elif op == OP["USQ"]:
                vs.append(((x * x + y * y) + z * z) - 1.0)
                pc += 1
            elif op == OP["CONE"]:
                radius = abs(self.f(pc + 1) * (z + self.f(pc + 2)))
                vs.append((math.sqrt(x * x + y * y) - radius) * self.f(pc + 3))
                pc += 4
            elif op in (OP["GSPHERE"], OP["GCYL"]):
                a = self.box_distance(pc + 1, x, y, z)
                if a <= self.f(pc + 7):
                    r = self.f(pc + 8)
                    a = (math.sqrt((x * x + y * y) + z * z) - r) if op == OP["GSPHERE"] else (math.sqrt(x * x + y * y) - r)
                vs.append(a)
                pc += 9
            elif op == OP["GBOX"]:
                a = self.box_distance(pc + 1, x, y, z)
                vs.append(self.box_fold(pc + 8, small, x, y, z) if a <= self.f(pc + 7) else a)
                pc += 16
            elif op == OP["BOX"]:
                vs.append(self.box_fold(pc + 1, small, x, y, z))
                pc += 9
            elif op in (OP["GUARD"], OP["GAPUSH"]):
                a = self.box_distance(pc + 1, x, y, z)
                if not (a <= self.f(pc + 7)):
                    vs.append(a)
                    pc = idx                      # past the matching ENDG
                    continue
                pc += 8
                if op == OP["GAPUSH"]:
                    ps.append((x, y, z))
                    x, y, z = self.affine(pc - 1, x, y, z)
                    pc += 12
            elif op == OP["APUSH"]:
                ps.append((x, y, z))
                x, y, z = self.affine(pc, x, y, z)
                pc += 13
            elif op == OP["APOP"]:
                x, y, z = ps.pop()
                vs[-1] = vs[-1] * self.f(pc + 1)
                pc += 2
            elif op in (OP["ENDG"], OP["CHILD"]):
                pc += 1                           # only points that passed the guard get here
            elif op == OP["NEG"]:
                vs[-1] = -vs[-1]
                pc += 1
            elif op == OP["BOOL"]:
                k = small & 0x7f
                vals = vs[len(vs) - k:]
                del vs[len(vs) - k:]
                vs.append(self.fold(vals, bool(small & 0x80), self.f(pc + 1), self.f(pc + 2)))
                pc += 3
            elif op == OP["BOXDIST"]:
                vs.append(self.box_distance(pc + 1, x, y, z))
                pc += 7
            elif op == OP["SPHERE"]:
                vs.append(math.sqrt((x * x + y * y) + z * z) - self.f(pc + 1))
                pc += 2
            elif op == OP["CYL"]:
                vs.append(math.sqrt(x * x + y * y) - self.f(pc + 1))
                pc += 2
            else:
                raise ValueError(f"unknown opcode {op} at word {pc}")
        raise RuntimeError("program does not terminate")

    def affine(self, at, x, y, z):
        """m[12] (3 x 4, row major) starts at word at + 1: ((m0 x + m1 y) + m2 z) + t per row (nalgebra transform_point)."""
        f = self.f
        return (((f(at + 1) * x + f(at + 2) * y) + f(at + 3) * z) + f(at + 4),
                ((f(at + 5) * x + f(at + 6) * y) + f(at + 7) * z) + f(at + 8),
                ((f(at + 9) * x + f(at + 10) * y) + f(at + 11) * z) + f(at + 12))
