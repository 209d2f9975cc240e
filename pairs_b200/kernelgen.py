"""Generic kernels: CUDA source for kernel bodies that are not one of the hand-written families.

The reference traces a kernel function into IR and prints code for it (src/pairs/mapping/funcs.py:39-334 BuildParticleIR,
keywords in src/pairs/mapping/keywords.py, Apply in src/pairs/ir/apply.py, the pair loop of sim/interaction.py:168-295).  This
module prints CUDA for the same vocabulary directly from the Python AST, for the properties the MD path stores on the device
(position, one velocity-like vector, one volatile force-like vector, mass, the `type` feature and its feature properties) and
for every further real / vector property the script declares (rows of the user-property block, csrc/props.cu):

  pair kernels     def k(i, j): locals, delta(i, j), squared_distance(i, j) (or the legacy bare names delta / rsq),
                   prop[i], prop[j], featprop[i, j], sqrt, select, min, max, abs, dot, length, squared_length, normalized,
                   zero_vector, vector(x, y, z), + - * / unary -, comparisons, apply(prop, expr), local op= expr,
                   if / else (mapping/funcs.py:179-195; a local assigned inside an arm lives in that arm), v[0] / prop[i][2]
                   (one component of a vector), cross, skip_when(cond) (keywords.py:62-65: `continue` with the next partner),
                   is_point_mass / is_sphere / is_halfspace(i | j), the integer properties uid / shape / flags and the feature
                   (`type`) as values, % & | ^ ~ on integers, and / or / not (both sides evaluated, as the reference prints them)
  particle kernels def k(i): the same expressions and statements, prop[i] = / += / -= expr

  contact models   def k(i, j) of a DEM script (translate_dem_model): the same expressions plus penetration_depth / contact_point /
                   contact_normal(i, j), contact properties cp[i, j] (read and assigned, mapping/funcs.py:230-263), apply() to
                   force and torque; printed as a device function that the library's contact kernel calls per touching pair
                   (csrc/dem_force_kernel.cuh)

Like the reference's generated code the output has ONE statement per operation, in Python's evaluation order, vectors
scalarised per component, `select` evaluating both arms, symbols substituted as literals; compiled with --fmad=false every
operation is the IEEE operation of the reference's C++ (-ffp-contract=off), so a pair term is bit-identical and only the
order in which a particle's terms are summed is ours (the list order).  apply() accumulates in registers and adds to the
property once after the loop (sim/interaction.py:280-292); FIXED particles are skipped (mapping/funcs.py:305-310).
"""
import ast
import inspect
import os
import textwrap


class KernelGenError(Exception):
    pass


# Pair kernels over neighbour lists fetch this many partners at a time -- their indices first, then their positions (independent
# 256-bit gathers in flight), then the bodies in list order -- as the hand-written Lennard-Jones kernel does (csrc/md_kernels.cu,
# `lj_unroll`); the order of the pair terms, hence every bit of the result, is that of the plain loop (1 = plain loop).
PAIR_PREFETCH = int(os.environ.get("PAIRS_B200_PAIR_PREFETCH", "4"))


def _lit(x):
    if isinstance(x, bool):
        return "true" if x else "false"
    if isinstance(x, int):
        return str(x)
    r = repr(float(x))
    if r in ("inf", "-inf"):
        return "PB_INFINITY" if r == "inf" else "(-PB_INFINITY)"
    if r == "nan":
        raise KernelGenError("NaN symbol value")
    return r if ("." in r or "e" in r or "E" in r) else r + ".0"


def _xref(store, d, idx):
    """Component d of a user-defined property (csrc/props.cu: SoA rows of [cap])."""
    return f"a.xdata[{store[1] + d} * (size_t) a.cap + {idx}]"


def _store_tag(store):
    return store if isinstance(store, str) else f"x{store[1]}"


class _Gen:
    """Expression / statement printer.  Values are (type, code) with type in {'f', 'i', 'b'} or ('v', [c0, c1, c2])."""

    def __init__(self, name, kind, storage, feature_tables, ntypes, symbols, glob):
        self.name, self.kind = name, kind
        self.storage = storage                # property name -> 'pos' | 'vel' | 'force' | 'mass' | ('x', first row, components)
        self.tables = feature_tables          # feature property name -> list of nk*nk floats
        self.ntypes = ntypes
        self.symbols, self.glob = symbols, glob
        self.lines, self.locals, self.n = [], {}, 0
        self.loaded = {}                      # (storage, who) -> value, loaded once per (pair) iteration
        self.applied = {}                     # storage -> accumulator names
        self.hoisted = []                     # statements before the neighbour loop (loads of i)
        self.types_needed = False
        self.stored = set()                   # properties written so far (particle kernels)
        self.mutable = {}                     # locals that an if / else arm assigns: name -> value held in mutable C variables
        self.contact = {}                     # contact property name -> 'c_tsd' | 'c_ivm' | 'c_stick' (contact models)
        self.depth = 0                        # nesting depth of if / else arms

    # -- helpers --
    def tmp(self, ctype, code, hoist=False):
        self.n += 1
        t = f"t{self.n}"
        (self.hoisted if hoist else self.lines).append(f"const {ctype} {t} = {code};")
        return t

    def vec(self, comps):
        return ("v", list(comps))

    @staticmethod
    def is_vec(v):
        return isinstance(v, tuple) and v[0] == "v"

    @staticmethod
    def is_mat(v):                           # 3x3 matrix, row-major, 9 components
        return isinstance(v, tuple) and v[0] == "m"

    @staticmethod
    def is_quat(v):                          # quaternion (w, x, y, z)
        return isinstance(v, tuple) and v[0] == "q"

    def sum3(self, terms):
        """(t0 + t1) + t2 (+ t3): Python's left-to-right sum of the reference's keyword implementations."""
        acc = terms[0]
        for t in terms[1:]:
            acc = self.tmp("double", f"{acc} + {t}")
        return acc

    def mul(self, x, y):
        return self.tmp("double", f"{x} * {y}")

    def matmul(self, a, b):
        """keywords.py:158-196 (matrix_multiplication) and :219-235 (quaternion_multiplication)."""
        if self.is_quat(a) and self.is_quat(b):
            l, r = a[1], b[1]

            def comb(t0, t1, t2, t3, signs):
                acc = self.mul(l[t0[0]], r[t0[1]])
                for (x, y), sg in zip((t1, t2, t3), signs):
                    acc = self.tmp("double", f"{acc} {sg} {self.mul(l[x], r[y])}")
                return acc
            rr = comb((0, 0), (1, 1), (2, 2), (3, 3), "---")
            ii = comb((0, 1), (1, 0), (2, 3), (3, 2), "++-")
            jj = comb((0, 2), (2, 0), (3, 1), (1, 3), "++-")
            kk = comb((0, 3), (3, 0), (1, 2), (2, 1), "++-")
            len2 = self.sum3([self.mul(x, x) for x in (rr, ii, jj, kk)])
            off = self.tmp("double", f"{len2} - 1.0")
            near = self.tmp("bool", f"fabs({off}) < 1e-08")
            root = self.tmp("double", f"sqrt({len2})")
            inv = self.tmp("double", f"1.0 / {root}")
            ilen = self.tmp("double", f"({near}) ? (1.0) : ({inv})")
            return ("q", [self.mul(x, ilen) for x in (rr, ii, jj, kk)])
        if self.is_mat(a) and self.is_mat(b):
            l, r = a[1], b[1]
            return ("m", [self.sum3([self.mul(l[3 * i + k], r[3 * k + j]) for k in range(3)]) for i in range(3) for j in range(3)])
        if self.is_mat(a) and self.is_vec(b):
            return self.vec([self.sum3([self.mul(a[1][3 * i + k], b[1][k]) for k in range(3)]) for i in range(3)])
        if self.is_vec(a) and self.is_mat(b):      # as the reference defines it: out[i] = sum_k v[k] * M[3 i + k]
            return self.vec([self.sum3([self.mul(a[1][k], b[1][3 * i + k]) for k in range(3)]) for i in range(3)])
        if self.is_mat(b) and not isinstance(a[1], list):
            return ("m", [self.mul(x, a[1]) for x in b[1]])
        if self.is_mat(a) and not isinstance(b[1], list):
            return ("m", [self.mul(x, b[1]) for x in a[1]])
        raise KernelGenError("unsupported product of matrices / quaternions")

    def load(self, store, who):
        key = (store, who)
        if key in self.loaded:
            return self.loaded[key]
        idx = "i" if who == "i" else "j"
        hoist = who == "i" and self.kind == "pair"
        if self.kind == "dem":                                # contact model: the kernel hands the pair's data in as arguments
            names = {"pos": "x", "vel": "v", "angvel": "w"}
            if store in names:
                val = self.vec([f"{names[store]}{idx}[{d}]" for d in range(3)])
            elif store in ("mass", "radius"):
                val = ("f", f"{'m' if store == 'mass' else 'r'}{idx}")
            else:
                raise KernelGenError(f"'{store}' is not available inside a contact model (position, linear and angular velocity, mass, radius)")
            self.loaded[key] = val
            return val
        if store == "pos":
            src = "pi" if who == "i" else "pj"
            if who == "i" and self.kind != "pair":
                # a per-particle kernel may assign position[i]: what was read before the assignment keeps its value (as every other
                # property: snapshots in const temporaries), `old = position[i]; position[i] = ...; position[i] - old` is not 0
                val = self.vec([self.tmp("double", f"{src}.{c}", False) for c in "xyz"])
            else:
                val = self.vec([f"{src}.x", f"{src}.y", f"{src}.z"])
        elif store in ("vel", "force", "angvel", "torque"):       # angvel / torque: DEM scripts
            val = self.vec([self.tmp("double", f"a.{store}[{d} * (size_t) a.cap + {idx}]" if d else f"a.{store}[{idx}]", hoist) for d in range(3)])
        elif store in ("mass", "radius"):
            val = ("f", self.tmp("double", f"a.{store}[{idx}]", hoist))
        elif store in ("inv_inertia", "rotmat", "quat"):      # DEM scripts: matrix / quaternion properties, SoA rows of [cap]
            n = 4 if store == "quat" else 9
            val = ("q" if store == "quat" else "m",
                   [self.tmp("double", f"a.{store}[{d} * (size_t) a.cap + {idx}]" if d else f"a.{store}[{idx}]", hoist) for d in range(n)])
        elif store in ("uid", "shape", "flags"):
            val = ("i", self.tmp("int", f"a.{store}[{idx}]", hoist))
        elif store == "type":                                 # the feature index rides in the low bits of position.w
            val = ("i", self.tmp("int", f"pb_w_type({'pi' if who == 'i' else 'pj'}.w)", hoist))
        elif isinstance(store, tuple):                        # user-defined property: rows of a.xdata
            if len(store) > 3:                                # integer property kept in a double row
                val = ("i", self.tmp("int", f"(int) {_xref(store, 0, idx)}", hoist))
            else:
                comps = [self.tmp("double", _xref(store, d, idx), hoist) for d in range(store[2])]
                val = ("f", comps[0]) if store[2] == 1 else self.vec(comps)
        else:
            raise KernelGenError(f"no device storage for '{store}'")
        self.loaded[key] = val
        return val

    # -- expressions --
    def expr(self, node):
        if isinstance(node, ast.Constant):
            if isinstance(node.value, (int, float)) and not isinstance(node.value, bool):
                return ("i" if isinstance(node.value, int) else "f", _lit(node.value))
            raise KernelGenError(f"unsupported constant {node.value!r}")
        if isinstance(node, ast.Name):
            return self.name_value(node.id)
        if isinstance(node, ast.UnaryOp):
            v = self.expr(node.operand)
            if isinstance(node.op, ast.USub):
                if self.is_vec(v):
                    return self.vec([self.tmp("double", f"-({c})") for c in v[1]])
                return (v[0], self.tmp("double" if v[0] == "f" else "int", f"-({v[1]})"))
            if isinstance(node.op, ast.UAdd):
                return v
            if isinstance(node.op, ast.Not):
                return ("b", self.tmp("bool", f"!({v[1]})"))
            if isinstance(node.op, ast.Invert) and v[0] == "i":
                return ("i", self.tmp("int", f"~({v[1]})"))
            raise KernelGenError("unsupported unary operator")
        if isinstance(node, ast.BinOp):
            ops = {ast.Add: "+", ast.Sub: "-", ast.Mult: "*", ast.Div: "/"}
            int_ops = {ast.Mod: "%", ast.BitAnd: "&", ast.BitOr: "|", ast.BitXor: "^"}       # mapping/funcs.py:54-69, integers only
            if type(node.op) in int_ops:
                a, b = self.expr(node.left), self.expr(node.right)
                if self.is_vec(a) or self.is_vec(b) or a[0] != "i" or b[0] != "i":
                    raise KernelGenError(f"operator {int_ops[type(node.op)]} needs integer operands")
                return ("i", self.tmp("int", f"{a[1]} {int_ops[type(node.op)]} {b[1]}"))
            if type(node.op) not in ops:
                raise KernelGenError(f"unsupported operator {type(node.op).__name__}")
            return self.binop(ops[type(node.op)], self.expr(node.left), self.expr(node.right))
        if isinstance(node, ast.Compare):
            if len(node.ops) != 1:
                raise KernelGenError("chained comparisons are not supported")
            ops = {ast.Lt: "<", ast.LtE: "<=", ast.Gt: ">", ast.GtE: ">=", ast.Eq: "==", ast.NotEq: "!="}
            a, b = self.expr(node.left), self.expr(node.comparators[0])
            if self.is_vec(a) or self.is_vec(b):
                raise KernelGenError("vectors cannot be compared")
            return ("b", self.tmp("bool", f"{a[1]} {ops[type(node.ops[0])]} {b[1]}"))
        if isinstance(node, ast.BoolOp):
            vals = [self.expr(v) for v in node.values]
            op = "&&" if isinstance(node.op, ast.And) else "||"
            return ("b", self.tmp("bool", f" {op} ".join(v[1] for v in vals)))
        if isinstance(node, ast.Subscript):
            return self.subscript(node)
        if isinstance(node, ast.Call):
            return self.call(node)
        raise KernelGenError(f"unsupported expression {type(node).__name__}")

    def binop(self, op, a, b):
        if any(self.is_mat(x) or self.is_quat(x) for x in (a, b)):
            if op != "*":
                raise KernelGenError("matrices and quaternions: only * is defined")
            return self.matmul(a, b)
        if self.is_vec(a) and self.is_vec(b):
            if op in ("*", "/"):
                raise KernelGenError("vector * vector is ambiguous: use dot()")
            return self.vec([self.tmp("double", f"{x} {op} {y}") for x, y in zip(a[1], b[1])])
        if self.is_vec(a):
            if op in ("+", "-"):
                raise KernelGenError("vector +- scalar is not defined")
            return self.vec([self.tmp("double", f"{x} {op} {b[1]}") for x in a[1]])
        if self.is_vec(b):
            if op != "*":
                raise KernelGenError("scalar op vector: only * is defined")
            return self.vec([self.tmp("double", f"{a[1]} {op} {y}") for y in b[1]])
        # two integers stay an integer, `/` included: the reference types the operation by its operands (ir/scalars.py:66-91) and
        # prints C, so 7 / 2 is 3 there -- not Python's 3.5
        t = "i" if (a[0] == "i" and b[0] == "i") else "f"
        return (t, self.tmp("double" if t == "f" else "int", f"{a[1]} {op} {b[1]}"))

    def name_value(self, name):
        if name in self.locals:
            v = self.locals[name]
            if name in self.mutable:                          # a variable that if / else arms store into: read = snapshot of its value
                ctype = {"f": "double", "i": "int", "b": "bool"}.get(v[0], "double")
                return (v[0], [self.tmp("double", c) for c in v[1]]) if isinstance(v[1], list) else (v[0], self.tmp(ctype, v[1]))
            return v
        if self.kind == "dem" and name in self.storage and name not in self.symbols:
            return self.load(self.storage[name], "i")         # a bare property inside apply() means prop[i] (ir/apply.py:99-100)
        if self.kind == "pair" and name == "rsq":            # legacy bare names of examples/lj_onetype.py
            return ("f", "rsq")
        if self.kind == "pair" and name == "delta":
            return self.vec(["dx", "dy", "dz"])
        if name in self.symbols:
            v = self.symbols[name]
        elif name in self.glob and isinstance(self.glob[name], (int, float)) and not isinstance(self.glob[name], bool):
            v = self.glob[name]
        else:
            raise KernelGenError(f"symbol '{name}' has no value (pass it in symbols={{...}})")
        return ("i" if isinstance(v, int) else "f", _lit(v))

    def subscript(self, node):
        idx = node.slice
        if isinstance(idx, ast.Constant) and isinstance(idx.value, int) and not isinstance(idx.value, bool):
            base = self.expr(node.value)                      # v[0], position[i][2]: one component (ir/vectors.py VectorAccess)
            if not isinstance(base[1], list) or not 0 <= idx.value < len(base[1]):
                raise KernelGenError("component access needs a vector / matrix / quaternion and an index inside it")
            return ("f", base[1][idx.value])
        if not isinstance(node.value, ast.Name):
            raise KernelGenError("unsupported subscript")
        prop = node.value.id
        if isinstance(idx, ast.Tuple) and self.kind == "dem":
            names = [e.id for e in idx.elts if isinstance(e, ast.Name)]
            if names != ["i", "j"]:
                raise KernelGenError(f"'{prop}[...]': contact and feature properties are indexed [i, j]")
            if prop in self.contact:                          # contact property of this pair (mapping/funcs.py:230-263)
                kind = self.contact[prop]
                if kind == "c_tsd":
                    return self.vec([self.tmp("double", f"tsd[{d}]") for d in range(3)])
                if kind.startswith("cx:"):                    # a further contact property: lanes of cx[] (pb_dem_enable_ex)
                    _, ctype, off = kind.split(":")
                    if ctype == "vec":
                        return self.vec([self.tmp("double", f"cx[{int(off) + d}]") for d in range(3)])
                    return ("f", self.tmp("double", f"cx[{off}]")) if ctype == "real" else ("i", self.tmp("int", f"(int) cx[{off}]"))
                return ("f", self.tmp("double", "*ivm")) if kind == "c_ivm" else ("i", self.tmp("int", "*sticking"))
            if prop in self.tables:
                return ("f", self.tmp("double", f"fp_{prop}[tij]"))
            raise KernelGenError(f"'{prop}' is neither a contact property nor a feature property")
        if isinstance(idx, ast.Tuple):                        # feature property: fp[i, j]
            names = [e.id for e in idx.elts if isinstance(e, ast.Name)]
            if prop not in self.tables or names != ["i", "j"] or self.kind != "pair":
                raise KernelGenError(f"'{prop}[...]': only feature_property[i, j] inside a pair kernel is supported")
            self.types_needed = True
            return ("f", self.tmp("double", f"fp_{prop}[ti + tj]"))
        if not isinstance(idx, ast.Name) or idx.id not in ("i", "j") or (idx.id == "j" and self.kind not in ("pair", "dem")):
            raise KernelGenError(f"'{prop}[...]': index must be the particle argument")
        if prop not in self.storage:
            raise KernelGenError(f"'{prop}' is not a declared real / vector property")
        return self.load(self.storage[prop], idx.id)

    def call(self, node):
        if not isinstance(node.func, ast.Name):
            raise KernelGenError("unsupported call")
        f = node.func.id
        if self.kind == "dem" and f in ("penetration_depth", "contact_point", "contact_normal", "delta", "squared_distance"):
            # sim/interaction.py:234-264: the geometry of the touching pair, computed by the kernel (dem_math.h pb_dem_geom_*)
            if f == "contact_point":
                return self.vec(["cp[0]", "cp[1]", "cp[2]"])
            if f == "contact_normal":
                return self.vec(["n[0]", "n[1]", "n[2]"])
            if f == "penetration_depth":
                return ("f", self.tmp("double", "-(delta)"))   # the kernel passes delta = -penetration_depth; negation is exact
            if "dem_delta" not in self.locals:
                self.locals["dem_delta"] = self.vec([self.tmp("double", f"xi[{d}] - xj[{d}]") for d in range(3)])
            dv = self.locals["dem_delta"]
            if f == "delta":
                return dv
            p = [self.tmp("double", f"{x} * {x}") for x in dv[1]]
            q = self.tmp("double", f"{p[0]} + {p[1]}")
            return ("f", self.tmp("double", f"{q} + {p[2]}"))
        if f in ("delta", "squared_distance"):
            if self.kind != "pair":
                raise KernelGenError(f"{f}() needs a pair kernel")
            return self.vec(["dx", "dy", "dz"]) if f == "delta" else ("f", "rsq")
        args = [] if f in ("is_point_mass", "is_sphere", "is_halfspace") else [self.expr(a) for a in node.args]
        if f in ("sqrt", "abs") and (len(args) != 1 or isinstance(args[0][1], list)):
            raise KernelGenError(f"{f}() takes one scalar")
        if f == "sqrt":
            x = args[0][1] if args[0][0] == "f" else f"(double) {args[0][1]}"        # an integer argument: no overload guessing
            return ("f", self.tmp("double", f"sqrt({x})"))
        if f == "abs":
            if args[0][0] == "i":
                return ("i", self.tmp("int", f"({args[0][1]} < 0) ? -({args[0][1]}) : ({args[0][1]})"))
            return ("f", self.tmp("double", f"fabs({args[0][1]})"))
        if f in ("min", "max"):                               # keywords.py:67-79: e = a0; for a in rest: e = select(a < e, a, e)
            if len(args) < 1 or any(self.is_vec(a) for a in args):
                raise KernelGenError(f"{f}() takes scalars")
            e = args[0]
            for a in args[1:]:
                c = self.tmp("bool", f"{a[1]} {'<' if f == 'min' else '>'} {e[1]}")
                t = "i" if (a[0] == "i" and e[0] == "i") else "f"
                e = (t, self.tmp("int" if t == "i" else "double", f"({c}) ? ({a[1]}) : ({e[1]})"))
            return e
        if f == "select":
            c, a, b = args
            if self.is_vec(a) != self.is_vec(b):
                raise KernelGenError("select(): both arms must have the same type")
            if self.is_vec(a):
                return self.vec([self.tmp("double", f"({c[1]}) ? ({x}) : ({y})") for x, y in zip(a[1], b[1])])
            if a[0] == "i" and b[0] == "i":
                return ("i", self.tmp("int", f"({c[1]}) ? ({a[1]}) : ({b[1]})"))
            return ("f", self.tmp("double", f"({c[1]}) ? ({a[1]}) : ({b[1]})"))
        if f in ("is_point_mass", "is_sphere", "is_halfspace"):      # keywords.py:34-50: shape[p] == Shapes.<...>
            if len(node.args) != 1 or not isinstance(node.args[0], ast.Name) or node.args[0].id not in ("i", "j") or \
                    (node.args[0].id == "j" and self.kind != "pair"):
                raise KernelGenError(f"{f}() takes the particle argument")
            sh = self.load("shape", node.args[0].id)
            return ("b", self.tmp("bool", f"{sh[1]} == {dict(is_sphere=0, is_halfspace=1, is_point_mass=2)[f]}"))
        if f == "transposed":                                 # keywords.py:124-130
            if not self.is_mat(args[0]):
                raise KernelGenError("transposed() takes a matrix")
            m = args[0][1]
            return ("m", [m[0], m[3], m[6], m[1], m[4], m[7], m[2], m[5], m[8]])
        if f == "inversed":                                   # keywords.py:132-149
            if not self.is_mat(args[0]):
                raise KernelGenError("inversed() takes a matrix")
            m = args[0][1]

            def minor(p, q, r, t):
                return self.tmp("double", f"{self.mul(m[p], m[q])} - {self.mul(m[r], m[t])}")
            det = self.sum3([self.mul(m[0], minor(4, 8, 7, 5)), self.mul(m[1], minor(5, 6, 8, 3)), self.mul(m[2], minor(3, 7, 4, 6))])
            inv = self.tmp("double", f"1.0 / {det}")
            idx = ((4, 8, 5, 7), (7, 2, 8, 1), (1, 5, 2, 4), (5, 6, 3, 8), (8, 0, 6, 2), (2, 3, 0, 5), (3, 7, 4, 6), (6, 1, 7, 0), (0, 4, 1, 3))
            return ("m", [self.mul(inv, minor(*t)) for t in idx])
        if f == "diagonal_matrix":                            # keywords.py:151-156
            if isinstance(args[0][1], list):
                raise KernelGenError("diagonal_matrix() takes a scalar")
            return ("m", [args[0][1] if k % 4 == 0 else "0.0" for k in range(9)])
        if f == "default_quaternion":
            return ("q", ["1.0", "0.0", "0.0", "0.0"])
        if f == "quaternion":                                 # keywords.py:202-217 (`A or B` on IR nodes keeps A: only the axis test)
            axis, angle = args
            if not self.is_vec(axis) or isinstance(angle[1], list):
                raise KernelGenError("quaternion() takes an axis vector and an angle")
            ln = self.tmp("double", f"sqrt({self.sum3([self.mul(x, x) for x in axis[1]])})")
            zero = self.tmp("bool", f"fabs({ln}) < 1e-06")
            half = self.mul(angle[1], "0.5")
            sina = self.tmp("double", f"sin({half})")
            cosa = self.tmp("double", f"cos({half})")
            inv = self.tmp("double", f"1.0 / {ln}")
            comps = [cosa] + [self.mul(sina, self.mul(x, inv)) for x in axis[1]]
            return ("q", [self.tmp("double", f"({zero}) ? ({d}) : ({c})") for d, c in zip(("1.0", "0.0", "0.0", "0.0"), comps)])
        if f == "quaternion_to_rotation_matrix":              # keywords.py:237-251
            if not self.is_quat(args[0]):
                raise KernelGenError("quaternion_to_rotation_matrix() takes a quaternion")
            q = args[0][1]

            def diag(x, y):
                first = self.tmp("double", f"1.0 - {self.mul(self.mul('2.0', q[x]), q[x])}")
                return self.tmp("double", f"{first} - {self.mul(self.mul('2.0', q[y]), q[y])}")

            def off(x, y, sg, u, v):
                inner = self.tmp("double", f"{self.mul(q[x], q[y])} {sg} {self.mul(q[u], q[v])}")
                return self.mul("2.0", inner)
            return ("m", [diag(2, 3), off(1, 2, "-", 0, 3), off(1, 3, "+", 0, 2), off(1, 2, "+", 0, 3), diag(1, 3), off(2, 3, "-", 0, 1),
                          off(1, 3, "-", 0, 2), off(2, 3, "+", 0, 1), diag(1, 2)])
        if f == "cross":                                      # keywords.py:95-103
            a, b = args
            if not (self.is_vec(a) and self.is_vec(b)):
                raise KernelGenError("cross() takes two vectors")
            out = []
            for p, q in ((1, 2), (2, 0), (0, 1)):
                l = self.tmp("double", f"{a[1][p]} * {b[1][q]}")
                r = self.tmp("double", f"{a[1][q]} * {b[1][p]}")
                out.append(self.tmp("double", f"{l} - {r}"))
            return self.vec(out)
        if f == "dot":
            a, b = args
            p = [self.tmp("double", f"{x} * {y}") for x, y in zip(a[1], b[1])]
            s = self.tmp("double", f"{p[0]} + {p[1]}")
            return ("f", self.tmp("double", f"{s} + {p[2]}"))
        if f in ("squared_length", "length", "normalized"):
            v = args[0]
            p = [self.tmp("double", f"{x} * {x}") for x in v[1]]
            s = self.tmp("double", f"{p[0]} + {p[1]}")
            sq = self.tmp("double", f"{s} + {p[2]}")
            if f == "squared_length":
                return ("f", sq)
            ln = self.tmp("double", f"sqrt({sq})")
            if f == "length":
                return ("f", ln)
            # keywords.py:105-112: select(length > 0.0, v * (1.0 / length), zero_vector())
            pos = self.tmp("bool", f"{ln} > 0.0")
            inv = self.tmp("double", f"1.0 / {ln}")
            scaled = [self.tmp("double", f"{x} * {inv}") for x in v[1]]
            return self.vec([self.tmp("double", f"({pos}) ? ({x}) : (0.0)") for x in scaled])
        if f == "zero_vector":
            return self.vec(["0.0", "0.0", "0.0"])
        if f == "vector":
            return self.vec([a[1] for a in args])
        raise KernelGenError(f"unknown function '{f}'")

    # -- statements --
    def make_mutable(self, name):
        v = self.locals[name]
        self.n += 1
        if isinstance(v[1], list):
            names = [f"m{self.n}_{d}" for d in range(len(v[1]))]
            for nm, c in zip(names, v[1]):
                self.lines.append(f"double {nm} = {c};")
            self.locals[name] = (v[0], names)
        else:
            ctype = {"f": "double", "i": "int", "b": "bool"}[v[0]]
            self.lines.append(f"{ctype} m{self.n} = {v[1]};")
            self.locals[name] = (v[0], f"m{self.n}")
        self.mutable[name] = self.locals[name]

    def assign_local(self, name, value):
        slot = self.mutable.get(name)
        if slot is None or self.depth == 0:
            self.mutable.pop(name, None)                  # re-bound at the top level: a fresh value, no storage needed
            self.locals[name] = value
            return
        if slot[0] != value[0] and not (slot[0] == "f" and value[0] == "i"):
            raise KernelGenError(f"'{name}' changes its type inside an if / else arm")
        if isinstance(slot[1], list):
            if not isinstance(value[1], list) or len(value[1]) != len(slot[1]):
                raise KernelGenError(f"'{name}' changes its type inside an if / else arm")
            # all components are evaluated before the first store (the value may be built from the old one)
            for nm, c in zip(slot[1], [self.tmp("double", c) for c in value[1]]):
                self.lines.append(f"{nm} = {c};")
        else:
            self.lines.append(f"{slot[1]} = {value[1]};")
        self.locals[name] = slot

    def block(self, body):
        """Statements of an if / else arm: temporaries, loads and locals created inside stay inside (C scope = Python use)."""
        saved_locals, saved_loaded, before, saved_mutable = dict(self.locals), dict(self.loaded), set(self.stored), dict(self.mutable)
        self.stored = set()
        self.depth += 1
        for st in body:
            self.stmt(st)
        self.depth -= 1
        written = self.stored
        self.locals = saved_locals
        self.mutable = saved_mutable                          # variables declared inside the arm end with its C scope
        # a property the arm stored to has to be read again afterwards (the store may or may not have happened)
        self.loaded = {k: v for k, v in saved_loaded.items() if k[0] not in written}
        self.stored = before | written

    def stmt(self, node):
        if isinstance(node, ast.Expr) and isinstance(node.value, ast.Constant):
            return                                            # docstring
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            self.assign_local(node.targets[0].id, self.expr(node.value))
            return
        if isinstance(node, ast.AugAssign) and isinstance(node.target, ast.Name):       # local op= expr
            if node.target.id not in self.locals:
                raise KernelGenError(f"'{node.target.id}' is used before it is assigned")
            ops = {ast.Add: "+", ast.Sub: "-", ast.Mult: "*", ast.Div: "/"}
            if type(node.op) not in ops:
                raise KernelGenError(f"unsupported operator {type(node.op).__name__}")
            self.assign_local(node.target.id, self.binop(ops[type(node.op)], self.locals[node.target.id], self.expr(node.value)))
            return
        if isinstance(node, ast.If):
            # mapping/funcs.py:179-195: Filter (one arm) / Branch (two arms).  A local FIRST assigned inside an arm is visible in
            # that arm only; what an arm does to the outside world are apply(), property stores and new values of outer locals.
            cond = self.expr(node.test)
            if self.is_vec(cond):
                raise KernelGenError("if: the condition must be a scalar")
            # a local of the enclosing scope that an arm assigns keeps the new value afterwards (Python): it becomes a mutable
            # variable here, the arms store into it
            for sub_node in [n for arm in (node.body, node.orelse) for st in arm for n in ast.walk(st)]:
                tgt = sub_node.targets[0] if isinstance(sub_node, ast.Assign) and len(sub_node.targets) == 1 else \
                    (sub_node.target if isinstance(sub_node, ast.AugAssign) else None)
                if isinstance(tgt, ast.Name) and tgt.id in self.locals and tgt.id not in self.mutable:
                    self.make_mutable(tgt.id)
            self.lines.append(f"if({cond[1]}) {{")
            self.block(node.body)
            if node.orelse:
                self.lines.append("} else {")
                self.block(node.orelse)
            self.lines.append("}")
            return
        if isinstance(node, ast.Expr) and isinstance(node.value, ast.Call) and getattr(node.value.func, "id", None) == "skip_when":
            if len(node.value.args) != 1:
                raise KernelGenError("skip_when() takes one condition")
            cond = self.expr(node.value.args[0])
            if self.is_vec(cond):
                raise KernelGenError("skip_when(): the condition must be a scalar")
            # keywords.py:62-65 prints `continue`: the next partner in a pair kernel, the next particle otherwise
            leave = {"pair": "continue", "dem": "return false"}.get(self.kind, "return")
            self.lines.append(f"if({cond[1]}) {{ {leave}; }}")
            return
        if isinstance(node, ast.Expr) and isinstance(node.value, ast.Call) and getattr(node.value.func, "id", None) == "apply":
            if self.kind == "dem":
                tgt, val = node.value.args
                out = {"force": "F", "torque": "T"}.get(self.storage.get(getattr(tgt, "id", None)))
                if out is None:
                    raise KernelGenError("apply(): a contact model applies to the force and the torque property")
                v = self.expr(val)
                if not self.is_vec(v):
                    raise KernelGenError("apply(): vector property needs a vector expression")
                first = out not in self.applied and self.depth == 0
                self.applied[out] = True
                for d, c in enumerate(v[1]):
                    self.lines.append(f"{out}[{d}] = {c};" if first else f"{out}[{d}] = {out}[{d}] + {c};")
                return
            if self.kind != "pair":
                raise KernelGenError("apply() needs a pair kernel")
            tgt, val = node.value.args
            store = self.storage.get(getattr(tgt, "id", None))
            if (store not in ("force", "vel") and not isinstance(store, tuple)) or (isinstance(store, tuple) and len(store) > 3):
                raise KernelGenError("apply(): the target must be a declared vector property")
            v = self.expr(val)
            if isinstance(store, tuple) and store[2] == 1:
                # a scalar target: the reference accepts it but prints code that does not compile (its reduction variable is
                # always a 3-vector, ir/apply.py:41-42); here it accumulates the scalar, the evident meaning
                if self.is_vec(v):
                    raise KernelGenError("apply(): scalar property needs a scalar expression")
                comps = [v[1]]
            else:
                if not self.is_vec(v):
                    raise KernelGenError("apply(): vector property needs a vector expression")
                comps = v[1]
            acc = self.applied.setdefault(store, [f"acc_{_store_tag(store)}_{d}" for d in range(len(comps))])
            for a, c in zip(acc, comps):
                self.lines.append(f"{a} = {a} + {c};")
            if getattr(self, "half", False):              # Newton's third law: the partner gets the opposite term (ir/apply.py:111-125)
                self.lines.append("if(j < a.nlocal && (a.flags[j] & PB_FLAG_FIXED) == 0) {")
                for d, c in enumerate(comps):
                    ref = _xref(store, d, "j") if isinstance(store, tuple) else (f"a.{store}[{d} * (size_t) a.cap + j]" if d else f"a.{store}[j]")
                    self.lines.append(f"    atomicAdd(&{ref}, -({c}));")
                self.lines.append("}")
            return
        if self.kind == "dem" and isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Subscript) \
                and isinstance(node.targets[0].value, ast.Name) and node.targets[0].value.id in self.contact:
            tgt = node.targets[0]
            if not isinstance(tgt.slice, ast.Tuple) or [getattr(e, "id", None) for e in tgt.slice.elts] != ["i", "j"]:
                raise KernelGenError("contact properties are assigned as prop[i, j] = ...")
            kind, v = self.contact[tgt.value.id], self.expr(node.value)
            if kind.startswith("cx:"):
                _, ctype, off = kind.split(":")
                if (ctype == "vec") != bool(self.is_vec(v)):
                    raise KernelGenError(f"'{tgt.value.id}' is a {'vector' if ctype == 'vec' else 'scalar'} contact property")
                if ctype == "vec":
                    for d, c in enumerate(v[1]):
                        self.lines.append(f"cx[{int(off) + d}] = {c};")
                elif ctype == "real":
                    self.lines.append(f"cx[{off}] = {v[1]};")
                else:
                    self.lines.append(f"cx[{off}] = (double) (int) ({v[1]});")
            elif kind == "c_tsd":
                if not self.is_vec(v):
                    raise KernelGenError(f"'{tgt.value.id}' is a vector contact property")
                for d, c in enumerate(v[1]):
                    self.lines.append(f"tsd[{d}] = {c};")
            elif self.is_vec(v):
                raise KernelGenError(f"'{tgt.value.id}' is a scalar contact property")
            elif kind == "c_ivm":
                self.lines.append(f"*ivm = {v[1]};")
            else:
                self.lines.append(f"*sticking = (int) ({v[1]});")
            return
        if self.kind == "pair" and isinstance(node, ast.AugAssign) and isinstance(node.op, (ast.Add, ast.Sub)) \
                and isinstance(node.target, ast.Subscript) and isinstance(node.target.value, ast.Name) \
                and getattr(node.target.slice, "id", None) == "i":
            # the older API of examples/lj_onetype.py writes `force[i] += expr` inside the pair kernel: the same accumulation as
            # apply(force, expr)
            value = node.value if isinstance(node.op, ast.Add) else ast.UnaryOp(op=ast.USub(), operand=node.value)
            call = ast.Call(func=ast.Name(id="apply", ctx=ast.Load()), args=[ast.Name(id=node.target.value.id, ctx=ast.Load()), value], keywords=[])
            return self.stmt(ast.Expr(value=call))
        if self.kind == "particle" and isinstance(node, (ast.Assign, ast.AugAssign)):
            tgt = node.targets[0] if isinstance(node, ast.Assign) else node.target
            if isinstance(tgt, ast.Subscript) and isinstance(tgt.value, ast.Subscript) and isinstance(tgt.slice, ast.Constant) \
                    and isinstance(tgt.value.value, ast.Name) and getattr(tgt.value.slice, "id", None) == "i":
                # prop[i][k] = expr: one component of a vector property (examples/dem.py:90 force[i][2] = ...)
                store, k = self.storage.get(tgt.value.value.id), tgt.slice.value
                cur = self.load(store, "i") if store is not None else None
                if cur is None or not self.is_vec(cur) or not isinstance(k, int) or not 0 <= k < 3:
                    raise KernelGenError(f"'{ast.unparse(tgt)}': component assignment needs a declared vector property and an index 0..2")
                v = self.expr(node.value)
                if isinstance(node, ast.AugAssign):
                    ops = {ast.Add: "+", ast.Sub: "-", ast.Mult: "*", ast.Div: "/"}
                    v = self.binop(ops[type(node.op)], ("f", cur[1][k]), v)
                if self.is_vec(v):
                    raise KernelGenError("a component takes a scalar")
                comps = list(cur[1])
                comps[k] = v[1]
                self.store(store, self.vec(comps), only=k)
                return
            if isinstance(tgt, ast.Subscript) and isinstance(tgt.value, ast.Name) and getattr(tgt.slice, "id", None) == "i":
                store = self.storage.get(tgt.value.id)
                if store is None:
                    raise KernelGenError(f"'{tgt.value.id}' is not a declared real / vector property")
                v = self.expr(node.value)
                if isinstance(node, ast.AugAssign):
                    ops = {ast.Add: "+", ast.Sub: "-", ast.Mult: "*", ast.Div: "/"}
                    v = self.binop(ops[type(node.op)], self.load(store, "i"), v)
                self.store(store, v)
                return
        raise KernelGenError(f"unsupported statement: {ast.unparse(node)}")

    def store(self, store, v, only=None):
        """prop[i] = v; `only` = k writes component k alone (the value's other components are what was loaded)."""
        if store in ("uid", "shape", "flags", "type"):
            raise KernelGenError("integer properties (uid, shape, flags, the feature) are read-only in kernels")
        self.stored.add(store)
        if store in ("inv_inertia", "rotmat", "quat"):
            n = 4 if store == "quat" else 9
            if not isinstance(v[1], list):                    # a scalar assigned to a matrix: every element (examples/dem.py:15)
                v = ("q" if store == "quat" else "m", [v[1]] * n)
            if len(v[1]) != n or self.is_vec(v):
                raise KernelGenError(f"'{store}' takes a {'quaternion' if n == 4 else 'matrix'}")
            for d in range(n):
                self.lines.append(f"a.{store}[{d} * (size_t) a.cap + i] = {v[1][d]};")
            self.loaded[(store, "i")] = v
            return
        if store in ("mass", "radius"):
            if isinstance(v[1], list):
                raise KernelGenError(f"{store} is a scalar")
            self.lines.append(f"a_mass_w[i] = {v[1]};" if store == "mass" else f"a.radius[i] = {v[1]};")
            self.loaded[(store, "i")] = v
            return
        if isinstance(store, tuple):
            comps = [v[1]] if not self.is_vec(v) else v[1]
            if len(comps) != store[2]:
                raise KernelGenError("assignment: a real property takes a scalar, a vector property a vector")
            if len(store) > 3:                                # integer property: C conversion on assignment, as the reference's int array
                iv = v if v[0] == "i" else ("i", self.tmp("int", f"(int) ({v[1]})"))
                self.lines.append(f"{_xref(store, 0, 'i')} = (double) {iv[1]};")
                self.loaded[(store, "i")] = iv
                return
            for d, c in enumerate(comps):
                if only is None or only == d:
                    self.lines.append(f"{_xref(store, d, 'i')} = {c};")
            self.loaded[(store, "i")] = v
            return
        if not self.is_vec(v):
            raise KernelGenError(f"'{store}' is a vector property")
        if store == "pos":
            self.lines.append(f"pi.x = {v[1][0]}; pi.y = {v[1][1]}; pi.z = {v[1][2]}; a.pos_w[i] = pi;")
            self.loaded[(store, "i")] = v                      # later reads see the value just assigned (const temporaries)
        else:
            for d in range(3):
                if only is None or only == d:
                    self.lines.append(f"a.{store}[{d} * (size_t) a.cap + i] = {v[1][d]};")
            self.loaded[(store, "i")] = v


def _function_ast(func):
    """The kernel function's AST, as the reference obtains it (inspect.getsource, mapping/funcs.py:286-288)."""
    try:
        src = textwrap.dedent(inspect.getsource(func))
    except (OSError, TypeError):
        raise KernelGenError(f"the source text of '{getattr(func, '__name__', func)}' is not available (kernels are translated from their "
                             "source: define them in a file)") from None
    return ast.parse(src).body[0]


def translate(func, storage, feature_tables, ntypes, symbols, prelude, skip_fixed=True, traversal="lists", half=False):
    """-> (kind, kernel name, CUDA source).  `storage` maps the user's property names to device arrays.  skip_fixed=False is
    for setup() functions: the reference runs those over every local particle (no FIXED filter, mapping/funcs.py:305-310 applies
    to compute() only).  traversal: pair kernels walk the neighbour lists ("lists") or, for scripts that build cell lists only,
    cell 0 and the 27 stencil cells of the particle's cell ("cells": sim/interaction.py:92-118, the three z-adjacent cells of a
    stencil row being one run of the CSR).  half=True is Simulation.compute_half(): the lists hold every pair once and apply()
    also subtracts the term from the partner with an atomic add unless it is a ghost or FIXED (ir/apply.py:111-125; the own
    particle's sum is added atomically too, other threads may be updating it)."""
    tree = _function_ast(func)
    if not isinstance(tree, ast.FunctionDef):
        raise KernelGenError(f"{func.__name__}: not a plain function")
    params = [a.arg for a in tree.args.args]
    if params == ["i", "j"]:
        kind = "pair"
    elif params == ["i"]:
        kind = "particle"
    else:
        raise KernelGenError(f"{func.__name__}: kernels take (i) or (i, j)")
    name = f"user_{func.__name__}"
    g = _Gen(name, kind, storage, feature_tables, ntypes, symbols, func.__globals__)
    g.half = bool(half) and kind == "pair"
    if half and traversal != "lists":
        raise KernelGenError("compute_half() needs neighbour lists")
    for node in tree.body:
        g.stmt(node)
    out = [prelude]
    for fp, table in feature_tables.items():
        out.append(f"__device__ const double fp_{fp}[{len(table)}] = {{{', '.join(_lit(float(x)) for x in table)}}};")
    out.append(f'extern "C" __global__ void __launch_bounds__(128) {name}(PbJitArgs a) {{')
    out.append("    const int i = blockIdx.x * blockDim.x + threadIdx.x;")
    out.append("    if(i >= a.nlocal || (a.flags[i] & PB_FLAG_FIXED) != 0) { return; }" if skip_fixed else "    if(i >= a.nlocal) { return; }")
    if kind == "pair":
        out.append("    const double4 pi = pb_ld_pos(a.pos + i);")
        if g.types_needed:
            out.append(f"    const int ti = pb_w_type(pi.w) * {ntypes};")
        out += ["    " + ln for ln in g.hoisted]
        for store, acc in g.applied.items():
            out += [f"    double {x} = 0.0;" for x in acc]
        if traversal == "cells":
            out.append("    const int pc = a.particle_cell[i];")
            out.append("    for(int run = 0; run < 10; run++) {")
            out.append("    int c_lo = 0, c_hi = 0;")                     # run 0: cell 0 (disp = -1 of the reference's loop)
            out.append("    if(run > 0) {")
            out.append("        const int mid = pc + (((run - 1) / 3 - 1) * a.dim1 + ((run - 1) % 3 - 1)) * a.dim2;")
            out.append("        c_lo = (mid - 1 > 1) ? mid - 1 : 1;")      # 0 < cell < ncells
            out.append("        c_hi = (mid + 1 < a.ncells - 1) ? mid + 1 : a.ncells - 1;")
            out.append("        if(c_lo > c_hi) { continue; }")
            out.append("    }")
            out.append("    const int k_end = a.cell_start[c_hi + 1];")
            out.append("    for(int k = a.cell_start[c_lo]; k < k_end; k++) {")
            out.append("        const int j = __ldg(a.cell_list + k);")
            out.append("        if(j == i) { continue; }")
        elif PAIR_PREFETCH > 1:
            U = PAIR_PREFETCH
            out.append("    const int nn = a.numneigh[i];")
            out.append("    const int *nb = a.neigh + (size_t) (i >> 5) * a.nslots * 32 + (i & 31);")
            out.append(f"    for(int k0 = 0; k0 < nn; k0 += {U}) {{")
            out.append(f"    int jj[{U}];")
            out.append(f"    double4 pp[{U}];")
            out.append("#pragma unroll")
            out.append(f"    for(int u = 0; u < {U}; u++) {{ jj[u] = (k0 + u < nn) ? __ldg(nb + (size_t) (k0 + u) * 32) : i; }}")
            out.append("#pragma unroll")
            out.append(f"    for(int u = 0; u < {U}; u++) {{ pp[u] = pb_ld_pos(a.pos + jj[u]); }}")
            out.append("#pragma unroll")
            out.append(f"    for(int u = 0; u < {U}; u++) {{")
            out.append("        if(k0 + u >= nn) { break; }")
            out.append("        const int j = jj[u];")
        else:
            out.append("    const int nn = a.numneigh[i];")
            out.append("    const int *nb = a.neigh + (size_t) (i >> 5) * a.nslots * 32 + (i & 31);")
            out.append("    for(int k = 0; k < nn; k++) {")
            out.append("        const int j = __ldg(nb + (size_t) k * 32);")
        out.append("        const double4 pj = pp[u];" if (traversal == "lists" and PAIR_PREFETCH > 1) else "        const double4 pj = pb_ld_pos(a.pos + j);")
        out.append("        const double dx = pi.x - pj.x;")          # delta(i, j) = position[i] - position[j]
        out.append("        const double dy = pi.y - pj.y;")
        out.append("        const double dz = pi.z - pj.z;")
        out.append("        const double rsq_a = dx * dx;")           # (dx*dx + dy*dy) + dz*dz, one operation per statement
        out.append("        const double rsq_b = dy * dy;")
        out.append("        const double rsq_c = rsq_a + rsq_b;")
        out.append("        const double rsq_d = dz * dz;")
        out.append("        const double rsq = rsq_c + rsq_d;")
        out.append("        if(rsq < a.cutsq) {")
        if g.types_needed:
            out.append("            const int tj = pb_w_type(pj.w);")
        out += ["            " + ln for ln in g.lines]
        out.append("        }")
        out.append("    }")
        if traversal == "cells" or PAIR_PREFETCH > 1:
            out.append("    }")
        for store, acc in g.applied.items():                          # prop[i] = prop[i] + acc (sim/interaction.py:280-292)
            for d, x in enumerate(acc):
                if isinstance(store, tuple):
                    ref = _xref(store, d, "i")
                else:
                    ref = f"a.{store}[{d} * (size_t) a.cap + i]" if d else f"a.{store}[i]"
                out.append(f"    atomicAdd(&{ref}, {x});" if g.half else f"    {ref} = {ref} + {x};")
    else:
        out.append("    double4 pi = a.pos_w[i];")
        out.append("    double *a_mass_w = const_cast<double *>(a.mass);")
        out += ["    " + ln for ln in g.lines]
    out.append("}")
    return kind, name, "\n".join(out) + "\n"


def translate_dem_model(func, storage, contact, feature_tables, symbols, contact_defaults=None, extra_lanes=0):
    """A DEM contact model (the body of a pair kernel over contact history, e.g. examples/dem.py:18-74) -> (function name, CUDA
    source) of the device function the library's contact kernel calls for every touching pair (csrc/dem_force_kernel.cuh):

        bool f(xi, vi, wi, mi, ri, xj, vj, wj, mj, rj, n, cp, delta, tij, tsd, ivm, sticking, cx, F, T)

    xi.. = position, linear velocity, angular velocity, mass, radius of i and j; n / cp / delta = contact normal, contact point and
    -penetration_depth from the kernel's geometry pass; tij = type[i] * ntypes + type[j] (feature properties are literal tables);
    tsd / ivm / sticking = this pair's contact properties (in / out); F / T = what the body apply()s to force / torque.
    Returns false when a skip_when() left the pair.  `storage`: property name -> 'pos' | 'vel' | 'angvel' | 'mass' | 'radius' |
    'force' | 'torque'; `contact`: contact property name -> 'c_tsd' | 'c_ivm' | 'c_stick' (the three columns examples/dem.py
    declares) or 'cx:real:<lane>' | 'cx:vec:<lane>' | 'cx:int:<lane>' for further contact properties, which live in the `extra_lanes`
    double lanes cx[] of the contact (pb_dem_enable_ex; an integer is held as an exact double); `contact_defaults`: kind -> the
    default of add_contact_property() a fresh contact slot starts from (zeros if absent; the extra lanes' defaults go to
    pb_dem_enable_ex)."""
    tree = _function_ast(func)
    if not isinstance(tree, ast.FunctionDef) or [a.arg for a in tree.args.args] != ["i", "j"]:
        raise KernelGenError(f"{func.__name__}: a contact model takes (i, j)")
    name = f"user_model_{func.__name__}"
    g = _Gen(name, "dem", storage, feature_tables, 0, symbols, func.__globals__)
    g.contact = dict(contact)
    for node in tree.body:
        g.stmt(node)
    out = []
    if extra_lanes > 0:
        out.append(f"#define PB_DEM_NX {int(extra_lanes)}")
    for kind, value in (contact_defaults or {}).items():
        if kind == "c_stick":
            out.append(f"#define PB_DEM_DEFAULT_STICK {int(value)}")
        elif kind == "c_ivm":
            out.append(f"#define PB_DEM_DEFAULT_IVM {_lit(float(value))}")
        elif kind == "c_tsd":
            v = list(value) if isinstance(value, (list, tuple)) else [value] * 3
            out.append(f"#define PB_DEM_DEFAULT_TSD(d) ((d) == 0 ? {_lit(float(v[0]))} : ((d) == 1 ? {_lit(float(v[1]))} : {_lit(float(v[2]))}))")
    for fp, table in feature_tables.items():
        out.append(f"__device__ const double fp_{fp}[{len(table)}] = {{{', '.join(_lit(float(x)) for x in table)}}};")
    out.append(f"__device__ __forceinline__ bool {name}(const double *xi, const double *vi, const double *wi, double mi, double ri, "
               "const double *xj, const double *vj, const double *wj, double mj, double rj, const double *n, const double *cp, "
               "double delta, int tij, double *tsd, double *ivm, int *sticking, double *cx, double *F, double *T) {")
    out.append("    F[0] = 0.0; F[1] = 0.0; F[2] = 0.0; T[0] = 0.0; T[1] = 0.0; T[2] = 0.0;")
    out += ["    " + ln for ln in g.lines]
    out.append("    return true;")
    out.append("}")
    return name, "\n".join(out) + "\n"
