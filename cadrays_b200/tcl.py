"""TCL-subset scene loader: runs the path-tracing scripts CADRays ships
(data/scripts/CornellBox.tcl, data/scripts/Materials.tcl) and the model.tcl files its exporter
writes (src/ImportExport/ImportExport.cxx:155-231,444-606) without OCCT, Tcl or DRAW.

It is a small Tcl evaluator (set / expr / for / if / incr / eval / lrepeat / puts, $var and [cmd]
substitution, braces and quotes) plus the DRAW / ViewerTest commands those scripts use:

  shapes   box psphere pcylinder compound explode ttranslate trotate tcopy
  viewer   vclear vdisplay verase vremove vlocation vsetmaterial vbsdf vlight rtlight vcamera vviewparams
           vfront vback vtop vbottom vleft vright vaxo vfit vrenderparams vtextureenv vsetdispmode
           vinit vfps vdump(ignored)
  CADRays  rtmeshread (PLY, the exporter's mesh format) rtdisplay rterase

Everything produces a cadrays_b200.scenes.SceneDesc; rendering stays in libcadrays_b200.so.
Default BSDFs of OCCT's named materials (`vsetmaterial <obj> glass`) live in the absent
Graphic3d_MaterialAspect.cxx; NAMED_MATERIALS below is a re-derivation (parity unpinned).
"""
from __future__ import annotations

import ast
import math
import operator
import re
import os
from typing import Callable, Dict, List, Optional

import numpy as np

from . import scenes
from .view import (Graphic3d_BSDF, Graphic3d_Camera, Graphic3d_Fresnel, Graphic3d_RenderingParams,
                   Graphic3d_ToneMappingMethod_Filmic, make_light)


class TclError(RuntimeError):
    pass


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class _Return(Exception):
    def __init__(self, value=""):
        super().__init__(value)
        self.value = value


# ------------------------------------------------------------------ named materials (re-derived)

def _metal(f0, rough):
    return lambda: Graphic3d_BSDF(Ks=[0.985, 0.985, 0.985, rough], FresnelBase=Graphic3d_Fresnel.CreateSchlick(*f0))


def _diffuse(kd):
    return lambda: Graphic3d_BSDF(Kd=list(kd))


def _plastic(kd, rough=0.1):
    return lambda: Graphic3d_BSDF(Kd=list(kd), Ks=[0.04, 0.04, 0.04, rough], FresnelBase=Graphic3d_Fresnel.CreateConstant(1.0))


# Values: what the reference itself reveals, in this order of trust --
#   (1) BSDFs the shipped scripts export right after `vsetmaterial X` (Materials.tcl:40-190: Brass Schlick colour,
#       Glass absorption 0.75 0.95 0.9 / 0.05 with a 1.62 dielectric, Aluminium Schlick colour and roughness);
#   (2) the 64x64 icons data/materials/<MaterialName>.png, which are OCCT path-traced renders of
#       data/other/preview.tcl for every named material: the ball's colour relative to the plaster ball calibrates
#       albedo and tint here (tests/golden/material_icons.json, tests/test_tcl_cpu.py);
#   (3) recalled Graphic3d_MaterialAspect.cxx values where neither exists.
# The icons also settle which name is which: files are named Graphic3d_MaterialAspect::MaterialName(id)
# (main.cxx:120-132), in enum order "... Metalized, Ionized, Chrome, Aluminium, Obsidian, Neon, Jade ...", so
# NEON_GNC is the grey "Ionized" and NEON_PHC the green emissive "Neon".
NAMED_MATERIALS: Dict[str, Callable[[], Graphic3d_BSDF]] = {
    "brass": _metal((0.58, 0.42, 0.20), 0.045),            # Schlick colour as used with Brass in Materials.tcl:49
    "bronze": _metal((0.65, 0.35, 0.15), 0.045),
    "copper": _metal((0.955, 0.638, 0.538), 0.045),
    "gold": _metal((1.0, 0.782, 0.344), 0.045),
    "pewter": _metal((0.55, 0.57, 0.60), 0.2),
    "silver": _metal((0.972, 0.960, 0.915), 0.045),
    "steel": _metal((0.56, 0.57, 0.58), 0.1),
    "chrome": _metal((0.549, 0.556, 0.554), 0.02),
    "aluminium": _metal((0.913183, 0.921494, 0.924524), 0.026),   # Materials.tcl:159-171
    "aluminum": _metal((0.913183, 0.921494, 0.924524), 0.026),
    "metalized": _metal((0.28, 0.28, 0.26), 0.2),
    "plaster": _diffuse((0.482353, 0.482353, 0.482353)),
    "plastic": _plastic((0.2, 0.2, 0.2)),
    "shiny_plastic": _plastic((0.3, 0.3, 0.3), 0.02),
    "satin": _plastic((0.62, 0.62, 0.62), 0.3),
    "stone": _diffuse((0.25, 0.245, 0.24)),
    "charcoal": _plastic((0.08, 0.08, 0.08), 0.3),
    "obsidian": _plastic((0.03, 0.01, 0.027), 0.02),
    "jade": _plastic((0.23, 0.43, 0.23), 0.1),
    "neon_gnc": _plastic((0.28, 0.28, 0.28), 0.1),                                   # "Ionized"
    "neon_phc": lambda: Graphic3d_BSDF(Kd=[0.1, 0.1, 0.1], Le=[0.0, 1.0, 0.46]),     # "Neon"
    "glass": lambda: Graphic3d_BSDF.CreateGlass((1, 1, 1), (0.75, 0.95, 0.9), 0.05, 1.62),   # Materials.tcl:74-88
    "water": lambda: Graphic3d_BSDF.CreateGlass((1, 1, 1), (0.8, 0.9, 1.0), 0.05, 1.33),
    "diamond": lambda: Graphic3d_BSDF.CreateGlass((1, 1, 1), (1, 1, 1), 0.0, 2.42),
    "transparent": lambda: Graphic3d_BSDF.CreateGlass((1, 1, 1), (1, 1, 1), 0.0, 1.0),
    "default": _plastic((0.6, 0.6, 0.6)),
}
# Graphic3d_MaterialAspect::MaterialName() spellings (the GUI's and the icon files' names)
for _alias, _name in (("plastered", "plaster"), ("plastified", "plastic"), ("shiny_plastified", "shiny_plastic"),
                      ("satined", "satin"), ("ionized", "neon_gnc"), ("neon", "neon_phc")):
    NAMED_MATERIALS[_alias] = NAMED_MATERIALS[_name]


# ------------------------------------------------------------------ a small Tcl evaluator

_BIN = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: None, ast.Mod: operator.mod,
        ast.Pow: operator.pow, ast.BitAnd: operator.and_, ast.BitOr: operator.or_, ast.BitXor: operator.xor,
        ast.LShift: operator.lshift, ast.RShift: operator.rshift}
_CMP = {ast.Eq: operator.eq, ast.NotEq: operator.ne, ast.Lt: operator.lt, ast.LtE: operator.le, ast.Gt: operator.gt, ast.GtE: operator.ge}
_FUN = {"sin": math.sin, "cos": math.cos, "tan": math.tan, "sqrt": math.sqrt, "abs": abs, "int": int, "double": float,
        "round": round, "floor": math.floor, "ceil": math.ceil, "pow": math.pow, "atan": math.atan, "atan2": math.atan2,
        "asin": math.asin, "acos": math.acos, "exp": math.exp, "log": math.log, "min": min, "max": max, "fmod": math.fmod}


def tcl_expr(text: str):
    """Tcl `expr` for the arithmetic the scripts use (C operators, integer division, math functions)."""
    src = text.replace("&&", " and ").replace("||", " or ").replace("!", " not ").replace(" not =", "!=")
    src = re.sub(r"\beq\b", "==", re.sub(r"\bne\b", "!=", src))          # string comparison operators
    try:
        tree = ast.parse(src.strip(), mode="eval")
    except SyntaxError:
        raise TclError(f"expr: syntax error in '{text}'")

    def same_kind(a, b):
        # Tcl compares numerically when both operands are numbers, else as strings
        if isinstance(a, str) or isinstance(b, str):
            return str(a) if not isinstance(a, str) else a, str(b) if not isinstance(b, str) else b
        return a, b

    def ev(n):
        if isinstance(n, ast.Expression):
            return ev(n.body)
        if isinstance(n, ast.Constant) and isinstance(n.value, (int, float)):
            return n.value
        if isinstance(n, ast.Constant) and isinstance(n.value, str):
            return n.value
        if isinstance(n, ast.Name):              # a substituted variable that held a word: a string operand
            return n.id
        if isinstance(n, ast.BinOp):
            a, b = ev(n.left), ev(n.right)
            if isinstance(n.op, ast.Div):
                if isinstance(a, int) and isinstance(b, int):
                    return a // b          # Tcl integer division
                return a / b
            return _BIN[type(n.op)](a, b)
        if isinstance(n, ast.UnaryOp):
            v = ev(n.operand)
            if isinstance(n.op, ast.USub):
                return -v
            if isinstance(n.op, ast.UAdd):
                return +v
            if isinstance(n.op, ast.Not):
                return int(not v)
        if isinstance(n, ast.BoolOp):
            vals = [ev(v) for v in n.values]
            return int(all(vals)) if isinstance(n.op, ast.And) else int(any(vals))
        if isinstance(n, ast.Compare) and len(n.ops) == 1:
            return int(_CMP[type(n.ops[0])](*same_kind(ev(n.left), ev(n.comparators[0]))))
        if isinstance(n, ast.IfExp):
            return ev(n.body) if ev(n.test) else ev(n.orelse)
        if isinstance(n, ast.Call) and isinstance(n.func, ast.Name) and n.func.id in _FUN:
            return _FUN[n.func.id](*[ev(a) for a in n.args])
        raise TclError(f"expr: unsupported expression '{text}'")
    return ev(tree)


def _fmt(v) -> str:
    if isinstance(v, bool):
        return "1" if v else "0"
    if isinstance(v, float):
        return repr(v) if not v.is_integer() or abs(v) > 1e15 else f"{v:.1f}"
    return str(v)


class _Linked(dict):
    """A proc's variable table in which some names are links to the global table (`global name`)."""

    def __init__(self, local, globals_, name):
        super().__init__(local)
        self._g = globals_
        self._names = set(getattr(local, "_names", ())) | {name}
        for nm in self._names:
            super().pop(nm, None)

    def __contains__(self, k):
        return (k in self._names and k in self._g) or super().__contains__(k)

    def __getitem__(self, k):
        return self._g[k] if k in self._names else super().__getitem__(k)

    def __setitem__(self, k, v):
        if k in self._names:
            self._g[k] = v
        else:
            super().__setitem__(k, v)

    def get(self, k, default=None):
        return self._g.get(k, default) if k in self._names else super().get(k, default)

    def pop(self, k, *default):
        return self._g.pop(k, *default) if k in self._names else super().pop(k, *default)


class Interp:
    """Draw_Interpretor stand-in: command table + variables."""

    def __init__(self):
        self.vars: Dict[str, str] = {}
        self.cmds: Dict[str, Callable[[List[str]], Optional[str]]] = {}
        self.unknown: List[str] = []
        self.strict = False
        self.output: List[str] = []
        for name in ("set", "expr", "for", "if", "incr", "eval", "lrepeat", "puts", "list", "llength", "lindex", "foreach",
                     "while", "catch", "unset", "append", "string", "proc", "return", "break", "continue", "global",
                     "lappend", "lrange", "concat", "join", "split", "format", "info"):
            self.cmds[name] = getattr(self, "_c_" + name)
        self._frames: List[Dict[str, str]] = []      # saved variable tables of the callers while a proc body runs
        self._globals: Dict[str, str] = self.vars

    # -- parsing
    def eval(self, script: str) -> str:
        result = ""
        for words in self._commands(script):
            if not words:
                continue
            result = self._dispatch(words)
        return result

    def _dispatch(self, words: List[str]) -> str:
        name = words[0]
        fn = self.cmds.get(name)
        if fn is None:
            if self.strict:
                raise TclError(f'invalid command name "{name}"')
            self.unknown.append(name)
            return ""
        r = fn(words[1:])
        return "" if r is None else str(r)

    def _commands(self, script: str):
        i, n = 0, len(script)
        while i < n:
            # skip whitespace / separators
            while i < n and script[i] in " \t\r\n;":
                i += 1
            if i >= n:
                break
            if script[i] == "#":
                while i < n and script[i] != "\n":
                    if script[i] == "\\" and i + 1 < n:
                        i += 1
                    i += 1
                continue
            words: List[str] = []
            while i < n and script[i] not in "\n;":
                while i < n and script[i] in " \t\r":
                    i += 1
                if i >= n or script[i] in "\n;":
                    break
                if script[i] == "\\" and i + 1 < n and script[i + 1] == "\n":
                    i += 2
                    continue
                w, i = self._word(script, i)
                words.append(w)
            yield words

    def _word(self, s: str, i: int):
        n = len(s)
        if s[i] == "{":
            depth, j = 1, i + 1
            while j < n and depth:
                if s[j] == "\\":
                    j += 1
                elif s[j] == "{":
                    depth += 1
                elif s[j] == "}":
                    depth -= 1
                j += 1
            if depth:
                raise TclError("missing close-brace")
            return s[i + 1:j - 1], j
        out = []
        quoted = s[i] == '"'
        if quoted:
            i += 1
        while i < n:
            c = s[i]
            if quoted and c == '"':
                i += 1
                break
            if not quoted and c in " \t\r\n;":
                break
            if c == "\\" and i + 1 < n:
                nxt = s[i + 1]
                out.append({"n": "\n", "t": "\t", "\n": " "}.get(nxt, nxt))
                i += 2
            elif c == "$":
                j = i + 1
                if j < n and s[j] == "{":
                    k = s.index("}", j)
                    name, j = s[j + 1:k], k + 1
                else:
                    if s.startswith("::", j):          # $::name: the global namespace, the only one here
                        j += 2
                    while j < n and (s[j].isalnum() or s[j] == "_"):
                        j += 1
                    name = s[i + 1:j]
                if name.startswith("::"):
                    name = name[2:]
                if not name:
                    out.append("$")
                    i += 1
                    continue
                if name not in self.vars:
                    raise TclError(f'can\'t read "{name}": no such variable')
                out.append(self.vars[name])
                i = j
            elif c == "[":
                depth, j = 1, i + 1
                while j < n and depth:
                    if s[j] == "\\":
                        j += 1
                    elif s[j] == "[":
                        depth += 1
                    elif s[j] == "]":
                        depth -= 1
                    j += 1
                if depth:
                    raise TclError("missing close-bracket")
                out.append(self.eval(s[i + 1:j - 1]))
                i = j
            else:
                out.append(c)
                i += 1
        return "".join(out), i

    def subst(self, text: str) -> str:
        """$var and [cmd] substitution inside a braced expression (used by expr / if / for)."""
        w, _ = self._word('"' + text.replace('"', '\\"') + '"', 0)
        return w

    # -- core commands
    def _c_set(self, a):
        if len(a) == 1:
            return self.vars[a[0]]
        self.vars[a[0]] = a[1]
        return a[1]

    def _c_unset(self, a):
        for v in a:
            self.vars.pop(v, None)

    def _c_append(self, a):
        self.vars[a[0]] = self.vars.get(a[0], "") + "".join(a[1:])
        return self.vars[a[0]]

    def _c_expr(self, a):
        return _fmt(tcl_expr(self.subst(" ".join(a))))

    def _cond(self, text):
        return bool(tcl_expr(self.subst(text)))

    def _c_for(self, a):
        init, cond, step, body = a
        self.eval(init)
        guard = 0
        while self._cond(cond):
            try:
                self.eval(body)
            except _Break:
                break
            except _Continue:
                pass
            self.eval(step)
            guard += 1
            if guard > 10_000_000:
                raise TclError("for: runaway loop")

    def _c_while(self, a):
        guard = 0
        while self._cond(a[0]):
            try:
                self.eval(a[1])
            except _Break:
                break
            except _Continue:
                pass
            guard += 1
            if guard > 10_000_000:
                raise TclError("while: runaway loop")

    def _c_foreach(self, a):
        var, items, body = a
        names = self._split_list(var)
        values = self._split_list(items)
        for k in range(0, len(values), max(1, len(names))):
            for j, nm in enumerate(names):
                self.vars[nm] = values[k + j] if k + j < len(values) else ""
            try:
                self.eval(body)
            except _Break:
                break
            except _Continue:
                continue

    def _c_if(self, a):
        i = 0
        while i < len(a):
            if a[i] in ("else",):
                return self.eval(a[i + 1])
            if a[i] == "elseif":
                i += 1
            cond = a[i]
            body_i = i + 2 if i + 1 < len(a) and a[i + 1] == "then" else i + 1
            if self._cond(cond):
                return self.eval(a[body_i])
            i = body_i + 1
        return ""

    def _c_incr(self, a):
        v = int(self.vars.get(a[0], "0")) + (int(a[1]) if len(a) > 1 else 1)
        self.vars[a[0]] = str(v)
        return str(v)

    def _c_eval(self, a):
        return self.eval(" ".join(a))

    # -- procedures: one variable table per call, `global` links names to the outermost table
    def _c_proc(self, a):
        name, params, body = a
        spec = [self._split_list(p) for p in self._split_list(params)]

        def call(args, _spec=spec, _body=body, _name=name):
            local: Dict[str, str] = {}
            rest = list(args)
            for k, p in enumerate(_spec):
                if p[0] == "args" and k == len(_spec) - 1:
                    local["args"] = self._join_list(rest)
                    rest = []
                elif rest:
                    local[p[0]] = rest.pop(0)
                elif len(p) > 1:
                    local[p[0]] = p[1]
                else:
                    raise TclError(f'wrong # args: should be "{_name} {params}"')
            if rest:
                raise TclError(f'wrong # args: should be "{_name} {params}"')
            self._frames.append(self.vars)
            self.vars = local
            try:
                return self.eval(_body)
            except _Return as r:
                return r.value
            finally:
                self.vars = self._frames.pop()

        self.cmds[name] = call
        return ""

    def _c_return(self, a):
        raise _Return(a[-1] if a else "")

    def _c_break(self, a):
        raise _Break()

    def _c_continue(self, a):
        raise _Continue()

    def _c_global(self, a):
        if self.vars is self._globals:
            return ""
        for nm in a:
            self.vars = _Linked(self.vars, self._globals, nm)
        return ""

    def _c_lappend(self, a):
        items = self._split_list(self.vars.get(a[0], "")) + list(a[1:])
        self.vars[a[0]] = self._join_list(items)
        return self.vars[a[0]]

    def _c_lrange(self, a):
        items = self._split_list(a[0])
        idx = lambda t: len(items) - 1 + (int(t[3:]) if t[3:4] in ("-", "+") else 0) if t.startswith("end") else int(t)
        return self._join_list(items[max(0, idx(a[1])):idx(a[2]) + 1])

    def _c_concat(self, a):
        return " ".join(x.strip() for x in a if x.strip())

    def _c_join(self, a):
        return (a[1] if len(a) > 1 else " ").join(self._split_list(a[0]))

    def _c_split(self, a):
        seps = a[1] if len(a) > 1 else " \t\n"
        out, cur = [], ""
        for ch in a[0]:
            if ch in seps:
                out.append(cur); cur = ""
            else:
                cur += ch
        out.append(cur)
        return self._join_list(out)

    def _c_format(self, a):
        vals = []
        for v in a[1:]:
            try:
                vals.append(int(v))
            except ValueError:
                try:
                    vals.append(float(v))
                except ValueError:
                    vals.append(v)
        try:
            return a[0] % tuple(vals)
        except (TypeError, ValueError) as e:
            raise TclError(f"format: {e}")

    def _c_info(self, a):
        if a and a[0] == "exists":
            return "1" if a[1] in self.vars else "0"
        if a and a[0] in ("commands", "procs"):
            return self._join_list(sorted(self.cmds))
        raise TclError("info: unsupported subcommand " + (a[0] if a else ""))

    def _c_catch(self, a):
        try:
            r = self.eval(a[0])
            if len(a) > 1:
                self.vars[a[1]] = r
            return "0"
        except (_Break, _Continue, _Return):
            raise
        except Exception as e:  # noqa: BLE001
            if len(a) > 1:
                self.vars[a[1]] = str(e)
            return "1"

    @staticmethod
    def _split_list(text: str) -> List[str]:
        out, i, n = [], 0, len(text)
        while i < n:
            while i < n and text[i].isspace():
                i += 1
            if i >= n:
                break
            if text[i] == "{":
                depth, j = 1, i + 1
                while j < n and depth:
                    depth += (text[j] == "{") - (text[j] == "}")
                    j += 1
                out.append(text[i + 1:j - 1])
                i = j
            else:
                j = i
                while j < n and not text[j].isspace():
                    j += 1
                out.append(text[i:j])
                i = j
        return out

    @staticmethod
    def _join_list(items) -> str:
        return " ".join("{" + x + "}" if (not x or any(c.isspace() for c in x)) else x for x in items)

    def _c_lrepeat(self, a):
        return self._join_list(a[1:] * int(a[0]))

    def _c_list(self, a):
        return self._join_list(a)

    def _c_llength(self, a):
        return str(len(self._split_list(a[0])))

    def _c_lindex(self, a):
        return self._split_list(a[0])[int(a[1])]

    def _c_string(self, a):
        if a[0] == "length":
            return str(len(a[1]))
        if a[0] in ("tolower", "toupper"):
            return a[1].lower() if a[0] == "tolower" else a[1].upper()
        raise TclError("string: unsupported subcommand " + a[0])

    def _c_puts(self, a):
        self.output.append(a[-1] if a else "")


# ------------------------------------------------------------------ DRAW / ViewerTest subset

class _Shape:
    """A DRAW shape: tessellated parts with the shape-level transform baked into the vertices."""

    def __init__(self, parts):
        self.parts = parts          # list of (pos, nrm, idx)
        self.uvs = [None] * len(parts)   # texel coordinates per part (PLY meshes), or None
        self.faces = None           # for `explode <box> FACE`

    def copy(self):
        s = _Shape([(p.copy(), n.copy(), i.copy()) for p, n, i in self.parts])
        s.uvs = [None if u is None else u.copy() for u in self.uvs]
        if self.faces:
            s.faces = [f.copy() for f in self.faces]
        return s

    def transform(self, m: np.ndarray):
        R, t = m[:, :3].astype(np.float64), m[:, 3].astype(np.float64)
        self.parts = [((p @ R.T + t).astype(np.float32), (n @ np.linalg.inv(R)).astype(np.float32), i) for p, n, i in self.parts]
        for k, (p, n, i) in enumerate(self.parts):
            ln = np.linalg.norm(n, axis=1, keepdims=True)
            self.parts[k] = (p, (n / np.maximum(ln, 1e-30)).astype(np.float32), i)
        if self.faces:
            for f in self.faces:
                f.transform(m)

    def merged(self):
        return scenes._merge(self.parts)

    def merged_uv(self):
        if all(u is None for u in self.uvs):
            return None
        return np.concatenate([np.zeros((p.shape[0], 2), np.float32) if u is None else u.astype(np.float32)
                               for (p, _, _), u in zip(self.parts, self.uvs)])


class _Object:
    def __init__(self, name, shape):
        self.name = name
        self.shape = shape
        self.location = np.eye(3, 4)
        self.material_name = "default"
        self.bsdf = NAMED_MATERIALS["default"]()
        self.displayed = True


def _quat_to_mat(x, y, z, w):
    n = math.sqrt(x * x + y * y + z * z + w * w) or 1.0
    x, y, z, w = x / n, y / n, z / n, w / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _compose(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """a * b for row-major 3x4 affine matrices."""
    A, B = np.eye(4), np.eye(4)
    A[:3], B[:3] = a, b
    return (A @ B)[:3]


class DrawSession(Interp):
    """Evaluates a script and accumulates the viewer state; `scene()` returns the SceneDesc."""

    SPHERE_RES = (64, 32)

    def __init__(self, width=512, height=512, root: Optional[str] = None):
        super().__init__()
        self.width, self.height = width, height
        self.shapes: Dict[str, _Shape] = {}
        self.objects: Dict[str, _Object] = {}     # displayed AIS objects, insertion ordered
        # V3d_Viewer::SetDefaultLights(): a directional head light (0) and an ambient light (1, no path-traced effect)
        self.lights: List[dict] = [dict(kind="directional", head=True, vec=(0, 0, -1), smooth=0.0, intensity=1.0, color=(1, 1, 1)),
                                   dict(kind="ambient", head=False, vec=(0, 0, 0), smooth=0.0, intensity=1.0, color=(1, 1, 1))]
        self.size_fixed = False                    # True: `vinit w= h=` does not change the size the caller chose
        self.on_dump: Optional[Callable] = None    # on_dump(session, path, frames) renders a `vdump`
        self.dumps: List[tuple] = []               # (path, frames, SceneDesc) of every `vdump` when on_dump is None
        self.params = Graphic3d_RenderingParams()
        self.proj = np.array([0.0, -1.0, 0.0])     # direction from the scene towards the eye
        self.up = np.array([0.0, 0.0, 1.0])
        self.at: Optional[np.ndarray] = None
        self.eye: Optional[np.ndarray] = None
        self.fovy = 45.0
        self.ortho = False
        self.size: Optional[float] = None
        self.distance: Optional[float] = None
        self.fit_requested = True
        self.envmap: Optional[np.ndarray] = None
        self.textures: List[np.ndarray] = []
        self._texture_ids: Dict[str, int] = {}
        self.frames: Optional[int] = None
        if root is not None:
            self.vars["Root"] = root
        for name in ("box psphere pcylinder compound explode ttranslate trotate tcopy vclear vdisplay verase vremove vlocation "
                     "vsetmaterial vbsdf vlight rtlight vcamera vviewparams vfront vback vtop vbottom vleft vright vaxo vfit "
                     "vrenderparams vtextureenv vsetdispmode vinit vfps vdump pload vglinfo vzbufftrihedron vrepaint vupdate "
                     "rtmeshread rtdisplay rterase rttexture vtexture vvbo incmesh vsetlocation").split():
            self.cmds[name] = getattr(self, "_d_" + name, self._d_ignore)

    def _d_ignore(self, a):
        return ""

    # -- shapes
    def _d_box(self, a):
        v = [float(x) for x in a[1:]]
        if len(v) == 3:
            o, d = (0.0, 0.0, 0.0), v
        elif len(v) == 6:
            o, d = v[:3], v[3:]
        else:
            raise TclError("box name [x y z] dx dy dz")
        faces = scenes.box_faces(d[0], d[1], d[2], 1, o)
        s = _Shape([scenes._merge(faces)])
        s.faces = [_Shape([scenes._merge([f])]) for f in faces]
        self.shapes[a[0]] = s

    def _d_psphere(self, a):
        self.shapes[a[0]] = _Shape([scenes.uv_sphere(float(a[1]), *self.SPHERE_RES)])

    def _d_pcylinder(self, a):
        self.shapes[a[0]] = _Shape([scenes.cylinder(float(a[1]), float(a[2]), seg=64)])

    def _shape(self, name) -> _Shape:
        if name not in self.shapes:
            raise TclError(f"shape '{name}' does not exist")
        return self.shapes[name]

    def _d_compound(self, a):
        *members, result = a
        c = _Shape([])
        c.members = [self._shape(m).copy() for m in members]
        for m in c.members:
            c.parts.extend(m.parts)
            c.uvs.extend(m.uvs)
        self.shapes[result] = c

    def _d_explode(self, a):
        s = self._shape(a[0])
        kind = a[1].upper() if len(a) > 1 else ""
        if kind in ("FACE", "F") and s.faces:
            subs = s.faces
        elif hasattr(s, "members"):
            subs = s.members
        else:
            subs = [s]
        names = []
        for k, sub in enumerate(subs, 1):
            self.shapes[f"{a[0]}_{k}"] = sub.copy()
            names.append(f"{a[0]}_{k}")
        return " ".join(names)

    def _d_ttranslate(self, a):
        *names, dx, dy, dz = a
        for n in names:
            self._shape(n).transform(scenes.trsf((float(dx), float(dy), float(dz))))

    def _d_trotate(self, a):
        *names, x, y, z, dx, dy, dz, ang = a
        p = np.array([float(x), float(y), float(z)])
        r = scenes.trsf((0, 0, 0), (float(dx), float(dy), float(dz)), float(ang)).astype(np.float64)
        r[:, 3] = p - r[:, :3] @ p
        for n in names:
            self._shape(n).transform(r)

    def _d_tcopy(self, a):
        self.shapes[a[1]] = self._shape(a[0]).copy()

    # -- CADRays' own DRAW commands (src/ImportExport/ImportExportPlugin.cxx:973-994), mesh subset
    def _d_rtmeshread(self, a):
        """rtmeshread <file> <name> [-group] [-up X|Y|Z|-X|-Y|-Z] ...: PLY files as the exporter writes them."""
        from . import ply
        path, name = a[0], a[1]
        if not os.path.exists(path):
            raise TclError(f"rtmeshread: cannot read '{path}'")
        if not path.lower().endswith(".ply"):
            raise TclError("rtmeshread: only PLY meshes are supported (the exporter's format); OBJ/FBX need Assimp")
        pos, nrm, uv, idx = ply.read_ply(path)
        up = None
        for k, t in enumerate(a):
            if t.lower() == "-up" and k + 1 < len(a):
                up = a[k + 1].upper()
        if up in ("Y", "-Y"):                      # MeshImporter's axis flip to the viewer's Z-up
            sgn = 1.0 if up == "Y" else -1.0
            rot = np.array([[1, 0, 0], [0, 0, -sgn], [0, sgn, 0]], dtype=np.float32)
            pos = pos @ rot.T
            nrm = None if nrm is None else nrm @ rot.T
        if nrm is None:
            e0, e1 = pos[idx[:, 1]] - pos[idx[:, 0]], pos[idx[:, 2]] - pos[idx[:, 0]]
            fn = np.cross(e0, e1)
            acc = np.zeros_like(pos, dtype=np.float64)
            for k in range(3):
                np.add.at(acc, idx[:, k], fn)
            ln = np.linalg.norm(acc, axis=1, keepdims=True)
            nrm = np.where(ln > 0, acc / np.maximum(ln, 1e-30), [0, 0, 1]).astype(np.float32)
        self.shapes[name] = _Shape([(pos.astype(np.float32), nrm.astype(np.float32), idx.astype(np.uint32))])
        self.shapes[name].uvs = [uv]
        return name

    def _d_rttexture(self, a):
        """rttexture <node> [<file>] [-scale S T] [-on|-off]  (ImportExportPlugin.cxx:610-750; exporter
        ImportExport.cxx:259-264); vtexture <node> <file> -scale S T is the AIS_TexturedShape spelling."""
        o = self._obj(a[0])
        i = 1
        while i < len(a):
            t = a[i]
            if t.lower() == "-scale":
                o.bsdf.TextureScale = (float(a[i + 1]), float(a[i + 2])); i += 3
            elif t.lower() == "-off":
                o.bsdf.TextureId = None; i += 1
            elif t.lower() in ("-on", "-noupdate"):
                i += 1
            else:
                if t not in self._texture_ids:
                    if not os.path.exists(t):
                        raise TclError(f"rttexture: failed to find image file at the path '{t}'")
                    from PIL import Image
                    self.textures.append(np.asarray(Image.open(t).convert("RGBA"), dtype=np.uint8))
                    self._texture_ids[t] = len(self.textures) - 1
                o.bsdf.TextureId = self._texture_ids[t]
                i += 1
        return ""

    _d_vtexture = _d_rttexture

    def _d_rtdisplay(self, a):
        self._d_vdisplay(a)

    def _d_rterase(self, a):
        self._d_verase(a)

    # -- viewer content
    def _d_vclear(self, a):
        self.objects.clear()

    def _names(self, a):
        return [x for x in a if not x.startswith("-")]

    def _d_vdisplay(self, a):
        for n in self._names(a):
            if n in self.objects:
                self.objects[n].displayed = True
            else:
                self.objects[n] = _Object(n, self._shape(n))

    def _d_verase(self, a):
        for n in self._names(a):
            if n in self.objects:
                self.objects[n].displayed = False

    def _d_vremove(self, a):
        for n in self._names(a):
            self.objects.pop(n, None)

    def _obj(self, name) -> _Object:
        if name not in self.objects:
            raise TclError(f"object '{name}' is not displayed")
        return self.objects[name]

    def _d_vlocation(self, a):
        obj = self._obj(self._names(a[:1] if not a[0].startswith("-") else a)[0])
        i = 0
        while i < len(a):
            t = a[i].lower()
            if t in ("-setlocation", "-location"):
                obj.location[:, 3] = [float(a[i + 1]), float(a[i + 2]), float(a[i + 3])]
                i += 4
            elif t == "-rotate":
                x, y, z, dx, dy, dz, ang = (float(v) for v in a[i + 1:i + 8])
                r = scenes.trsf((0, 0, 0), (dx, dy, dz), ang).astype(np.float64)
                p = np.array([x, y, z])
                r[:, 3] = p - r[:, :3] @ p
                obj.location = _compose(obj.location, r)      # LocalTransformation() * rotation
                i += 8
            elif t in ("-rotation", "-setrotation"):
                obj.location[:, :3] = _quat_to_mat(*(float(v) for v in a[i + 1:i + 5]))
                i += 5
            elif t == "-reset":
                obj.location = np.eye(3, 4)
                i += 1
            else:
                i += 1

    def _d_vsetmaterial(self, a):
        names = self._names(a)
        mat = names[-1].lower()
        if mat not in NAMED_MATERIALS:
            raise TclError(f"unknown material '{names[-1]}'")
        for n in names[:-1]:
            o = self._obj(n)
            o.material_name = mat
            o.bsdf = NAMED_MATERIALS[mat]()

    @staticmethod
    def _fresnel(a, i):
        kind = a[i].lower()
        if kind == "constant":
            return Graphic3d_Fresnel.CreateConstant(float(a[i + 1])), i + 2
        if kind == "schlick":
            return Graphic3d_Fresnel.CreateSchlick(float(a[i + 1]), float(a[i + 2]), float(a[i + 3])), i + 4
        if kind == "conductor":
            return Graphic3d_Fresnel.CreateConductor(float(a[i + 1]), float(a[i + 2])), i + 3
        if kind == "dielectric":
            return Graphic3d_Fresnel.CreateDielectric(float(a[i + 1])), i + 2
        raise TclError("unknown Fresnel model " + a[i])

    @staticmethod
    def _is_num(s):
        try:
            float(s)
            return True
        except ValueError:
            return False

    def _color(self, a, i):
        """1 or 3 numbers after a flag (`-kd 0.85` is the scalar shorthand, Materials.tcl:31)."""
        vals = []
        while i < len(a) and len(vals) < 3 and self._is_num(a[i]):
            vals.append(float(a[i]))
            i += 1
        if len(vals) == 1:
            vals = vals * 3
        if len(vals) != 3:
            raise TclError("expected 1 or 3 colour components")
        return vals, i

    def _d_vbsdf(self, a):
        o = self._obj(a[0])
        b = o.bsdf
        i = 1
        normalize = False
        while i < len(a):
            t = a[i].lower()
            if t in ("-kc", "-kd", "-ks", "-kt", "-le", "-absorpcolor"):
                vals, i = self._color(a, i + 1)
                tgt = {"-kc": b.Kc, "-kd": b.Kd, "-ks": b.Ks, "-kt": b.Kt, "-le": b.Le, "-absorpcolor": b.Absorption}[t]
                tgt[0:3] = vals
            elif t == "-baseroughness":
                b.Ks[3] = float(a[i + 1]); i += 2
            elif t == "-coatroughness":
                b.Kc[3] = float(a[i + 1]); i += 2
            elif t == "-absorpcoeff":
                b.Absorption[3] = float(a[i + 1]); i += 2
            elif t == "-coatfresnel":
                b.FresnelCoat, i = self._fresnel(a, i + 1)
            elif t == "-basefresnel":
                b.FresnelBase, i = self._fresnel(a, i + 1)
            elif t in ("-n", "-normalize"):
                normalize = True; i += 1
            elif t in ("-noupdate", "-update"):
                i += 1
            else:
                raise TclError(f"vbsdf: unknown option {a[i]}")
        if normalize:
            b.Normalize()
        return ""

    # -- lights
    def _d_vlight(self, a):
        if not a:
            return ""
        sub = a[0].lower()
        if sub == "clear":
            self.lights = []
            return ""
        if sub in ("add", "new"):
            kind = a[1].lower()
            l = dict(kind=kind, head=False, vec=(0.0, 0.0, -1.0) if kind == "directional" else (0.0, 0.0, 0.0),
                     smooth=0.0, intensity=1.0, color=(1.0, 1.0, 1.0))
            self.lights.append(l)
            self._light_opts(l, a[2:])
            return str(len(self.lights) - 1)
        if sub == "change":
            self._light_opts(self.lights[int(a[1])], a[2:])
            return ""
        if sub in ("del", "delete", "remove"):
            self.lights.pop(int(a[1]))
            return ""
        raise TclError("vlight: unsupported subcommand " + a[0])

    def _light_opts(self, l, a):
        i = 0
        while i < len(a):
            t = a[i].lower().lstrip("-")
            if t in ("pos", "position", "dir", "direction"):
                l["vec"] = tuple(float(v) for v in a[i + 1:i + 4]); i += 4
            elif t in ("sm", "smoothness"):
                l["smooth"] = float(a[i + 1]); i += 2
            elif t in ("int", "intensity"):
                l["intensity"] = float(a[i + 1]); i += 2
            elif t in ("head", "headlight"):
                l["head"] = bool(int(a[i + 1])); i += 2
            elif t in ("color", "colour"):
                if self._is_num(a[i + 1]):
                    l["color"] = tuple(float(v) for v in a[i + 1:i + 4]); i += 4
                else:
                    i += 2
            else:
                i += 1

    def _d_rtlight(self, a):
        l = self.lights[int(a[0])]
        self._light_opts(l, a[1:])

    # -- camera
    def _d_vcamera(self, a):
        i = 0
        while i < len(a):
            t = a[i].lower()
            if t in ("-persp", "-perspective"):
                self.ortho = False; i += 1
            elif t in ("-ortho", "-orthographic"):
                self.ortho = True; i += 1
            elif t in ("-fovy", "-fov"):
                self.fovy = float(a[i + 1]); i += 2
            elif t in ("-distance", "-dist"):
                self.distance = float(a[i + 1]); i += 2
            else:
                i += 1

    def _d_vviewparams(self, a):
        i = 0
        while i < len(a):
            t = a[i].lower()
            if t in ("-proj", "-up", "-at", "-eye"):
                v = np.array([float(x) for x in a[i + 1:i + 4]])
                setattr(self, t[1:], v)
                self.fit_requested = False if t in ("-at", "-eye") else self.fit_requested
                i += 4
            elif t in ("-size", "-scale"):
                if t == "-size":
                    self.size = float(a[i + 1])
                i += 2
            else:
                i += 1

    def _view(self, proj, up):
        self.proj, self.up = np.array(proj, float), np.array(up, float)
        self.eye = self.at = None
        self.fit_requested = True

    def _d_vfront(self, a): self._view((0, -1, 0), (0, 0, 1))
    def _d_vback(self, a): self._view((0, 1, 0), (0, 0, 1))
    def _d_vtop(self, a): self._view((0, 0, 1), (0, 1, 0))
    def _d_vbottom(self, a): self._view((0, 0, -1), (0, -1, 0))
    def _d_vleft(self, a): self._view((-1, 0, 0), (0, 0, 1))
    def _d_vright(self, a): self._view((1, 0, 0), (0, 0, 1))
    def _d_vaxo(self, a): self._view((1, -1, 1), (0, 0, 1))

    def _d_vfit(self, a):
        self.eye = self.at = None
        self.fit_requested = True

    def _d_vrenderparams(self, a):
        i = 0
        p = self.params
        while i < len(a):
            t = a[i].lower()
            nxt = a[i + 1] if i + 1 < len(a) else ""
            if t in ("-raydepth", "-reflections") and self._is_num(nxt):
                p.RaytracingDepth = int(float(nxt)); i += 2
            elif t in ("-gi", "-ray", "-raytrace", "-shadows", "-refl", "-fsaa", "-rebuildglsl", "-rebuild"):
                i += 1 + (1 if nxt.lower() in ("on", "off", "0", "1") else 0)
            elif t in ("-rasterization", "-raster"):
                raise TclError("rasterization is out of scope: only the path-traced mode is implemented")
            elif t in ("-iss", "-adaptive"):      # CornellBox.tcl:78-79
                p.AdaptiveScreenSampling = nxt.lower() not in ("off", "0"); i += 1 + (1 if nxt.lower() in ("on", "off", "0", "1") else 0)
            elif t in ("-nbtiles", "-tiles") and self._is_num(nxt):
                p.NbRayTracingTiles = int(float(nxt)); i += 2
            elif t == "-issd":                    # ShowSamplingTiles
                p.ShowSamplingTiles = nxt.lower() not in ("off", "0"); i += 1 + (1 if nxt.lower() in ("on", "off", "0", "1") else 0)
            elif t in ("-maxrad", "-radianceclamping") and self._is_num(nxt):
                p.RadianceClampingValue = float(nxt); i += 2
            elif t in ("-twoside", "-twosided"):
                p.TwoSidedBsdfModels = nxt.lower() not in ("off", "0"); i += 1 + (1 if nxt.lower() in ("on", "off", "0", "1") else 0)
            elif t in ("-coherent", "-brng"):
                p.CoherentPathTracingMode = nxt.lower() not in ("off", "0"); i += 1 + (1 if nxt.lower() in ("on", "off", "0", "1") else 0)
            elif t == "-env":
                p.UseEnvironmentMapBackground = nxt.lower() not in ("off", "0"); i += 2
            elif t in ("-aperture",) and self._is_num(nxt):
                p.CameraApertureRadius = float(nxt); i += 2
            elif t in ("-focal",) and self._is_num(nxt):
                p.CameraFocalPlaneDist = float(nxt); i += 2
            elif t == "-exposure" and self._is_num(nxt):
                p.Exposure = float(nxt); i += 2
            elif t == "-whitepoint" and self._is_num(nxt):
                p.WhitePoint = float(nxt); i += 2
            elif t == "-tonemapping":
                p.ToneMappingMethod = Graphic3d_ToneMappingMethod_Filmic if nxt.lower() == "filmic" else 0; i += 2
            elif t in ("-spp", "-samples") and self._is_num(nxt):
                p.SamplesPerPixel = int(float(nxt)); i += 2
            else:
                i += 1

    def _d_vtextureenv(self, a):
        if a and a[0].lower() == "off":
            self.envmap = None
            return
        path = a[-1]
        if not os.path.exists(path):
            # OCCT reports the unreadable file and goes on without a map (preview.tcl:56 names a file of its author's disk)
            self.output.append(f"vtextureenv: cannot read '{path}', no environment map")
            self.envmap = None
            return
        ext = os.path.splitext(path)[1].lower()
        if ext in (".hdr", ".pic"):              # floating-point maps keep their range (crt_envmap_set_rgb32f)
            from .imageio import read_hdr
            self.envmap = read_hdr(path)
        elif ext == ".pfm":
            from .imageio import read_pfm
            self.envmap = read_pfm(path)
        elif ext == ".png":
            from .imageio import read_png_rgb8
            self.envmap = read_png_rgb8(path)
        else:
            from PIL import Image
            self.envmap = np.asarray(Image.open(path).convert("RGB"), dtype=np.uint8)

    def _d_vfps(self, a):
        if a and self._is_num(a[0]):
            self.frames = int(float(a[0]))

    def _d_vinit(self, a):
        """`vinit name=View1 w=128 h=128` (data/other/preview.tcl:10): the window size, unless the caller fixed one."""
        for t in a:
            k, _, v = t.partition("=")
            k = k.lower().lstrip("-")
            if k in ("w", "width") and v.isdigit() and not self.size_fixed:
                self.width = int(v)
            elif k in ("h", "height") and v.isdigit() and not self.size_fixed:
                self.height = int(v)

    def _d_vsetlocation(self, a):
        """`vsetlocation [-noupdate] name x y z` (data/other/preview.tcl:23): translation of the object."""
        names = [t for t in a if not t.startswith("-") or self._is_num(t)]
        if len(names) != 4:
            raise TclError("vsetlocation name x y z")
        self._obj(names[0]).location[:, 3] = [float(v) for v in names[1:]]

    def _d_vdump(self, a):
        """`vfps N` + `vdump file` is how preview.tcl renders one icon per material (:62-65): each vdump hands the
        current scene and frame count to `on_dump` (a renderer), or records them in `dumps`."""
        paths = [t for t in a if not t.startswith("-")]
        if not paths:
            raise TclError("vdump file")
        if self.on_dump is not None:
            self.on_dump(self, paths[0], self.frames)
        else:
            import copy
            self.dumps.append((paths[0], self.frames, copy.deepcopy(self.scene())))

    # -- result
    def scene(self) -> scenes.SceneDesc:
        s = scenes.SceneDesc("tcl", width=self.width, height=self.height)
        lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
        for o in self.objects.values():
            if not o.displayed:
                continue
            pos, nrm, idx = o.shape.merged()
            s.add((pos, nrm, idx), o.location.astype(np.float32), o.bsdf)
            uv = o.shape.merged_uv()
            if uv is not None:
                s.mesh_uvs[len(s.meshes) - 1] = uv
            w = pos.astype(np.float64) @ o.location[:, :3].T + o.location[:, 3]
            lo, hi = np.minimum(lo, w.min(0)), np.maximum(hi, w.max(0))
        if not np.isfinite(lo).all():
            lo, hi = np.zeros(3), np.ones(3)
        s.params = self.params
        s.envmap = self.envmap
        s.textures = list(self.textures)
        proj = self.proj / np.linalg.norm(self.proj)
        auto_size = None
        if self.eye is not None and self.at is not None:
            eye, at = self.eye, self.at
        else:
            at = 0.5 * (lo + hi) if self.at is None else self.at
            radius = 0.5 * float(np.linalg.norm(hi - lo))
            if self.distance is not None and not self.fit_requested:
                dist = self.distance
            elif self.ortho:
                dist = 2.0 * radius + 1.0
            else:
                half = math.radians(self.fovy) * 0.5
                aspect = self.width / self.height
                half_min = math.atan(math.tan(half) * min(1.0, aspect))
                dist = radius / math.sin(half_min)
            eye = at + proj * dist
            if self.size is None:
                auto_size = 2.0 * radius   # per call: a later vdump of a changed scene fits the scene as it is then
        s.camera = Graphic3d_Camera(Eye=tuple(eye), Direction=tuple(np.asarray(at) - np.asarray(eye)), Up=tuple(self.up),
                                    FOVy=self.fovy, IsOrthographic=self.ortho, Scale=self.size or auto_size or 1.0)
        fwd = np.asarray(at) - np.asarray(eye)
        fwd = fwd / np.linalg.norm(fwd)
        for l in self.lights:
            if l["kind"] == "directional":
                d = fwd if l["head"] else np.array(l["vec"], float)
                s.lights.append(make_light(False, d, l["color"], l["intensity"], l["smooth"]))
            elif l["kind"] in ("positional", "spotlight"):
                p = np.array(l["vec"], float) + (np.asarray(eye) if l["head"] else 0.0)
                s.lights.append(make_light(True, p, l["color"], l["intensity"], l["smooth"]))
            # ambient lights have no path-traced counterpart (the GUI hides them, LightSourcesEditor.cxx:157-178)
        return s


def load_script(path: str, width=512, height=512, strict=False) -> DrawSession:
    sess = DrawSession(width, height, root=os.path.dirname(os.path.abspath(path)))
    sess.strict = strict
    with open(path, "r", encoding="utf-8", errors="replace") as f:
        sess.eval(f.read())
    return sess
