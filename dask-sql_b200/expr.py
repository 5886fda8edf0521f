"""Host-side expression IR of the B200 layer.

The reference turns every REX node into one pandas call on whole Series
(physical/rex/core/call.py:1158-1216).  Here the same operators build a small typed tree; at
execution time a tree is either recognised as a conjunction of `column <cmp> literal` terms
(evaluated inside the fused scan kernels) or compiled to the postfix program of b2_expr_eval.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L

I64, F64, U8 = L.I64, L.F64, L.U8
_DT_NAME = {I64: "int64", F64: "float64", U8: "bool"}

_CMP = {"eq": L.EQ, "ne": L.NE, "lt": L.LT, "le": L.LE, "gt": L.GT, "ge": L.GE}
_FLIP = {"eq": "eq", "ne": "ne", "lt": "gt", "le": "ge", "gt": "lt", "ge": "le"}


class Expr:
    dtype: int = I64

    def refs(self, out=None):
        out = set() if out is None else out
        self._refs(out)
        return out

    def _refs(self, out):
        pass


class ColRef(Expr):
    __slots__ = ("name", "dtype", "logical")

    def __init__(self, name, dtype, logical=None):
        self.name, self.dtype, self.logical = name, dtype, logical or _DT_NAME[dtype]

    def _refs(self, out):
        out.add(self.name)

    def __repr__(self):
        return f"col({self.name})"


class Lit(Expr):
    __slots__ = ("value", "dtype")

    def __init__(self, value, dtype=None):
        if dtype is None:
            if value is None:
                dtype = F64
            elif isinstance(value, (bool, np.bool_)):
                dtype, value = U8, bool(value)
            elif isinstance(value, (int, np.integer)):
                dtype, value = I64, int(value)
            elif isinstance(value, (float, np.floating)):
                dtype, value = F64, float(value)
            else:
                raise NotImplementedError(f"literal {value!r} of type {type(value).__name__} is outside the "
                                          "int64/float64/bool hot path")
        self.value, self.dtype = value, dtype

    def __repr__(self):
        return f"lit({self.value!r})"


class Call(Expr):
    __slots__ = ("op", "args", "dtype")

    def __init__(self, op, args, dtype):
        self.op, self.args, self.dtype = op, tuple(args), dtype

    def _refs(self, out):
        for a in self.args:
            a._refs(out)

    def __repr__(self):
        return f"{self.op}({', '.join(map(repr, self.args))})"


def as_expr(x) -> Expr:
    return x if isinstance(x, Expr) else Lit(x)


def cast(e: Expr, dtype: int) -> Expr:
    if e.dtype == dtype:
        return e
    if isinstance(e, Lit):
        if e.value is None:
            return Lit(None, dtype)
        if dtype == F64:
            return Lit(float(e.value), F64)
        if dtype == I64:
            return Lit(int(e.value), I64)
        return Lit(bool(e.value), U8)
    return Call("cast", [e], dtype)


def _arith_type(a: Expr, b: Expr):
    return F64 if F64 in (a.dtype, b.dtype) else I64


def binop(op: str, a, b) -> Expr:
    a, b = as_expr(a), as_expr(b)
    if op in ("add", "sub", "mul", "mod"):
        t = _arith_type(a, b)
        return Call(op, [cast(a, t), cast(b, t)], t)
    if op == "truediv":
        return Call("truediv", [cast(a, F64), cast(b, F64)], F64)
    if op == "divt":  # SQL integer division: truncates toward zero (call.py:165-189)
        t = _arith_type(a, b)
        if t == F64:
            return Call("truediv", [cast(a, F64), cast(b, F64)], F64)
        return Call("divt", [cast(a, I64), cast(b, I64)], I64)
    if op in _CMP:
        t = _arith_type(a, b)
        return Call(op, [cast(a, t), cast(b, t)], U8)
    if op in ("and", "or"):
        return Call(op, [cast(a, U8), cast(b, U8)], U8)
    raise NotImplementedError(f"operator {op}")


def unop(op: str, a) -> Expr:
    a = as_expr(a)
    if op == "neg":
        return Call("neg", [cast(a, I64) if a.dtype == U8 else a], F64 if a.dtype == F64 else I64)
    if op == "abs":
        return Call("abs", [a], a.dtype)
    if op == "sqrt":
        return Call("sqrt", [cast(a, F64)], F64)
    if op == "not":
        return Call("not", [cast(a, U8)], U8)
    if op == "isnull":
        if isinstance(a, Lit):
            return Lit(a.value is None or (isinstance(a.value, float) and a.value != a.value))
        return Call("isnull", [a], U8)
    raise NotImplementedError(f"operator {op}")


def case(cond, then, other) -> Expr:
    cond, then, other = as_expr(cond), as_expr(then), as_expr(other)
    if isinstance(then, Lit) and then.value is None:
        t = other.dtype
    elif isinstance(other, Lit) and other.value is None:
        t = then.dtype
    else:
        t = F64 if F64 in (then.dtype, other.dtype) else (I64 if I64 in (then.dtype, other.dtype) else U8)
    return Call("case", [cast(cond, U8), cast(then, t), cast(other, t)], t)


def fillna(a, fill) -> Expr:
    a, fill = as_expr(a), as_expr(fill)
    return Call("fillna", [a, cast(fill, a.dtype)], a.dtype)


def substitute(e: Expr, mapping: Dict[str, Expr]) -> Expr:
    """Rewrite column references through `mapping` (composition of projections)."""
    if isinstance(e, ColRef):
        return mapping[e.name]
    if isinstance(e, Call):
        return Call(e.op, [substitute(a, mapping) for a in e.args], e.dtype)
    return e


def conjuncts(e: Expr) -> List[Expr]:
    """Flatten AND under 'is TRUE' semantics; fillna(x, False) is the identity there
    (filter.py:38-39: a NULL predicate drops the row)."""
    if isinstance(e, Call):
        if e.op == "and":
            return conjuncts(e.args[0]) + conjuncts(e.args[1])
        if e.op == "fillna" and isinstance(e.args[1], Lit) and e.args[1].value in (False, 0):
            return conjuncts(e.args[0])
    return [e]


def _strip_cast(e: Expr):
    """cast(ColRef int -> f64) compares as float: report (colref, as_f64)."""
    if isinstance(e, Call) and e.op == "cast" and e.dtype == F64 and isinstance(e.args[0], ColRef) \
            and e.args[0].dtype == I64:
        return e.args[0], True
    if isinstance(e, ColRef):
        return e, False
    return None, False


def as_term(e: Expr) -> Optional[Tuple[str, int, object]]:
    """Recognise `col <cmp> literal`, `col IS [NOT] NULL`, or a boolean column.
    Returns (column name, B2 term op, literal) or None."""
    if isinstance(e, ColRef) and e.dtype == U8:
        return e.name, L.IS_TRUE, 0
    if not isinstance(e, Call):
        return None
    if e.op in _CMP:
        a, b = e.args
        op = e.op
        if isinstance(a, Lit) and not isinstance(b, Lit):
            a, b, op = b, a, _FLIP[op]
        if isinstance(b, Lit) and b.value is not None:
            col, as_f = _strip_cast(a)
            if col is None or col.dtype == U8:
                return None
            lit = b.value
            if as_f or col.dtype == F64:
                lit = float(lit)
                if col.dtype == I64 and lit.is_integer() and abs(lit) < 2 ** 62:
                    lit = int(lit)
            return col.name, _CMP[op], lit
        return None
    if e.op == "isnull" and isinstance(e.args[0], ColRef):
        return e.args[0].name, L.IS_NULL, 0
    if e.op == "not" and isinstance(e.args[0], Call) and e.args[0].op == "isnull" \
            and isinstance(e.args[0].args[0], ColRef):
        return e.args[0].args[0].name, L.IS_NOT_NULL, 0
    return None


# ---------------------------------------------------------------------------------------------
# compilation to the postfix program of b2_expr_eval
# ---------------------------------------------------------------------------------------------
class _Compiler:
    def __init__(self, col_index: Dict[str, int]):
        self.col_index = col_index
        self.code: List[Tuple[int, int, int, float]] = []

    def emit(self, op, a=0, imm_i=0, imm_f=0.0):
        self.code.append((op, a, imm_i, imm_f))

    def lit(self, e: Lit):
        if e.value is None:
            self.emit(L.OP_CONST_NULL)
        elif e.dtype == F64:
            self.emit(L.OP_CONST_F, imm_f=float(e.value))
        else:
            self.emit(L.OP_CONST_I, imm_i=int(e.value))

    def visit(self, e: Expr):
        if isinstance(e, ColRef):
            self.emit(L.OP_LOAD, a=self.col_index[e.name])
            return
        if isinstance(e, Lit):
            self.lit(e)
            return
        op, args = e.op, e.args
        if op == "cast":
            src = args[0]
            self.visit(src)
            if e.dtype == F64 and src.dtype != F64:
                self.emit(L.OP_I2F)
            elif e.dtype == I64 and src.dtype == F64:
                self.emit(L.OP_F2I)
            elif e.dtype == U8 and src.dtype == I64:
                self.emit(L.OP_CONST_I, imm_i=0)
                self.emit(L.OP_EQ_I + L.NE)
            elif e.dtype == U8 and src.dtype == F64:
                self.emit(L.OP_CONST_F, imm_f=0.0)
                self.emit(L.OP_EQ_F + L.NE)
            return
        if op in ("add", "sub", "mul", "mod", "truediv", "divt"):
            self.visit(args[0])
            self.visit(args[1])
            f = e.dtype == F64
            table = {"add": (L.OP_ADD_I, L.OP_ADD_F), "sub": (L.OP_SUB_I, L.OP_SUB_F),
                     "mul": (L.OP_MUL_I, L.OP_MUL_F), "truediv": (None, L.OP_DIV_F),
                     "divt": (L.OP_DIV_I, None), "mod": (L.OP_MOD_I, None)}
            code = table[op][1 if f else 0]
            if code is None:
                raise NotImplementedError(f"{op} on {_DT_NAME[e.dtype]}")
            self.emit(code)
            return
        if op in _CMP:
            self.visit(args[0])
            self.visit(args[1])
            base = L.OP_EQ_F if args[0].dtype == F64 else L.OP_EQ_I
            self.emit(base + _CMP[op])
            return
        if op in ("and", "or"):
            self.visit(args[0])
            self.visit(args[1])
            self.emit(L.OP_AND if op == "and" else L.OP_OR)
            return
        if op == "not":
            self.visit(args[0])
            self.emit(L.OP_NOT)
            return
        if op == "neg":
            self.visit(args[0])
            self.emit(L.OP_NEG_F if e.dtype == F64 else L.OP_NEG_I)
            return
        if op == "abs":
            self.visit(args[0])
            self.emit(L.OP_ABS_F if e.dtype == F64 else L.OP_ABS_I)
            return
        if op == "sqrt":
            self.visit(args[0])
            self.emit(L.OP_SQRT_F)
            return
        if op == "isnull":
            self.visit(args[0])
            self.emit(L.OP_ISNULL_F if args[0].dtype == F64 else L.OP_ISNULL_I)
            return
        if op == "case":
            for a in args:
                self.visit(a)
            self.emit(L.OP_CASE)
            return
        if op == "fillna":
            # NaN in a float operand is NULL for fillna: route through CASE(isnull(x), fill, x)
            if args[0].dtype == F64:
                self.visit(args[0])
                self.emit(L.OP_ISNULL_F)
                self.visit(args[1])
                self.visit(args[0])
                self.emit(L.OP_CASE)
            else:
                self.visit(args[0])
                self.visit(args[1])
                self.emit(L.OP_FILLNA)
            return
        if op == "ord2f":
            self.visit(args[0])
            self.emit(L.OP_ORD2F)
            return
        raise NotImplementedError(f"expression operator {op}")


def compile_expr(e: Expr, col_names: Sequence[str]) -> L.Prog:
    """Compile `e` over the columns `col_names` (their order defines LOAD indices)."""
    comp = _Compiler({n: i for i, n in enumerate(col_names)})
    comp.visit(e)
    if len(comp.code) > L.MAX_PROG:
        raise NotImplementedError(f"expression needs {len(comp.code)} instructions (max {L.MAX_PROG})")
    p = L.Prog()
    p.n = len(comp.code)
    p.out_dtype = e.dtype
    for i, (op, a, ii, ff) in enumerate(comp.code):
        ins = p.code[i]
        ins.op, ins.a, ins.imm_i, ins.imm_f = op, a, ii, ff
    return p


def may_be_null(e: Expr, col_nullable) -> bool:
    """Conservative: can the result carry a validity bitmap NULL?"""
    if isinstance(e, ColRef):
        return col_nullable(e.name)
    if isinstance(e, Lit):
        return e.value is None
    if e.op == "isnull":
        return False
    if e.op in ("divt", "mod"):
        return True
    if e.op == "cast" and e.dtype == I64 and e.args[0].dtype == F64:
        return True
    return any(may_be_null(a, col_nullable) for a in e.args)
