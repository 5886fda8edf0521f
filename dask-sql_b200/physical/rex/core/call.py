"""Operator calls (dask_sql/physical/rex/core/call.py:1029-1216), restricted to the operators of
the int64/float64/bool hot path: comparisons, boolean logic, arithmetic,
IS [NOT] NULL / TRUE / FALSE, BETWEEN, IN (list), CAST, CASE, negative, abs.

Operands are LazySeries (device expressions) or python scalars; every operation only BUILDS the
expression tree — evaluation happens fused inside the consuming kernel at compute time.
"""
import logging
import operator
from functools import partial, reduce

import numpy as np

from ....mappings import SqlTypeName, cast_column_to_type, sql_to_python_type, sql_to_python_value
from ....utils import LoggableDataFrame, is_frame
from ..base import BaseRexPlugin
from ..convert import RexConverter

logger = logging.getLogger(__name__)


class Operation:
    """Wrapper around a callable used as SQL operator (call.py:59-103)."""

    needs_dc = False
    needs_rex = False
    needs_context = False
    needs_rel = False

    @staticmethod
    def op_needs_dc(op):
        return getattr(op, "needs_dc", False)

    @staticmethod
    def op_needs_rex(op):
        return getattr(op, "needs_rex", False)

    @staticmethod
    def op_needs_context(op):
        return getattr(op, "needs_context", False)

    @staticmethod
    def op_needs_rel(op):
        return getattr(op, "needs_rel", False)

    def __init__(self, f):
        self.f = f

    def __call__(self, *operands, **kwargs):
        return self.f(*operands, **kwargs)

    def of(self, op: "Operation") -> "Operation":
        new_op = Operation(lambda *x, **kwargs: self(op(*x, **kwargs)))
        new_op.needs_dc = Operation.op_needs_dc(op)
        new_op.needs_rex = Operation.op_needs_rex(op)
        new_op.needs_context = Operation.op_needs_context(op)
        new_op.needs_rel = Operation.op_needs_rel(op)
        return new_op


class ReduceOperation(Operation):
    """n-ary operator applied by reduction over the operands (call.py:140-162)."""

    def __init__(self, operation, unary_operation=None):
        self.operation = operation
        self.unary_operation = unary_operation or operation
        self.needs_dc = Operation.op_needs_dc(self.operation)
        self.needs_rex = Operation.op_needs_rex(self.operation)
        super().__init__(self.reduce)

    def reduce(self, *operands, **kwargs):
        if len(operands) > 1:
            return reduce(partial(self.operation, **kwargs), operands)
        return self.unary_operation(*operands, **kwargs)


def _null_safe(f):
    """SQL: any NULL scalar operand makes a scalar comparison / arithmetic NULL."""
    def g(a, b):
        if a is None and not is_frame(b):
            return None
        if b is None and not is_frame(a):
            return None
        return f(a, b)
    return g


def _and(a, b):
    if not is_frame(a) and not is_frame(b):
        if a is False or b is False:
            return False
        if a is None or b is None:
            return None
        return bool(a) and bool(b)
    return operator.and_(a, b)


def _or(a, b):
    if not is_frame(a) and not is_frame(b):
        if a is True or b is True:
            return True
        if a is None or b is None:
            return None
        return bool(a) or bool(b)
    return operator.or_(a, b)


class SQLDivisionOperator(Operation):
    """SQL '/' truncates toward zero for integer results (call.py:165-189)."""

    needs_rex = True

    def __init__(self):
        super().__init__(self.div)

    def div(self, lhs, rhs, rex=None):
        output_type = sql_to_python_type(SqlTypeName.fromString(str(rex.getType()).upper()))
        is_float = np.issubdtype(output_type, np.floating)
        if is_frame(lhs):
            return lhs / rhs if is_float else lhs.sql_div(rhs)
        if is_frame(rhs):
            return lhs / rhs if is_float else rhs.sql_div(lhs, rev=True)
        if lhs is None or rhs is None:
            return None
        if is_float:
            return lhs / rhs
        return int(np.trunc(lhs / rhs)) if rhs != 0 else None


class CaseOperation(Operation):
    """CASE WHEN ... (call.py:212-253): operands = when, then[, when, then ...][, else]."""

    def __init__(self):
        super().__init__(self.case)

    def case(self, *operands):
        assert operands
        where, then = operands[0], operands[1]
        if len(operands) > 3:
            other = self.case(*operands[2:])
        elif len(operands) == 2:
            other = None
        else:
            other = operands[2]
        if is_frame(then):
            return then.where(where, other=other)
        if is_frame(where):
            from ....frame import LazySeries
            from .... import expr as E
            return LazySeries(where.source, where.pred,
                              E.case(where.expr, E.as_expr(then) if not is_frame(then) else then.expr,
                                     other.expr if is_frame(other) else E.as_expr(other)))
        # `where` is a scalar here: the CASE folds to one branch
        return then if where else other


class CastOperation(Operation):
    """CAST(x AS type) (call.py:256-292)."""

    needs_rex = True

    def __init__(self):
        super().__init__(self.cast)

    def cast(self, operand, rex=None):
        sql_type = SqlTypeName.fromString(rex.getType())
        if not is_frame(operand):
            return sql_to_python_value(sql_type, operand)
        python_type = sql_to_python_type(sql_type)
        out = cast_column_to_type(operand, python_type)
        return operand if out is None else out


class IsFalseOperation(Operation):
    def __init__(self):
        super().__init__(self.false_)

    def false_(self, df):
        if is_frame(df):
            return ~(df.astype("boolean").fillna(True))
        return df is not None and not bool(df)


class IsTrueOperation(Operation):
    def __init__(self):
        super().__init__(self.true_)

    def true_(self, df):
        if is_frame(df):
            return df.astype("boolean").fillna(False)
        return df is not None and bool(df)


class NegativeOperation(Operation):
    def __init__(self):
        super().__init__(lambda df: None if df is None else -df)


class NotOperation(Operation):
    """NOT x (call.py:348-364)."""

    def __init__(self):
        super().__init__(self.not_)

    def not_(self, df):
        if is_frame(df):
            return ~(df.astype("boolean"))
        return None if df is None else not df


class IsNullOperation(Operation):
    """x IS NULL (call.py:367-383); NaN counts as NULL for floats, as in pandas isna()."""

    def __init__(self):
        super().__init__(self.null)

    def null(self, df):
        if is_frame(df):
            return df.isna()
        return df is None or (isinstance(df, float) and df != df)


class BetweenOperation(Operation):
    """x [NOT] BETWEEN low AND high, bounds inclusive (call.py:963-978)."""

    needs_rex = True

    def __init__(self):
        super().__init__(self.between)

    def between(self, series, low, high, rex=None):
        if is_frame(series):
            res = series.between(low, high, inclusive="both")
            return ~res if rex.isNegated() else res
        res = (series >= low) & (series <= high) if is_frame(low) or is_frame(high) else low <= series <= high
        return ~res if (rex.isNegated() and is_frame(res)) else ((not res) if rex.isNegated() else res)


class InListOperation(Operation):
    """x [NOT] IN (v1, v2, ...) (call.py:981-993)."""

    needs_rex = True

    def __init__(self):
        super().__init__(self.inList)

    def inList(self, series, *operands, rex=None):
        if is_frame(series):
            result = series.isin(operands)
            return ~result if rex.isNegated() else result
        result = series in operands
        return (not result) if rex.isNegated() else result


def _abs(x):
    return x.abs() if is_frame(x) else (None if x is None else abs(x))


class RexCallPlugin(BaseRexPlugin):
    """Operator name -> Operation (call.py:1047-1156, hot-path rows)."""

    class_name = "RexCall"

    OPERATION_MAPPING = {
        "between": BetweenOperation(),
        "and": ReduceOperation(operation=_and),
        "or": ReduceOperation(operation=_or),
        ">": ReduceOperation(operation=_null_safe(operator.gt)),
        ">=": ReduceOperation(operation=_null_safe(operator.ge)),
        "<": ReduceOperation(operation=_null_safe(operator.lt)),
        "<=": ReduceOperation(operation=_null_safe(operator.le)),
        "=": ReduceOperation(operation=_null_safe(operator.eq)),
        "!=": ReduceOperation(operation=_null_safe(operator.ne)),
        "<>": ReduceOperation(operation=_null_safe(operator.ne)),
        "+": ReduceOperation(operation=_null_safe(operator.add), unary_operation=lambda x: x),
        "-": ReduceOperation(operation=_null_safe(operator.sub), unary_operation=lambda x: -x),
        "/": ReduceOperation(operation=SQLDivisionOperator()),
        "*": ReduceOperation(operation=_null_safe(operator.mul)),
        "%": ReduceOperation(operation=_null_safe(operator.mod)),
        "cast": CastOperation(),
        "case": CaseOperation(),
        "negative": NegativeOperation(),
        "not": NotOperation(),
        "in list": InListOperation(),
        "is null": IsNullOperation(),
        "is not null": NotOperation().of(IsNullOperation()),
        "is true": IsTrueOperation(),
        "is not true": NotOperation().of(IsTrueOperation()),
        "is false": IsFalseOperation(),
        "is not false": NotOperation().of(IsFalseOperation()),
        "is unknown": IsNullOperation(),
        "is not unknown": NotOperation().of(IsNullOperation()),
        "abs": Operation(_abs),
    }

    def convert(self, rel, expr, dc, context):
        operands = [RexConverter.convert(rel, o, dc, context=context) for o in expr.getOperands()]
        schema_name = context.schema_name
        operator_name = expr.getOperatorName().lower()
        try:
            operation = self.OPERATION_MAPPING[operator_name]
        except KeyError:
            try:
                operation = context.schema[schema_name].functions[operator_name]
            except KeyError:  # pragma: no cover
                raise NotImplementedError(f"RexCall operator '{operator_name}' not (yet) implemented")
        logger.debug(f"Executing {operator_name} on {[str(LoggableDataFrame(df)) for df in operands]}")
        kwargs = {}
        if Operation.op_needs_dc(operation):
            kwargs["dc"] = dc
        if Operation.op_needs_rex(operation):
            kwargs["rex"] = expr
        if Operation.op_needs_context(operation):
            kwargs["context"] = context
        if Operation.op_needs_rel(operation):
            kwargs["rel"] = rel
        return operation(*operands, **kwargs)
