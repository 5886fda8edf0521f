"""ORACLE — test infrastructure only (see pandas_oracle.py)."""
