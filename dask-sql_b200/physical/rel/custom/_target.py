"""Name resolution and the IF [NOT] EXISTS / OR REPLACE policy shared by the DDL plugins."""


def split_qualified(context, qualified_name):
    """'table' | 'schema.table' -> (schema, table); anything longer is an error
    (create_memory_table.py:42-50, drop_table.py:28-36)."""
    parts = qualified_name.split(".")
    if len(parts) > 2:
        raise RuntimeError(f"Expected unqualified or fully qualified table name, got {qualified_name}.")
    schema = parts[0] if len(parts) == 2 else context.schema_name
    return schema, parts[-1]


def table_exists(context, schema, table):
    return schema in context.schema and table.lower() in context.schema[schema].tables


def may_create(context, schema, table, if_not_exists, or_replace):
    """True: go ahead.  False: the table exists and IF NOT EXISTS says keep it.  Raises when it exists
    and neither clause was given (create_table.py:47-53, create_memory_table.py:55-61)."""
    if schema not in context.schema:
        raise RuntimeError(f"A schema with the name {schema} is not present.")
    if not table_exists(context, schema, table):
        return True
    if if_not_exists:
        return False
    if or_replace:
        return True
    raise RuntimeError(f"A table with the name {table} is already present.")
