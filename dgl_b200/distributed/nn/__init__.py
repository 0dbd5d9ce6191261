"""Reference package path ``dgl.distributed.nn`` (contents are lazy re-exports)."""
