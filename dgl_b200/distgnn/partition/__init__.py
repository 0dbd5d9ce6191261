"""Reference package path ``dgl.distgnn.partition`` (contents are lazy re-exports)."""
