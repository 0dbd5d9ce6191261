"""Reference package path ``dgl.distgnn.tools`` (contents are lazy re-exports)."""
