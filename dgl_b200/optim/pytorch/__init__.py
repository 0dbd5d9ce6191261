"""Reference package path ``dgl.optim.pytorch`` (contents are lazy re-exports)."""
