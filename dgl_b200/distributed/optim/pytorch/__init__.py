"""Reference package path ``dgl.distributed.optim.pytorch`` (contents are lazy re-exports)."""
