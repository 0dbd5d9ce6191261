"""Reference package path ``dgl.distributed.nn.pytorch`` (contents are lazy re-exports)."""
