"""Datapipe stages re-exported under one roof: the classes live in the modules the reference names
(``graphbolt/minibatch_transformer.py``, ``subgraph_sampler.py``, ``feature_fetcher.py``, ``dataloader.py``)."""
from ..dataloader import CopyTo, DataLoader, _record_stream  # noqa: F401
from ..feature_fetcher import FeatureFetcher  # noqa: F401
from ..minibatch_transformer import MiniBatchTransformer, _Stage  # noqa: F401
from ..subgraph_sampler import SubgraphSampler  # noqa: F401
