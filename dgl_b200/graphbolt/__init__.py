"""GraphBolt: the data-loading stack (parity: /root/reference/python/dgl/graphbolt/__init__.py + graphbolt/ C++)."""
from .base import (CANONICAL_ETYPE_DELIMITER, ORIGINAL_EDGE_ID, CSCFormatBase, etype_str_to_tuple, etype_tuple_to_str,  # noqa: F401
                   expand_indptr, index_select, isin, seed_type_str_to_ntypes)
from .datapipes import CopyTo, DataLoader, FeatureFetcher, MiniBatchTransformer, SubgraphSampler  # noqa: F401
from .impl import *  # noqa: F401,F403
from .impl.fused_csc_sampling_graph import load_from_shared_memory  # noqa: F401
from .itemset import DistributedItemSampler, HeteroItemSet, ItemSampler, ItemSet, ItemSetDict  # noqa: F401
from .minibatch import MiniBatch, SampledSubgraph, SampledSubgraphImpl  # noqa: F401
from .utils import (add_reverse_edges, compact_csc_format, exclude_seed_edges, unique_and_compact,  # noqa: F401
                    unique_and_compact_csc_formats)
from .impl.cooperative import rank_assignment, rank_sort  # noqa: F401


def set_seed(seed):
    from .. import random as _r
    _r.seed(seed)


def numpy_save_aligned(path, arr, alignment=4096):
    """``numpy.save`` whose data section starts at a multiple of ``alignment`` bytes, so DiskBasedFeature can read rows
    with O_DIRECT-style aligned preads (reference graphbolt/internal_utils.py:numpy_save_aligned)."""
    import numpy as np
    from numpy.lib import format as npf
    arr = np.ascontiguousarray(arr)
    header = npf.header_data_from_array_1_0(arr)
    import io
    buf = io.BytesIO()
    npf.write_array_header_2_0(buf, header)
    hdr = buf.getvalue()
    pad = (-len(hdr)) % alignment
    if pad:  # grow the (space padded) header dict so that the data offset is aligned; fix the length field
        hlen = int.from_bytes(hdr[8:12], "little") + pad
        hdr = hdr[:8] + hlen.to_bytes(4, "little") + hdr[12:-1] + b" " * pad + b"\n"
    with open(path, "wb") as f:
        f.write(hdr)
        arr.tofile(f)

from .compat import *  # noqa: F401,F403,E402
from .compat import (FeatureKey, GBWarning, all_to_all, built_with_cuda, bytes_to_number_of_items, check_sha1,  # noqa: F401,E402
                     convert_to_hetero, download, extract_archive, gb_warning, gb_warning_format, get_attributes,
                     get_device_to_host_uva_stream, get_host_to_device_uva_stream, get_nonproperty_attributes,
                     indptr_edge_ids, is_listlike, is_wsl, minibatcher_default, preprocess_ondisk_dataset, recursive_apply,
                     recursive_apply_reduce_all, revert_to_homo, seed, wrap_with_cached_feature)
from .impl.negative_sampler import NegativeSampler  # noqa: F401,E402

from .datapipes.utils import datapipe_graph_to_adjlist, find_dps, replace_dp, traverse_dps  # noqa: F401,E402
from .datapipes.visualization import to_graph  # noqa: F401,E402
