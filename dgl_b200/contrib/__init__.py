"""``dgl.contrib`` (reference python/dgl/contrib): only the unified-tensor remnant is still referenced by user code;
UVA access lives in ``dgl_b200.utils.pin_memory`` / ``dgl_b200.ops.cuda_struct``."""
from ..utils import gather_pinned_tensor_rows, pin_memory_inplace  # noqa: F401


class UnifiedTensor:
    """Host-pinned tensor indexable from the GPU (reference contrib/unified_tensor.py, deprecated there in favour of
    ``pin_memory_inplace`` + ``gather_pinned_tensor_rows``)."""

    def __init__(self, input, device):  # noqa: A002
        import torch
        self._input = input if input.is_pinned() else input.pin_memory()
        self._device = torch.device(device)

    def __len__(self):
        return len(self._input)

    def __getitem__(self, key):
        import torch
        if torch.is_tensor(key) and key.is_cuda:
            return gather_pinned_tensor_rows(self._input, key)
        return self._input[key]

    shape = property(lambda self: self._input.shape)
    dtype = property(lambda self: self._input.dtype)
    device = property(lambda self: self._device)
