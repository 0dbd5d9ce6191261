"""Feature storages (parity: /root/reference/python/dgl/storages/): ``fetch(indices, device)`` abstraction used by
the DataLoader prefetcher; pinned CPU tensors are gathered by the UVA kernel."""
import torch

from ..ops.index import index_select


class FeatureStorage:
    def requires_ddp(self):
        return False

    def fetch(self, indices, device, pin_memory=False, **kwargs):
        raise NotImplementedError


class TensorStorage(FeatureStorage):
    def __init__(self, tensor):
        self.storage = tensor
        self.feature_shape = tensor.shape[1:]

    def fetch(self, indices, device, pin_memory=False, **kwargs):
        device = torch.device(device)
        if self.storage.is_cuda or (self.storage.is_pinned() and indices.is_cuda):
            return index_select(self.storage, indices.to("cuda") if not indices.is_cuda else indices).to(device)
        out = self.storage[indices.cpu().long()]
        if pin_memory and device.type == "cuda":
            out = out.pin_memory()
        return out.to(device, non_blocking=True)


PyTorchTensorStorage = TensorStorage


class NumpyStorage(FeatureStorage):
    def __init__(self, arr):
        self.storage = arr

    def fetch(self, indices, device, pin_memory=False, **kwargs):
        return torch.from_numpy(self.storage[indices.cpu().numpy()]).to(device)


def wrap_storage(storage):
    if torch.is_tensor(storage):
        return TensorStorage(storage)
    import numpy as np
    if isinstance(storage, np.ndarray):
        return NumpyStorage(storage)
    return storage
