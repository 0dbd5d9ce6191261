"""Feature storages (parity: /root/reference/python/dgl/storages/): ``fetch(indices, device)`` abstraction used by
the DataLoader prefetcher; pinned CPU tensors are gathered by the UVA kernel."""
import threading

import torch

from ..ops.index import index_select

STORAGE_WRAPPERS = {}


def register_storage_wrapper(type_):
    """Class decorator: objects of ``type_`` placed in a frame are wrapped by the decorated FeatureStorage class
    (reference storages/base.py:8)."""
    def deco(cls):
        STORAGE_WRAPPERS[type_] = cls
        return cls
    return deco


class ThreadedFuture:
    """Run ``target(*args)`` on a daemon thread right away; :meth:`wait` joins and returns the result, re-raising the
    worker's exception in the caller (reference storages/base.py:42)."""

    def __init__(self, target, args):
        self._result, self._exc = None, None

        def run():
            try:
                self._result = target(*args)
            except BaseException as exc:  # noqa: BLE001 - surfaced by wait()
                self._exc = exc

        self.thread = threading.Thread(target=run, daemon=True)
        self.thread.start()

    def wait(self):
        self.thread.join()
        if self._exc is not None:
            raise self._exc
        return self._result


class FeatureStorage:
    def requires_ddp(self):
        return False

    def fetch(self, indices, device, pin_memory=False, **kwargs):
        raise NotImplementedError


class BaseTensorStorage(FeatureStorage):
    """Synchronous slice + transfer (reference storages/tensor.py:6)."""

    def __init__(self, tensor):
        self.storage = tensor

    def fetch(self, indices, device, pin_memory=False, **kwargs):
        return self.storage[indices.to(self.storage.device).long()].to(device, **kwargs)


@register_storage_wrapper(torch.Tensor)
class TensorStorage(BaseTensorStorage):
    def __init__(self, tensor):
        super().__init__(tensor)
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


def _numpy_type():
    import numpy as np
    return np.ndarray


@register_storage_wrapper(_numpy_type())
class NumpyStorage(FeatureStorage):
    def __init__(self, arr):
        self.storage = arr

    def fetch(self, indices, device, pin_memory=False, **kwargs):
        return torch.from_numpy(self.storage[indices.cpu().numpy()]).to(device)


    def fetch_async(self, indices, device, pin_memory=False, **kwargs):
        """Memory-mapped arrays are slow to slice; do it off-thread."""
        return ThreadedFuture(self.fetch, (indices, device, pin_memory))


def wrap_storage(storage):
    """Object -> FeatureStorage through the registered wrappers; FeatureStorage instances pass through."""
    for type_, cls in STORAGE_WRAPPERS.items():
        if isinstance(storage, type_):
            return cls(storage)
    assert isinstance(storage, FeatureStorage), \
        f"The frame column must be a tensor or a FeatureStorage object, got {type(storage)}"
    return storage
