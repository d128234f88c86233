"""Host-side windowing of dataset/dataset.py: items are independent tensors by default."""
import numpy as np
import torch


def _dataset(zero_copy):
    from neurosync_trainer_lite_b200.dataset.dataset import AudioFacialDataset
    ds = AudioFacialDataset.__new__(AudioFacialDataset)
    ds.micro_batch_size = 128
    ds.zero_copy_windows = zero_copy
    ds.clips, ds._starts = [], [0]
    ds.add_clip(np.arange(300 * 4, dtype=np.float64).reshape(300, 4), np.ones((300, 3)))
    return ds


def test_items_are_independent_copies_by_default():
    ds = _dataset(False)
    a0, _ = ds[0]
    before = ds[1][0].clone()
    a0.mul_(0)                                   # an in-place augmentation on one item ...
    assert torch.equal(ds[1][0], before)         # ... must not leak into the 127 windows overlapping it
    assert len(ds) == 174                        # reference count for N = 300 (173 regular + duplicated tail)
    assert torch.equal(ds[172][0], ds[173][0])


def test_zero_copy_windows_share_storage_when_asked():
    ds = _dataset(True)
    a0, _ = ds[0]
    a1, _ = ds[1]
    assert a0.data_ptr() + a0.stride(0) * a0.element_size() == a1.data_ptr()
