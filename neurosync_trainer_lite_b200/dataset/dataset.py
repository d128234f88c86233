"""Mirror of the reference's ``dataset/dataset.py`` - the consumer of the feature rows.

``AudioFacialDataset`` keeps the reference's constructor, ``__len__``, ``__getitem__`` and
``collate_fn`` contracts: item ``i`` is ``(float32[128, 256], float32[128, 61])``.  The difference
is memory: the reference materialises every stride-1 window (``process_example`` :58-98, the source
of its 128-256 GB host-RAM advice); here each clip's rows are stored once as float32 and a window is
produced at ``__getitem__`` time: an independent tensor by default (what the reference hands out, safe
for in-place augmentation), or - ``config['zero_copy_windows'] = True`` - a read-only-by-contract
``narrow`` view that shares the clip's storage with the 127 windows overlapping it.  The duplicated, reflected last window
the reference appends when ``N % 128 != 0`` (:77-96) and its failure for ``N < 128`` are preserved.
Pure host-side indexing - no arithmetic - so it needs no kernel.
"""
import numpy as np
import torch
from torch.nn.utils.rnn import pad_sequence
from torch.utils.data import DataLoader, Dataset, random_split

from .data_processing import load_data


def prepare_dataloader_with_split(config, val_split=0.1):
    """reference :12-21."""
    dataset = AudioFacialDataset(config)
    val_size = int(len(dataset) * val_split)
    train_dataset, val_dataset = random_split(dataset, [len(dataset) - val_size, val_size])
    train_dataloader = DataLoader(train_dataset, batch_size=config['batch_size'], shuffle=True,
                                  collate_fn=AudioFacialDataset.collate_fn)
    val_dataloader = DataLoader(val_dataset, batch_size=config['batch_size'], shuffle=False,
                                collate_fn=AudioFacialDataset.collate_fn)
    return train_dataset, val_dataset, train_dataloader, val_dataloader


def prepare_dataloader(config):
    """reference :23-26."""
    dataset = AudioFacialDataset(config)
    return dataset, DataLoader(dataset, batch_size=config['batch_size'], shuffle=True,
                               collate_fn=AudioFacialDataset.collate_fn)


class _WindowedClip:
    """Rows of one clip + the window list ``process_example`` would have materialised."""

    def __init__(self, audio_features, facial_data, window):
        na, nf = len(audio_features), len(facial_data)
        top = max(na, nf)
        if top < window:
            # the reference crashes here too (broadcast error in the tail branch, dataset.py:77-91)
            raise ValueError(f"clip has {top} rows, fewer than micro_batch_size={window}")
        self.window = window
        self.n_audio, self.n_facial = na, nf
        # float32 once per clip (the reference casts every window separately, :75)
        self.audio = torch.from_numpy(np.ascontiguousarray(audio_features, dtype=np.float32))
        self.facial = torch.from_numpy(np.ascontiguousarray(facial_data, dtype=np.float32))
        self.n_regular = top - window + 1
        self.has_tail = top % window != 0
        self.top = top

    def __len__(self):
        return self.n_regular + (1 if self.has_tail else 0)

    def _window(self, rows, n_rows, start):
        avail = min(self.window, n_rows - start)
        if avail == self.window:
            return rows.narrow(0, start, self.window)            # zero-copy view
        out = torch.zeros((self.window, rows.shape[1]), dtype=torch.float32)
        if avail > 0:
            out[:avail] = rows[start:start + avail]
        return out

    def get(self, i):
        # the tail window (:77-96) starts at top - window: a full window whenever both streams have
        # `top` rows (always true after collect_features), i.e. a duplicate of the last regular one
        if i < self.n_regular:
            return (self._window(self.audio, self.n_audio, i),
                    self._window(self.facial, self.n_facial, i))
        start = self.top - self.window
        return (self._tail(self.audio, self.n_audio, start),
                self._tail(self.facial, self.n_facial, start))

    def _tail(self, rows, n_rows, start):
        seg = rows[start:min(self.top, n_rows)]
        if len(seg) == self.window:
            return rows.narrow(0, start, self.window)
        out = torch.zeros((self.window, rows.shape[1]), dtype=torch.float32)
        out[:len(seg)] = seg
        fill = torch.flip(seg, dims=(0,))[:self.window - len(seg)]   # reflection fill (:87-94)
        out[len(seg):] = fill                                        # raises like the reference if short
        return out


class AudioFacialDataset(Dataset):
    def __init__(self, config):
        self.root_dir = config['root_dir']
        self.sr = config['sr']
        self.frame_rate = config['frame_rate']
        self.micro_batch_size = config['micro_batch_size']
        # False (default): items are independent copies, as in the reference.  True: items are views into the
        # clip's rows - no copy, but modifying one in place corrupts every overlapping window
        self.zero_copy_windows = bool(config.get('zero_copy_windows', False))
        self.processed_folders = set()
        self.clips = []
        self._starts = [0]
        for audio_features, facial_data in load_data(self.root_dir, self.sr, self.processed_folders):
            self.add_clip(audio_features, facial_data)

    def add_clip(self, audio_features, facial_data):
        clip = _WindowedClip(audio_features, facial_data, self.micro_batch_size)
        self.clips.append(clip)
        self._starts.append(self._starts[-1] + len(clip))

    @property
    def examples(self):
        """The reference's materialised list, produced lazily (for code that iterates it)."""
        return [self[i] for i in range(len(self))]

    def __len__(self):
        return self._starts[-1]

    def __getitem__(self, idx):
        if idx < 0:
            idx += len(self)
        if not 0 <= idx < len(self):
            raise IndexError(idx)
        c = int(np.searchsorted(self._starts, idx, side="right")) - 1
        a, f = self.clips[c].get(idx - self._starts[c])
        if getattr(self, "zero_copy_windows", False):
            return a, f
        return a.clone(), f.clone()

    @staticmethod
    def collate_fn(batch):
        """reference :52-56."""
        src_batch, trg_batch = zip(*batch)
        return (pad_sequence(src_batch, batch_first=True, padding_value=0),
                pad_sequence(trg_batch, batch_first=True, padding_value=0))

    def process_example(self, audio_features, facial_data):
        """reference :58-98 -- kept for API parity: the explicit list of windows of ONE clip."""
        clip = _WindowedClip(audio_features, facial_data, self.micro_batch_size)
        return [tuple(t.clone() for t in clip.get(i)) for i in range(len(clip))]
