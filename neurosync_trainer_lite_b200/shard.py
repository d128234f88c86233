"""Sharding of the feature path by clip over the GPUs of one box (one process per GPU).

The path has no cross-clip dependency - peak, dB max, CMVN statistics, edge fix, length matching
and blending are all per clip - so a dataset is split into independent
shards, each rank extracts its own clips with its own CUDA context, and the rows meet again on the
HOST: there is no NCCL (or any) collective on the data path.  ``torch.distributed`` appears only as
process plumbing (``gather_rows`` moves the *host* arrays to rank 0 over gloo when the caller wants
one array; ``SharedRows`` avoids even that by letting every rank write its slice of one /dev/shm
array).

Order contract: the reference builds its example list in ``os.listdir`` order
(``dataset/data_processing.py:16``); ``assemble`` restores exactly the input order.
"""
import os

import numpy as np


def lpt_partition(lengths, world):
    """Greedy longest-processing-time partition of clips by sample count.

    Returns ``world`` ascending index lists.  Deterministic: ties are broken by clip index, and
    equal-length clips degenerate to round-robin."""
    lengths = [int(n) for n in lengths]
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    load = [0] * world
    parts = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], len(parts[k]), k))
        parts[r].append(i)
        load[r] += lengths[i]
    return [sorted(p) for p in parts]


def block_partition(lengths, world):
    """Contiguous split of the clip list into ``world`` runs of near-equal total length (each cut placed at the
    clip boundary closest to its ideal position).  Every rank's rows then form ONE slice of the gathered array,
    so a rank can download straight into it."""
    lengths = np.asarray([int(n) for n in lengths], dtype=np.int64)
    n = len(lengths)
    cum = np.concatenate([[0], np.cumsum(lengths)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        i = int(np.searchsorted(cum, target))            # first boundary at or beyond the target
        if i > 0 and target - cum[i - 1] <= cum[min(i, n)] - target:
            i -= 1
        cuts.append(min(max(i, cuts[-1]), n))
    cuts.append(n)
    return [list(range(cuts[r], cuts[r + 1])) for r in range(world)]


def extract_collect_multi_device(pcm, offsets, facial, facial_offsets, devices, sr=88200, flags=0, include_fast=True,
                                 include_slow=False, blend_boundaries=True, blend_frames=30, out_audio=None,
                                 out_facial=None):
    """Features + ``collect_features`` augmentation of one dataset on SEVERAL GPUs from ONE process: one thread and
    one library context per device, every device working on a contiguous run of clips and downloading its rows
    straight into its slice of ONE page-locked host array ("gather" = independent D2H
    copies, no collective).  ``pcm`` / ``facial`` should be page-locked (``engine.PinnedBuffer``) for full speed.
    Returns ``(audio rows, facial rows, output row offsets)`` in the input clip order."""
    import threading

    from . import engine as _engine
    f_len, h_len = _engine.frame_params(sr)
    off = np.asarray(offsets, dtype=np.int64)
    f_off = np.asarray(facial_offsets, dtype=np.int64)
    engines = [_engine.get_engine(sr, f_len, h_len, device=d) for d in devices]
    kw = dict(include_fast=include_fast, include_slow=include_slow, blend_boundaries=blend_boundaries, blend_frames=blend_frames)
    o_off = engines[0].collect_rows(engines[0].row_offsets(off, flags), f_off, **kw)
    cols = engines[0].plan.feature_cols(flags)
    n_out = int(o_off[-1])
    if out_audio is None:
        out_audio = _engine.PinnedBuffer(n_out * cols * 4).view(np.float32, (n_out, cols))
    if out_facial is None:
        out_facial = _engine.PinnedBuffer(n_out * facial.shape[1] * 4).view(np.float32, (n_out, facial.shape[1]))
    parts = block_partition(np.diff(off), len(devices))
    errors = []

    def work(eng, idx):
        try:
            if not idx:
                return
            c0, c1 = idx[0], idx[-1] + 1
            eng.extract_collect_host(pcm, off[c0:c1 + 1], facial, f_off[c0:c1 + 1], flags,
                                     out_audio=out_audio[o_off[c0]:o_off[c1]], out_facial=out_facial[o_off[c0]:o_off[c1]], **kw)
        except Exception as e:  # noqa: BLE001 - re-raised in the caller's thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(e, p)) for e, p in zip(engines, parts)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return out_audio, out_facial, o_off


def row_layout(row_counts, parts):
    """Global row offsets (input order) and, per rank, the global row offset of each of its clips."""
    row_counts = np.asarray(row_counts, dtype=np.int64)
    offsets = np.zeros(len(row_counts) + 1, dtype=np.int64)
    np.cumsum(row_counts, out=offsets[1:])
    return offsets, [[int(offsets[i]) for i in p] for p in parts]


def default_extract(sr, frame_length=None, hop_length=None, flags=0, device=None):
    """The real extractor: one ``nsf_extract_host`` call on this rank's GPU."""
    from . import engine as _engine

    def run(clips):
        f, h = (frame_length, hop_length) if frame_length else _engine.frame_params(sr)
        eng = _engine.get_engine(sr, f, h, device=device)
        packed, off = _engine.pack_clips(clips)
        rows = eng.extract_host(packed, off, flags)
        return rows, eng.row_offsets(off, flags)
    return run


def extract_shard(clips, rank, world, extract_fn):
    """Extract this rank's share of ``clips``.  Returns ``(indices, rows, local_row_offsets)``."""
    parts = lpt_partition([len(c) for c in clips], world)
    mine = parts[rank]
    if not mine:
        return mine, np.zeros((0, 0), dtype=np.float32), np.zeros(1, dtype=np.int64)
    rows, roff = extract_fn([clips[i] for i in mine])
    return mine, rows, np.asarray(roff, dtype=np.int64)


def assemble(n_clips, pieces):
    """``pieces``: iterable of ``(indices, rows, local_row_offsets)`` from every rank.
    Returns ``(rows_in_input_order, global_row_offsets)``."""
    counts = np.zeros(n_clips, dtype=np.int64)
    cols, dtype = 0, np.float32
    for idx, rows, roff in pieces:
        for k, i in enumerate(idx):
            counts[i] = roff[k + 1] - roff[k]
        if len(idx):
            cols, dtype = rows.shape[1], rows.dtype
    offsets = np.zeros(n_clips + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    out = np.empty((int(offsets[-1]), cols), dtype=dtype)
    for idx, rows, roff in pieces:
        for k, i in enumerate(idx):
            out[offsets[i]:offsets[i + 1]] = rows[roff[k]:roff[k + 1]]
    return out, offsets


def gather_rows(n_clips, piece, group=None, dst=0):
    """Host-side gather of every rank's piece to ``dst`` over the process group (gloo): returns
    ``(rows, offsets)`` on ``dst`` and ``(None, None)`` elsewhere.  Data never touches NCCL."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(piece, bucket, dst=dst, group=group)
    if rank != dst:
        return None, None
    return assemble(n_clips, bucket)


class SharedRows:
    """One ``(rows, cols)`` float32 array in /dev/shm that every rank of the box maps and fills at
    its clips' global row offsets - the 'gather to host' of BASELINE config 4 without any copy
    between processes.  Rank 0 creates, everybody opens after a barrier, rank 0 unlinks."""

    def __init__(self, name, rows, cols, create):
        self.path = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp", name)
        self.shape = (int(rows), int(cols))
        mode = "w+" if create else "r+"
        self.array = np.memmap(self.path, dtype=np.float32, mode=mode, shape=self.shape)
        self.owner = create

    def write(self, indices, rows, local_row_offsets, global_offsets):
        for k, i in enumerate(indices):
            n = int(local_row_offsets[k + 1] - local_row_offsets[k])
            self.array[global_offsets[i]:global_offsets[i] + n] = rows[local_row_offsets[k]:local_row_offsets[k + 1]]
        self.array.flush()

    def close(self):
        del self.array
        if self.owner and os.path.exists(self.path):
            os.unlink(self.path)
