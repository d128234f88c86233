"""Mirror of the reference's ``utils/audio/processing/audio_processing.py`` - the inference-side consumer
of the feature rows: ``frame_size``-row chunks with ``overlap`` rows of
linear cross-fade between the decoded chunks, a reflect-padded last chunk, ``[:, :61] /= 100`` at the end.

Same names, arguments and results as the reference.  Two things differ in *how* the work is done:

* the chunk list is built once, every chunk is reflect-padded with one vectorised NumPy expression, and -
  with ``batched=True`` (default) - all chunks go through ``model.encoder`` / ``model.decoder`` as ONE
  batch: the chunks are independent, so the per-chunk Python loop of the reference (one kernel-launch
  storm and one device->host copy per chunk) collapses into a single forward pass and a single copy.
  ``batched=False`` decodes chunk by chunk exactly like the reference;
* the cross-fade is one broadcast expression per chunk boundary instead of a Python loop over rows
  (same float32 arithmetic: ``(1 - i/n) * a + (i/n) * b``).

On a CUDA device the chunker is device-resident: the feature rows - a CUDA tensor
straight from ``Engine.extract_device``, or host rows uploaded once - are cut into the ``[n_chunks, frame, 256]``
batch by ``nsf_chunk_gather`` (reflect-completed tail included), go through the model as one batch, and
``nsf_chunk_blend`` reassembles the decoded chunks with the reference's cross-fade arithmetic and the final
``[:, :61] /= 100``; one device-to-host copy at the end (none with ``return_tensor=True``).  With the model on the
CPU the NumPy route below runs, which chunk by chunk is bit-identical to the reference module.
"""
import ctypes as C

import numpy as np
import torch


def concatenate_outputs(all_decoded_outputs, num_frames):
    """reference :4-7."""
    return np.concatenate(all_decoded_outputs, axis=0)[:num_frames]


def ensure_2d(final_decoded_outputs):
    """reference :9-12."""
    if final_decoded_outputs.ndim == 3:
        final_decoded_outputs = final_decoded_outputs.reshape(-1, final_decoded_outputs.shape[-1])
    return final_decoded_outputs


def pad_audio_chunk(audio_chunk, frame_length, num_features):
    """reference :14-23 -- short chunks are extended with their own reflection (np.pad 'reflect')."""
    if audio_chunk.shape[0] < frame_length:
        pad_length = frame_length - audio_chunk.shape[0]
        padding = np.pad(audio_chunk, pad_width=((0, pad_length), (0, 0)), mode='reflect')
        audio_chunk = np.vstack((audio_chunk, padding[-pad_length:, :num_features]))
    return audio_chunk


def decode_audio_chunk(audio_chunk, model, device):
    """reference :25-31 -- one chunk through encoder + decoder."""
    src_tensor = torch.tensor(audio_chunk, dtype=torch.float32).unsqueeze(0).to(device)
    with torch.no_grad():
        encoder_outputs = model.encoder(src_tensor)
        output_sequence = model.decoder(encoder_outputs)
        decoded_outputs = output_sequence.squeeze(0).cpu().numpy()
    return decoded_outputs


def decode_audio_chunks(audio_chunks, model, device):
    """All (equal-length) chunks as one batch: ``[n, frame_size, features] -> [n, frame_size, out]``."""
    src = torch.as_tensor(np.stack(audio_chunks), dtype=torch.float32).to(device)
    with torch.no_grad():
        return model.decoder(model.encoder(src)).cpu().numpy()


def blend_chunks(chunk1, chunk2, overlap):
    """reference :33-48 -- linear cross-fade of the last / first ``overlap`` rows."""
    actual_overlap = min(overlap, len(chunk1), len(chunk2))
    if actual_overlap == 0:
        return np.vstack((chunk1, chunk2))
    blended_chunk = np.copy(chunk1)
    # row i of the overlap: (1 - i/n) * chunk1 + (i/n) * chunk2.  The reference multiplies array rows by
    # Python floats, which NumPy rounds to the ROW's dtype first (float32 rows stay float32): same here
    alpha = (np.arange(actual_overlap) / actual_overlap)[:, None]
    blended_chunk[-actual_overlap:] = ((1 - alpha).astype(chunk1.dtype) * chunk1[-actual_overlap:] +
                                       alpha.astype(chunk2.dtype) * chunk2[:actual_overlap])
    return np.vstack((blended_chunk, chunk2[actual_overlap:]))


def process_audio_features_device(audio_features, model, device, config, return_tensor=False, stream=None):
    """reference :50-112 with every row-wise step on the GPU (``device`` must be a CUDA device).

    ``audio_features``: ``(num_frames, features)`` float rows - a CUDA tensor (e.g. the output of
    ``Engine.extract_device``; not copied), a CPU tensor or a NumPy array (uploaded once).  Returns the
    ``(num_frames, out)`` float32 result as a NumPy array, or as a CUDA tensor with ``return_tensor=True``."""
    from .... import _native as nv
    from .... import engine as _engine
    dev = torch.device(device)
    if dev.type != "cuda":
        raise ValueError("process_audio_features_device needs a CUDA device")
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    dev = torch.device("cuda", index)
    frame_length = int(config['frame_size'])
    overlap = int(config.get('overlap', 16))
    rows = torch.as_tensor(audio_features)
    rows = rows.to(device=dev, dtype=torch.float32)
    if rows.dim() != 2 or rows.stride(1) != 1:
        rows = rows.contiguous()
    num_frames, num_features = rows.shape
    n_chunks = int(nv.lib.nsf_chunk_count(num_frames, frame_length, overlap))
    if n_chunks <= 0:
        raise ValueError("need at least one feature row and 0 <= overlap < frame_size")
    f_len, h_len = _engine.frame_params(88200)
    eng = _engine.get_engine(88200, f_len, h_len, device=index)       # any context of that device will do
    s = torch.cuda.current_stream(dev) if stream is None else stream
    model.eval()
    with torch.cuda.stream(s), torch.no_grad():
        chunks = torch.empty((n_chunks, frame_length, num_features), dtype=torch.float32, device=dev)
        nv.check(nv.lib.nsf_chunk_gather(eng.handle, C.c_void_p(s.cuda_stream), C.c_void_p(rows.data_ptr()), num_frames,
                                         num_features, rows.stride(0), frame_length, overlap,
                                         C.c_void_p(chunks.data_ptr())))
        decoded = model.decoder(model.encoder(chunks)).to(torch.float32).contiguous()
        out_cols = decoded.shape[-1]
        out = torch.empty((num_frames, out_cols), dtype=torch.float32, device=dev)
        nv.check(nv.lib.nsf_chunk_blend(eng.handle, C.c_void_p(s.cuda_stream), C.c_void_p(decoded.data_ptr()), num_frames,
                                        out_cols, frame_length, overlap, min(61, out_cols), C.c_float(100.0),
                                        C.c_void_p(out.data_ptr())))
    if return_tensor:
        return out
    s.synchronize()
    return out.cpu().numpy()


def process_audio_features(audio_features, model, device, config, batched=True):
    """reference :50-112.  ``audio_features``: ``(num_frames, 256)`` rows of ``extract_audio_features``.

    On a CUDA ``device`` (and ``batched=True``) the device-resident route above runs; otherwise the host route
    below, which with ``batched=False`` decodes chunk by chunk exactly like the reference."""
    if batched and torch.device(device).type == "cuda":
        return process_audio_features_device(audio_features, model, device, config)
    if isinstance(audio_features, torch.Tensor):
        audio_features = audio_features.detach().cpu().numpy()
    frame_length = config['frame_size']
    overlap = config.get('overlap', 16)
    num_features = audio_features.shape[1]
    num_frames = audio_features.shape[0]
    model.eval()

    # chunk starts exactly as the reference's while loop produces them
    starts, start_idx = [], 0
    while start_idx < num_frames:
        starts.append(start_idx)
        start_idx += frame_length - overlap
    ends = [min(s + frame_length, num_frames) for s in starts]
    chunks = [pad_audio_chunk(audio_features[s:e], frame_length, num_features) for s, e in zip(starts, ends)]
    if batched and chunks:
        decoded = decode_audio_chunks(chunks, model, device)
        decoded = [decoded[i][:e - s] for i, (s, e) in enumerate(zip(starts, ends))]
    else:
        decoded = [decode_audio_chunk(c, model, device)[:e - s] for c, s, e in zip(chunks, starts, ends)]

    all_decoded_outputs = []
    for out in decoded:
        if all_decoded_outputs:
            all_decoded_outputs.append(blend_chunks(all_decoded_outputs.pop(), out, overlap))
        else:
            all_decoded_outputs.append(out)

    current_length = sum(len(chunk) for chunk in all_decoded_outputs)
    if current_length < num_frames:                                   # reference :86-94
        remaining_frames = num_frames - current_length
        audio_chunk = pad_audio_chunk(audio_features[num_frames - remaining_frames:num_frames], frame_length,
                                      num_features)
        all_decoded_outputs.append(decode_audio_chunk(audio_chunk, model, device)[:remaining_frames])

    final_decoded_outputs = np.concatenate(all_decoded_outputs, axis=0)[:num_frames]
    final_decoded_outputs = ensure_2d(final_decoded_outputs)
    final_decoded_outputs[:, :61] /= 100
    return final_decoded_outputs


def zero_columns(data):
    """reference :114-119."""
    columns_to_zero = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60]
    modified_data = np.copy(data)
    modified_data[:, columns_to_zero] = 0
    return modified_data


def add_specified_dimensions_back(modified_data):
    """reference :123-141 -- scatter the kept columns back into the 68-wide layout."""
    original_dim = 68
    columns_to_remove = [0, 1, 2, 3, 4, 7, 8, 9, 10, 11, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60]
    new_data = np.zeros((modified_data.shape[0], original_dim))
    remaining_cols = [c for c in range(original_dim) if c not in columns_to_remove]
    new_data[:, remaining_cols] = modified_data
    return new_data
