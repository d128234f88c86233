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

This is host orchestration around a user-supplied model: there is no kernel here and nothing for the
C ABI to replace; the feature rows themselves come from ``extract_audio_features``.
"""
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


def process_audio_features(audio_features, model, device, config, batched=True):
    """reference :50-112.  ``audio_features``: ``(num_frames, 256)`` rows of ``extract_audio_features``."""
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
