"""numpy restatement of the reference's sliding-window prediction — TEST INFRASTRUCTURE (see oracle/__init__).
Pinned against the reference's own ``SliceBuilder`` / ``mirror_pad`` / ``remove_padding``
(dataset/unet3d_dataset/utils.py) by tests/golden/predictor_slices.npz (made by oracle/make_golden.py)."""
from __future__ import annotations

import numpy as np


def gen_indices(i, k, s):
    """utils.py:119-125."""
    assert i >= k, "Sample size has to be bigger than the patch size"
    j = 0
    for j in range(0, i - k + 1, s):
        yield j
    if j + k < i:
        yield i - k


def build_slices(shape, patch_shape, stride_shape):
    """utils.py:85-117 for a (D, H, W) array."""
    out = []
    for z in gen_indices(shape[0], patch_shape[0], stride_shape[0]):
        for y in gen_indices(shape[1], patch_shape[1], stride_shape[1]):
            for x in gen_indices(shape[2], patch_shape[2], stride_shape[2]):
                out.append((slice(z, z + patch_shape[0]), slice(y, y + patch_shape[1]), slice(x, x + patch_shape[2])))
    return out


def mirror_pad(image, padding_shape):
    """utils.py:314-342."""
    if all(p == 0 for p in padding_shape):
        return image
    pad_width = [(p, p) for p in padding_shape]
    if image.ndim == 4:
        pad_width = [(0, 0)] + pad_width
    return np.pad(image, pad_width, mode="reflect")


def remove_padding(m, padding_shape):
    """utils.py:345-361."""
    if padding_shape is None:
        return m
    return m[(..., *(slice(p, -p or None) for p in padding_shape))]


def standard_predict(model_fn, raw, patch_shape, stride_shape, halo_shape=(0, 0, 0), out_channels=1,
                     prediction_channel=None, save_segmentation=False, batch_size=1):
    """predictor.py:86-165 with the HDF5 dataset's patch extraction (hdf5.py:158-176: slices of the mirror-padded
    raw widened by 2 * halo).  ``model_fn`` maps a float32 (B, C, d, h, w) array to (B, out_channels, d, h, w)."""
    raw = np.asarray(raw, dtype=np.float32)
    if raw.ndim == 3:
        raw = raw[None]
    padded = mirror_pad(raw, halo_shape)
    channels = 1 if prediction_channel is not None else out_channels
    prediction_map = np.zeros((channels,) + raw.shape[1:], dtype="float32")
    normalization_mask = np.zeros((channels,) + raw.shape[1:], dtype="uint8")
    slices = build_slices(raw.shape[1:], patch_shape, stride_shape)
    for start in range(0, len(slices), batch_size):
        chunk = slices[start:start + batch_size]
        batch = np.stack([padded[(slice(None),) + tuple(slice(s.start, s.stop + 2 * h) for s, h in zip(idx, halo_shape))]
                          for idx in chunk])
        prediction = remove_padding(model_fn(batch), halo_shape)
        for pred, index in zip(prediction, chunk):
            if prediction_channel is None:
                channel_slice = slice(0, out_channels)
            else:
                channel_slice = slice(0, 1)
                pred = np.expand_dims(pred[prediction_channel], axis=0)
            index = (channel_slice,) + tuple(index)
            prediction_map[index] += pred
            normalization_mask[index] += 1
    result = prediction_map / normalization_mask
    if save_segmentation:
        result = np.argmax(result, axis=0).astype("uint16")
    return result


def lazy_predict(model_fn, raw, patch_shape, stride_shape, halo_shape=(0, 0, 0), out_channels=1,
                 prediction_channel=None, save_segmentation=False, batch_size=1):
    """predictor.py:171-223 (``LazyPredictor``): the same accumulation as ``standard_predict`` into datasets of the
    output file (arrays here — h5py is a storage detail), then ``_save_results``: block-wise normalisation over
    slices of 1/27 of the volume (``SliceBuilder._build_slices`` with patch = stride = shape // 3), the visited
    block's ``normalization_mask`` reset to 1 so that overlapping last blocks are not normalised twice, and the
    arg-max segmentation written block by block."""
    raw = np.asarray(raw, dtype=np.float32)
    if raw.ndim == 3:
        raw = raw[None]
    padded = mirror_pad(raw, halo_shape)
    channels = 1 if prediction_channel is not None else out_channels
    prediction_map = np.zeros((channels,) + raw.shape[1:], dtype="float32")
    normalization_mask = np.zeros((channels,) + raw.shape[1:], dtype="uint8")
    slices = build_slices(raw.shape[1:], patch_shape, stride_shape)
    for start in range(0, len(slices), batch_size):
        chunk = slices[start:start + batch_size]
        batch = np.stack([padded[(slice(None),) + tuple(slice(s.start, s.stop + 2 * h) for s, h in zip(idx, halo_shape))]
                          for idx in chunk])
        prediction = remove_padding(model_fn(batch), halo_shape)
        for pred, index in zip(prediction, chunk):
            if prediction_channel is None:
                channel_slice = slice(0, out_channels)
            else:
                channel_slice = slice(0, 1)
                pred = np.expand_dims(pred[prediction_channel], axis=0)
            index = (channel_slice,) + tuple(index)
            prediction_map[index] += pred
            normalization_mask[index] += 1
    z, y, x = prediction_map.shape[1:]
    block = (z // 3, y // 3, x // 3)
    segmentation = np.zeros((z, y, x), dtype="uint16") if save_segmentation else None
    for idx in build_slices((z, y, x), block, block):
        index = (slice(0, channels),) + tuple(idx)
        prediction_map[index] /= normalization_mask[index]
        normalization_mask[index] = 1
        if save_segmentation:
            segmentation[tuple(idx)] = np.argmax(prediction_map[index], axis=0).astype("uint16")
    return segmentation if save_segmentation else prediction_map
