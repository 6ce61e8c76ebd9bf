"""Block padding / splitting / merging with the reference's interface
(``engines/block_processor.py:7-48``).  Pure data movement: inside the round trip it does not
exist as a step (the kernels index the reflected position directly); these host-side NumPy
helpers are for callers that use the block list on its own."""

from typing import List, Tuple

import numpy as np


def pad_to_multiple(channel: np.ndarray, block_size: int) -> Tuple[np.ndarray, Tuple[int, int]]:
    """Reflect-pad the bottom / right edges to a multiple of ``block_size``
    (block_processor.py:7-16); returns (padded, (h, w))."""
    h, w = channel.shape
    pad_h = (block_size - h % block_size) % block_size
    pad_w = (block_size - w % block_size) % block_size
    if pad_h or pad_w:
        return np.pad(channel, ((0, pad_h), (0, pad_w)), mode='reflect'), (h, w)
    return channel.copy(), (h, w)


def split_into_blocks(channel: np.ndarray, block_size: int) -> List[Tuple[int, int, np.ndarray]]:
    """Raster list of (row, col, block copy) (block_processor.py:19-33); ragged edge blocks are
    zero-filled like the reference's."""
    h, w = channel.shape
    nby, nbx = -(-h // block_size), -(-w // block_size)
    full = np.zeros((nby * block_size, nbx * block_size), dtype=channel.dtype)
    full[:h, :w] = channel
    tiles = full.reshape(nby, block_size, nbx, block_size).swapaxes(1, 2)
    return [(i * block_size, j * block_size, tiles[i, j].copy())
            for i in range(nby) for j in range(nbx)]


def merge_blocks(blocks: List[Tuple[int, int, np.ndarray]], shape: Tuple[int, int],
                 block_size: int) -> np.ndarray:
    """Scatter blocks back into an fp64 channel of ``shape`` (block_processor.py:36-48)."""
    h, w = shape
    result = np.zeros((h, w), dtype=np.float64)
    for i, j, block in blocks:
        end_i, end_j = min(i + block_size, h), min(j + block_size, w)
        result[i:end_i, j:end_j] = block[:end_i - i, :end_j - j]
    return result
