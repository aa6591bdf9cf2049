"""`grid_sample` entry point kept for import compatibility (reference
`stylegan2ada/torch_utils/ops/grid_sample_gradfix.py`; only the ADA augmentation pipeline uses it and
`trainers.py:513` sets `.enabled`).  On torch >= 2.0 the reference's own custom op is inert
(grid_sample_gradfix.py:34-40) and it calls torch.nn.functional.grid_sample, which supports double
backward natively; this module does the same.  Augmentation is outside the op hot path."""
import torch

enabled = False


def grid_sample(input, grid):
    return torch.nn.functional.grid_sample(input=input, grid=grid, mode='bilinear', padding_mode='zeros', align_corners=False)
