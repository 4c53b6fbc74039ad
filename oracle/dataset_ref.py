"""TEST INFRASTRUCTURE (CPU oracle) - never imported by the product.

The training branch of SubjectLoader.fetch_data, datasets/dnerf_3d_video_IS.py:401-497, restated with the random draws
passed in, so that the GPU path and this restatement consume identical numbers.

Pinned: `multinomial_without_replacement` is checked in tests/test_oracle_golden.py against torch.multinomial itself
under the same generator state (torch's implementation - exponential_(1), div, topk - runs on the CPU in this
container); the rest of the function is the reference's own index arithmetic, line by line."""
import torch


def multinomial_without_replacement(weights, k, noise):
    """torch.multinomial(weights, k) [replacement=False] given its Exp(1) draws: the k largest of weights / noise
    (ATen multinomial_out: q = empty_like(w).exponential_(1); q = w / q; topk(q, k))."""
    return torch.topk(weights / noise, k).indices


def fetch_data_train(images, camtoworlds, K, timestamps, sampling_weights, weights_subsampled, num_rays, width, height,
                     opengl, subset, noise):
    """dnerf_3d_video_IS.py:404-497 with `subset` (the randint of :409-411, or None) and `noise` (the exponential draws
    inside torch.multinomial) given.  Returns the reference's dict plus the drawn cell indices."""
    batch_size = num_rays // (weights_subsampled ** 2)
    if subset is not None:
        samples = multinomial_without_replacement(sampling_weights[subset], batch_size, noise)
        index = subset[samples]
    else:
        index = multinomial_without_replacement(sampling_weights, batch_size, noise)
    cells = index
    hsub, wsub = height // weights_subsampled, width // weights_subsampled
    image_id = torch.div(index, hsub * wsub, rounding_mode="floor")
    ysub = torch.remainder(index, hsub * wsub).div(wsub, rounding_mode="floor")
    xsub = torch.remainder(index, hsub * wsub).remainder(wsub)
    x, y = [], []
    for ah in range(weights_subsampled):
        for aw in range(weights_subsampled):
            x.append(xsub * weights_subsampled + aw)
            y.append(ysub * weights_subsampled + ah)
    x, y = torch.cat(x), torch.cat(y)
    image_id = image_id.repeat(weights_subsampled ** 2)
    rgb = images[image_id, y, x] / 255.0
    c2w = camtoworlds[image_id]
    s = -1.0 if opengl else 1.0
    camera_dirs = torch.nn.functional.pad(
        torch.stack([(x - K[0, 2] + 0.5) / K[0, 0], (y - K[1, 2] + 0.5) / K[1, 1] * s], dim=-1), (0, 1), value=s)
    directions = (camera_dirs[:, None, :] * c2w[:, :3, :3]).sum(dim=-1)
    origins = torch.broadcast_to(c2w[:, :3, -1], directions.shape)
    viewdirs = directions / torch.linalg.norm(directions, dim=-1, keepdims=True)
    n = x.numel()
    return {"rgb": rgb.reshape(n, 3), "origins": origins.reshape(n, 3), "viewdirs": viewdirs.reshape(n, 3),
            "timestamps": timestamps[image_id], "idx": image_id, "cells": cells}
