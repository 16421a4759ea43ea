"""Sharding of an image batch over the GPUs of one box.

Images are independent, so the data path has NO collective: rank r of W encodes/decodes the images
``shard_indices(n, r, W)`` on its own GPU with its own context. The only thing that crosses ranks is the
per-image blob size (8 bytes per image), gathered on the host side so that any rank can lay out the blobs
of the whole batch (``gather_sizes``). Works with any torch.distributed backend (gloo on CPU in the tests,
nccl under torchrun on the GPU box).
"""
import numpy as np


def shard_indices(n_images, rank, world):
    """Round-robin: image i belongs to rank i % world (equal work per rank up to one image)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_images, world))


def owner_of(image, world):
    return image % world


def gather_sizes(local_sizes, n_images, rank, world, dist=None, device="cpu"):
    """All ranks receive the blob size of every image of the batch, in image order.

    local_sizes[k] is the size of image shard_indices(n_images, rank, world)[k]. One all_reduce of an
    int64 vector in which every rank fills only its own slots (sum == concatenation)."""
    mine = shard_indices(n_images, rank, world)
    if len(local_sizes) != len(mine):
        raise ValueError("one size per local image expected")
    full = np.zeros(n_images, np.int64)
    full[mine] = np.asarray(local_sizes, np.int64)
    if world == 1 or dist is None:
        return full
    import torch
    t = torch.from_numpy(full).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def blob_offsets(sizes, align=1):
    """Exclusive scan of (aligned) sizes: where each image's blob starts if the batch is stored back to back."""
    sizes = np.asarray(sizes, np.int64)
    padded = (sizes + align - 1) // align * align
    return np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64)


def encode_sharded(encode_one, images, rank, world, dist=None, device="cpu"):
    """Encode the local shard with `encode_one(image) -> bytes` and return (local blobs, sizes of ALL images)."""
    mine = shard_indices(len(images), rank, world)
    blobs = [encode_one(images[i]) for i in mine]
    sizes = gather_sizes([len(b) for b in blobs], len(images), rank, world, dist, device)
    return dict(zip(mine, blobs)), sizes
