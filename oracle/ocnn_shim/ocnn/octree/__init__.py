"""Restated `ocnn.octree` subset (test infrastructure; see package docstring)."""
import torch

_BATCH_SHIFT = 48            # reference dual_octree.py:75  (key >> 48 is the batch id)
_KEY_MASK = (1 << 48) - 1


def xyz2key(x, y, z, b=None, depth: int = 16):
    """Morton interleave, x at bit 3i+2, y at 3i+1, z at 3i (child index = 4x+2y+z,
    as the lookup tables in reference dual_octree.py:90-112 assume); batch at bit 48
    (dual_octree.py:75,138)."""
    x = x.long(); y = y.long(); z = z.long()
    key = torch.zeros_like(x)
    for i in range(depth):
        key = key | (((x >> i) & 1) << (3 * i + 2)) | (((y >> i) & 1) << (3 * i + 1)) \
                  | (((z >> i) & 1) << (3 * i))
    if b is not None:
        b = b.long() if torch.is_tensor(b) else torch.as_tensor(b, dtype=torch.long)
        key = key | (b << _BATCH_SHIFT)
    return key


def key2xyz(key, depth: int = 16):
    """Inverse of xyz2key -> (x, y, z, b) (reference dual_octree.py:47-51,130)."""
    key = key.long()
    b = key >> _BATCH_SHIFT
    k = key & _KEY_MASK
    x = torch.zeros_like(k); y = torch.zeros_like(k); z = torch.zeros_like(k)
    for i in range(depth):
        x = x | (((k >> (3 * i + 2)) & 1) << i)
        y = y | (((k >> (3 * i + 1)) & 1) << i)
        z = z | (((k >> (3 * i)) & 1) << i)
    return x, y, z, b


class Octree:
    """Field layout the reference reads (dual_octree.py:28-44, util_dualoctree.py:232-248):
    keys[d] int64, children[d] int32 (-1 empty, else rank among non-empty), nnum[d],
    nnum_nempty[d]."""

    def __init__(self, depth, full_depth=2, batch_size=1, device='cpu', **kw):
        self.depth = depth
        self.full_depth = full_depth
        self.batch_size = batch_size
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        n = depth + 1
        self.keys = [None] * n
        self.children = [None] * n
        self.nnum = torch.zeros(n, dtype=torch.long)
        self.nnum_nempty = torch.zeros(n, dtype=torch.long)

    # -- growth (ldm_diffusion_util.py:318-325, util_dualoctree.py:238-248) --------
    def octree_grow_full(self, depth, update_neigh=False):
        num = 8 ** depth
        k = torch.arange(num, dtype=torch.long, device=self.device)
        b = torch.arange(self.batch_size, dtype=torch.long, device=self.device)
        self.keys[depth] = ((b.unsqueeze(1) << _BATCH_SHIFT) | k.unsqueeze(0)).reshape(-1)
        self.children[depth] = torch.arange(num * self.batch_size, dtype=torch.int32,
                                            device=self.device)
        self.nnum[depth] = num * self.batch_size
        self.nnum_nempty[depth] = num * self.batch_size

    def octree_split(self, split, depth):
        split = split.to(self.device).long()
        rank = torch.cumsum(split, 0) - 1
        self.children[depth] = torch.where(split > 0, rank, torch.full_like(rank, -1)).int()
        self.nnum_nempty[depth] = int(split.sum())

    def octree_grow(self, depth, update_neigh=False):
        mask = self.children[depth - 1] >= 0
        kp = self.keys[depth - 1][mask]
        bb = kp >> _BATCH_SHIFT
        kk = ((kp & _KEY_MASK) << 3).unsqueeze(1) + torch.arange(8, device=self.device)
        self.keys[depth] = ((bb.unsqueeze(1) << _BATCH_SHIFT) | kk).reshape(-1)
        n = self.keys[depth].numel()
        self.children[depth] = torch.arange(n, dtype=torch.int32, device=self.device)
        self.nnum[depth] = n
        self.nnum_nempty[depth] = n

    # -- accessors -------------------------------------------------------------------
    def nempty_mask(self, depth):
        return self.children[depth] >= 0

    def key(self, depth, nempty=False):
        k = self.keys[depth]
        return k[self.nempty_mask(depth)] if nempty else k

    def batch_id(self, depth, nempty=False):
        return self.key(depth, nempty) >> _BATCH_SHIFT

    def xyzb(self, depth, nempty=False):
        return key2xyz(self.key(depth, nempty), depth)

    def search_key(self, query, depth, nempty=False):
        """index of each query key among the nodes of `depth` (keys are sorted: batch-major Morton order), -1 where
        absent -- the lookup reference mpu.py:72 relies on."""
        keys = self.key(depth, nempty)
        pos = torch.searchsorted(keys, query.long()).clamp(max=keys.numel() - 1)
        return torch.where(keys[pos] == query.long(), pos, torch.full_like(pos, -1))

    def to(self, device):
        device = torch.device(device)
        self.device = device
        self.keys = [k.to(device) if k is not None else None for k in self.keys]
        self.children = [c.to(device) if c is not None else None for c in self.children]
        return self

    def cuda(self):
        return self.to('cuda')

    def cpu(self):
        return self.to('cpu')


class Points:  # only referenced by data loading code that is out of scope
    def __init__(self, *a, **k):
        raise NotImplementedError('ocnn shim: Points is outside the hot path')


def merge_octrees(*a, **k):
    raise NotImplementedError('ocnn shim: merge_octrees is outside the hot path')
