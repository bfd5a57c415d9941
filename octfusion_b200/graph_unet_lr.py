"""Dense 16^3 U-Net with attention: drop-in for reference
models/networks/diffusion_networks/graph_unet_lr.py `UNet3DModel` (same constructor arguments, module
tree and state_dict keys).  Used (a) as the middle block of the sparse U-Net
(`forward_as_middle`, :175-182) and (b) stand-alone as the stage-1 split generator (`forward`, :184-230).

Layout: every dense tensor is [B * 8^r, C] with rows in Morton order.  On the full octree layer the
octree key *is* the Morton index (reference dual_octree.py:137-138), so `octree2voxel` and the gather back
(:176-181) are the identity here -- no permute, no scatter.
"""
from __future__ import annotations
import math
import torch
import torch.nn as nn

from . import ops
from .modules import (ResnetBlock, ConvDownsample, ConvUpsample, NormActAttention, LearnedSinusoidalPosEmb,
                      convnormalization, activation_function, our_Identity, conv_nd, DenseTables,
                      _to_morton, _from_morton, BatchedEmbedding)
from .graph_unet_hr import _Linear


class UNet3DModel(nn.Module):
    def __init__(self, full_depth, in_split_channels, model_channels, out_split_channels, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), dims=3, num_classes=None, use_checkpoint=False,
                 num_heads=-1, use_text_condition=False, context_dim=None, n_embed=None, **kwargs):
        super().__init__()
        assert dims == 3 and not use_text_condition
        self.full_depth = full_depth
        self.in_channels, self.model_channels, self.out_channels = in_split_channels, model_channels, out_split_channels
        self.attention_resolutions, self.dropout, self.channel_mult = attention_resolutions, dropout, list(channel_mult)
        self.num_classes, self.use_checkpoint, self.num_heads = num_classes, use_checkpoint, num_heads
        self.dtype = torch.float32
        channels = [model_channels, *[model_channels * m for m in self.channel_mult]]
        in_out = list(zip(channels[:-1], channels[1:]))
        time_embed_dim = model_channels * 4
        self.time_pos_emb = LearnedSinusoidalPosEmb(model_channels)
        self.time_emb = nn.Sequential(_Linear(model_channels + 1, time_embed_dim), activation_function(),
                                      _Linear(time_embed_dim, time_embed_dim))
        if num_classes is not None:
            self.label_emb = nn.Embedding(num_classes, time_embed_dim)
        self.input_emb = conv_nd(dims, 2 * self.in_channels, model_channels, 3, padding=1)
        self.downs, self.ups = nn.ModuleList([]), nn.ModuleList([])
        nres = len(in_out)
        ds = 1
        for ind, (dim_in, dim_out) in enumerate(in_out):
            is_last = ind >= nres - 1
            self.downs.append(nn.ModuleList([
                ResnetBlock(dims, dim_in, dim_out, emb_dim=time_embed_dim, dropout=dropout),
                NormActAttention(dim_out, num_heads) if ds in attention_resolutions else our_Identity(),
                ConvDownsample(dim_out, dims=dims) if not is_last else our_Identity()]))
            if not is_last:
                ds *= 2
        mid_dim = channels[-1]
        self.mid_block1 = ResnetBlock(dims, mid_dim, mid_dim, emb_dim=time_embed_dim, dropout=dropout)
        self.mid_self_attn = NormActAttention(mid_dim, num_heads) if ds in attention_resolutions else our_Identity()
        self.mid_block2 = ResnetBlock(dims, mid_dim, mid_dim, emb_dim=time_embed_dim, dropout=dropout)
        for ind, (dim_in, dim_out) in enumerate(reversed(in_out[1:])):
            is_last = ind >= nres - 1
            self.ups.append(nn.ModuleList([
                ResnetBlock(dims, dim_out * 2, dim_in, emb_dim=time_embed_dim, dropout=dropout),
                NormActAttention(dim_in, num_heads) if ds in attention_resolutions else our_Identity(),
                ConvUpsample(dim_in, dims=dims) if not is_last else our_Identity()]))
            if not is_last:
                ds //= 2
        self.end = nn.Sequential(convnormalization(model_channels), activation_function())
        self.out = conv_nd(dims, model_channels, self.out_channels, 3, padding=1)
        self._tables = {}

    def tables(self, batch, device):
        k = (batch, str(device))
        if k not in self._tables:
            self._tables[k] = DenseTables(batch, device)
        return self._tables[k]

    def embed(self, timesteps, label):
        e = ops.learned_sinusoidal(timesteps, self.time_pos_emb.weights)
        emb = self.time_emb[2].run(self.time_emb[0].run(e), a_silu=True)
        if self.num_classes is not None:
            assert label is not None and label.shape == (timesteps.shape[0],)
            ops.embedding_add(emb, self.label_emb.weight, label.to(torch.int32).contiguous())
        return emb

    def run(self, x, timesteps, label, batch, res_log2, as_middle):
        """x [B * 8^r, C] Morton rows; returns the same layout."""
        t = self.tables(batch, x.device)
        r = res_log2
        emb = self.embed(timesteps, label)
        if not hasattr(self, '_batched_emb'):
            blocks = [m[0] for m in self.downs] + [self.mid_block1, self.mid_block2] + [m[0] for m in self.ups]
            self._emb_blocks = blocks
            self._batched_emb = BatchedEmbedding([m.time_mlp[1] for m in blocks])
        es = {id(m): e for m, e in zip(self._emb_blocks, self._batched_emb(emb))}
        skips = []
        for resnet, attn, down in self.downs:
            x = resnet.run(x, emb, t, r, e=es[id(resnet)])
            if isinstance(attn, NormActAttention):
                x = attn.run(x, t, r)
            skips.append(x)
            if isinstance(down, ConvDownsample):
                x = down.run(x, t, r)
                r -= 1
        x = self.mid_block1.run(x, emb, t, r, e=es[id(self.mid_block1)])
        if isinstance(self.mid_self_attn, NormActAttention):
            x = self.mid_self_attn.run(x, t, r)
        x = self.mid_block2.run(x, emb, t, r, e=es[id(self.mid_block2)])
        for resnet, attn, up in self.ups:
            x = resnet.run(x, emb, t, r, x1=skips.pop(), e=es[id(resnet)])
            if isinstance(attn, NormActAttention):
                x = attn.run(x, t, r)
            if isinstance(up, ConvUpsample):
                x = up.run(x, t, r)
                r += 1
        x = self.end[0].run(x, t, r, act=True)
        if as_middle:
            return x
        return self.out.run(x, t.conv(r))

    @torch.no_grad()
    def forward_as_middle(self, h, doctree, timesteps, label, context):
        """h: features of the full octree layer [B * 8^full_depth, C] in octree order = Morton order."""
        assert h.shape[0] == doctree.batch_size * 8 ** self.full_depth
        return self.run(h.contiguous(), timesteps, label, doctree.batch_size, self.full_depth, True)

    @torch.no_grad()
    def forward(self, x=None, timesteps=None, x_self_cond=None, label=None, context=None, as_middle=False,
                **kwargs):
        """x [B, C, D, H, W] as in the reference (boundary conversion to Morton rows and back)."""
        assert (label is not None) == (self.num_classes is not None), \
            'must specify label if and only if the model is class-conditional'
        b = x.shape[0]
        t = self.tables(b, x.device)
        if not as_middle:
            sc = torch.zeros_like(x) if x_self_cond is None else x_self_cond
            xm, r = _to_morton(x, t)
            sm, _ = _to_morton(sc, t)
            xm = self.input_emb.run(xm, t.conv(r), x1=sm, stats=t.stat_plan(r))
        else:
            xm, r = _to_morton(x, t)
        y = self.run(xm, timesteps, label, b, r, as_middle)
        return _from_morton(y, t, b, r)
