"""Sparse (dual-octree) denoising U-Net: drop-in for reference
models/networks/diffusion_networks/graph_unet_hr.py `UNet3DModel` (same constructor arguments, module
tree and state_dict keys; forward :214-281).  The channel concatenations of the skip stack
(`torch.cat([h, hs.pop()], dim=1)`, :266) are never materialised: every consumer takes (x0 | x1).
"""
from __future__ import annotations
import torch
import torch.nn as nn

from . import ops
from .ops import PreparedWeight
from .modules import (GraphConv, GraphResBlockEmbed, GraphDownsample, GraphUpsample, graphnormalization,
                      zero_module, BatchedEmbedding)


class _Linear(nn.Linear):
    def prepared(self):
        if not hasattr(self, '_pw'):
            self._pw = PreparedWeight(1, self.in_features, 0, self.out_features)
        return self._pw.refresh(self.weight, 'linear')

    def run(self, x, a_silu=False):
        return ops.linear_small(x, self.weight, self.bias, a_silu=a_silu)


class UNet3DModel(nn.Module):
    def __init__(self, image_size, input_depth, full_depth, in_channels, model_channels, lr_model_channels,
                 out_channels, num_res_blocks, dropout=0, channel_mult=(1, 2, 4), dims=3, num_classes=None,
                 use_checkpoint=False, num_heads=-1, use_scale_shift_norm=False, **kwargs):
        super().__init__()
        self.image_size, self.input_depth, self.full_depth = image_size, input_depth, full_depth
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.num_res_blocks, self.dropout, self.channel_mult = num_res_blocks, dropout, list(channel_mult)
        self.num_classes, self.use_checkpoint, self.num_heads = num_classes, use_checkpoint, num_heads
        self.dtype = torch.float32
        n_edge_type, avg_degree = 7, 7
        time_embed_dim = model_channels * 4
        self.time_embed = nn.Sequential(_Linear(model_channels, time_embed_dim), nn.SiLU(),
                                        _Linear(time_embed_dim, time_embed_dim))
        if num_classes is not None:
            self.label_emb = nn.Embedding(num_classes, time_embed_dim)
        d = input_depth
        self.input_blocks = nn.ModuleList([GraphConv(in_channels, model_channels, n_edge_type, avg_degree, d - 1)])
        input_block_chans = [model_channels]
        ch = model_channels
        res = lambda cin, cout, dd: GraphResBlockEmbed(  # noqa: E731
            cin, time_embed_dim, dropout, out_channels=cout, n_edge_type=n_edge_type, avg_degree=avg_degree,
            n_node_type=dd - 1, dims=dims, use_checkpoint=use_checkpoint, use_scale_shift_norm=use_scale_shift_norm)
        for level, mult in enumerate(self.channel_mult):
            for _ in range(num_res_blocks[level]):
                self.input_blocks.append(res(ch, mult * model_channels, d))
                ch = mult * model_channels
                input_block_chans.append(ch)
            if level != len(self.channel_mult) - 1:
                d -= 1
                self.input_blocks.append(GraphDownsample(ch, ch, n_edge_type, avg_degree, d - 1))
                input_block_chans.append(ch)
        self.middle_block1 = res(ch, lr_model_channels, d)
        self.middle_block2 = res(lr_model_channels * 2, ch, d)
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(self.channel_mult))[::-1]:
            for i in range(num_res_blocks[level] + 1):
                ich = input_block_chans.pop()
                self.output_blocks.append(res(ch + ich, model_channels * mult, d))
                ch = model_channels * mult
                if level and i == num_res_blocks[level]:
                    d += 1
                    self.output_blocks.append(GraphUpsample(ch, ch, n_edge_type, avg_degree, d - 1))
        self.end_norm = graphnormalization(ch)
        self.end = nn.SiLU()
        self.out = zero_module(GraphConv(ch, out_channels, n_edge_type, avg_degree, input_depth - 1))

    def embed(self, timesteps, label):
        t = ops.timestep_embedding(timesteps, self.model_channels)
        emb = self.time_embed[2].run(self.time_embed[0].run(t), a_silu=True)
        if self.num_classes is not None:
            assert label is not None and label.shape == (timesteps.shape[0],)
            ops.embedding_add(emb, self.label_emb.weight, label.to(torch.int32).contiguous())
        return emb

    def forward_as_middle(self, h, doctree, timesteps, label, context):
        return self.forward(x=h, doctree=doctree, timesteps=timesteps, label=label, context=context, as_middle=True)

    @torch.no_grad()
    def forward(self, x=None, doctree=None, unet_lr=None, timesteps=None, label=None, context=None,
                as_middle=False, out_f32=True, **kwargs):
        assert (label is not None) == (self.num_classes is not None), \
            'must specify label if and only if the model is class-conditional'
        bsz = doctree.batch_size
        emb = self.embed(timesteps, label)
        if not hasattr(self, '_batched_emb'):
            blocks = [m for m in self.input_blocks if isinstance(m, GraphResBlockEmbed)]
            blocks += [self.middle_block1, self.middle_block2]
            blocks += [m for m in self.output_blocks if isinstance(m, GraphResBlockEmbed)]
            self._emb_blocks = blocks
            self._batched_emb = BatchedEmbedding([m.emb_layers[1] for m in blocks])
        es = {id(m): e for m, e in zip(self._emb_blocks, self._batched_emb(emb))}
        d = self.input_depth
        hs = []
        h = x.contiguous()
        if not as_middle:
            h = self.input_blocks[0].run(h, doctree.plan[d], stats=doctree.plan[d].stat)
        hs.append(h)
        for module in self.input_blocks[1:]:
            if isinstance(module, GraphResBlockEmbed):
                h = module.run(h, emb, doctree.plan[d], bsz, e=es[id(module)])
            elif isinstance(module, GraphDownsample):
                h = module(h, doctree, d)
                d -= 1
            else:
                h = module.run(h, doctree.plan[d], stats=doctree.plan[d].stat)
            hs.append(h)
        if unet_lr is not None:
            h = self.middle_block1.run(h, emb, doctree.plan[d], bsz, e=es[id(self.middle_block1)])
            h_lr = unet_lr.forward_as_middle(h, doctree, timesteps, label, context)
            h = self.middle_block2.run(h, emb, doctree.plan[d], bsz, x1=h_lr, e=es[id(self.middle_block2)])
        for module in self.output_blocks:
            if isinstance(module, GraphResBlockEmbed):
                h = module.run(h, emb, doctree.plan[d], bsz, x1=hs.pop(), e=es[id(module)])
            else:
                h = module(h, doctree, d)
                d += 1
        h = self.end_norm.run(h, doctree.plan[d], bsz, act=True)
        if as_middle:
            return h
        out = self.out.run(h, doctree.plan[d], out_f32=out_f32)
        assert out.shape[0] == x.shape[0]
        return out
