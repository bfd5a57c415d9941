"""TEST INFRASTRUCTURE -- CPU restatement (torch, fp32 or fp64) of the OctFusion denoising
U-Net hot path.  It is the checker for the CUDA kernels; it is never shipped, never timed as
the product and never imported by `octfusion_b200`.

Pinning: `tests/test_oracle_vs_reference.py` runs every function here against the
UNMODIFIED reference modules imported from /root/reference (oracle/ref_import.py) on seeded
inputs, and `oracle/gen_golden.py` stores reference outputs under tests/golden/ so that the
same check runs on the GPU box where the reference tree does not exist.
The only unpinned part is the third-party `ocnn` octree container (see ocnn_shim).

Everything is functional: parameters come from a flat state_dict with the reference's own
key names, so the same dict drives the reference, this oracle and the product.
All citations are file:line under /root/reference.
"""
from __future__ import annotations
import math
from types import SimpleNamespace
import numpy as np
import torch
import torch.nn.functional as F

N_DIR = 7          # 6 face directions + self loop (dual_octree.py:85-89,247)


# ---------------------------------------------------------------------------------------
# elementary operators
# ---------------------------------------------------------------------------------------
def silu(x):
    return x * torch.sigmoid(x)


def group_count(channels: int, group: int = 32) -> int:
    """Group rule of DualOctreeGroupNorm.__init__ (models/networks/modules.py:271-274)."""
    if channels <= 32:
        return channels // 4
    if channels % group != 0:
        return 30
    return group


def graph_conv(x, graph, weights, n_node_type: int, bias=None):
    """GraphConv.forward (models/networks/modules.py:194-220) with scatter_mean
    (diffusion_networks/utils/scatter.py:42-66): per (row, dir) slot the MEAN of the neighbour
    features (with the neighbours' one-hot node type appended when n_node_type > 1), empty
    slots stay zero, then one GEMM with weights [7*(Cin+nt), Cout]."""
    row, col = graph['edge_idx'][0], graph['edge_idx'][1]
    edir = graph['edge_dir']
    n = x.shape[0]
    if n_node_type > 1 and graph.get('node_type') is not None:
        onehot = F.one_hot(graph['node_type'].long(), n_node_type).to(x.dtype)
        x = torch.cat([x, onehot], 1)
    slot = row * N_DIR + edir
    acc = torch.zeros(n * N_DIR, x.shape[1], dtype=x.dtype).index_add_(0, slot, x[col])
    cnt = torch.zeros(n * N_DIR, dtype=x.dtype).index_add_(
        0, slot, torch.ones(slot.shape[0], dtype=x.dtype)).clamp_(min=1)
    out = (acc / cnt.unsqueeze(1)).reshape(n, -1) @ weights
    if bias is not None:
        out = out + bias
    return out


def doctree_group_norm(x, batch_id, batch_size, gamma, beta, eps=1e-5):
    """DualOctreeGroupNorm.forward (models/networks/modules.py:291-326): statistics per
    (sample, group) over that sample's nodes; eps is added to the element COUNT (:302) and to
    the variance (:310); two-pass variance."""
    c = x.shape[1]
    g = group_count(c)
    cpg = c // g
    cnt = torch.bincount(batch_id, minlength=batch_size).to(x.dtype) * cpg
    inv = 1.0 / (cnt + eps)
    s = torch.zeros(batch_size, c, dtype=x.dtype).index_add_(0, batch_id, x)
    mean = s.reshape(batch_size, g, cpg).sum(-1) * inv.unsqueeze(1)              # [B, G]
    xc = x - mean.repeat_interleave(cpg, 1)[batch_id]
    v = torch.zeros(batch_size, c, dtype=x.dtype).index_add_(0, batch_id, xc * xc)
    var = v.reshape(batch_size, g, cpg).sum(-1) * inv.unsqueeze(1)
    rstd = 1.0 / torch.sqrt(var + eps)
    return xc * rstd.repeat_interleave(cpg, 1)[batch_id] * gamma.reshape(1, -1) + beta.reshape(1, -1)


def downsample(x, w):
    """Downsample.forward (modules.py:392-395): flat-view arithmetic, w is [C, C, 8]."""
    c = w.shape[0]
    return x.reshape(-1, 8 * c) @ w.flatten(1).t()


def upsample(x, w):
    """Upsample.forward (modules.py:440-443)."""
    c = w.shape[0]
    return (x @ w.flatten(1)).reshape(-1, c)


def graph_downsample(x, doctree, d, w_down, w_conv, n_node_type):
    """GraphDownsample.forward (modules.py:409-428): depth-d graph features -> depth d-1."""
    numd = int(doctree.nnum[d])
    lnumd = int(doctree.lnum[d - 1])
    leaf = doctree.node_child(d - 1) < 0
    pooled = downsample(x[x.shape[0] - numd:], w_down)
    out = torch.zeros(leaf.shape[0], x.shape[1], dtype=x.dtype)
    out[leaf] = x[x.shape[0] - lnumd - numd: x.shape[0] - numd]
    out[~leaf] = pooled
    out = torch.cat([x[: x.shape[0] - numd - lnumd], out], 0)
    return graph_conv(out, doctree.graph[d - 1], w_conv, n_node_type)


def graph_upsample(x, doctree, d, w_up, w_conv, n_node_type):
    """GraphUpsample.forward (modules.py:458-472): depth-d graph features -> depth d+1."""
    numd = int(doctree.nnum[d])
    leaf = doctree.node_child(d) < 0
    outd = x[x.shape[0] - numd:]
    up = upsample(outd[~leaf], w_up)
    out = torch.cat([x[: x.shape[0] - numd], outd[leaf], up], 0)
    return graph_conv(out, doctree.graph[d + 1], w_conv, n_node_type)


def res_block_embed(x, emb, doctree, d, sd, prefix, n_node_type):
    """GraphResBlockEmbed._forward (modules.py:741-763), additive-shift branch (the
    scale-shift branch :747-751 is dead code).  The per-sample Python loop :757-758 is the
    broadcast add below."""
    bid = doctree.batch_id(d)
    bsz = doctree.batch_size
    g = doctree.graph[d]
    h = doctree_group_norm(x, bid, bsz, sd[prefix + 'block1_norm.weights'], sd[prefix + 'block1_norm.bias'])
    h = graph_conv(silu(h), g, sd[prefix + 'conv1.weights'], n_node_type)
    e = F.linear(silu(emb), sd[prefix + 'emb_layers.1.weight'], sd[prefix + 'emb_layers.1.bias'])
    h = h + e[bid]
    h = doctree_group_norm(h, bid, bsz, sd[prefix + 'block2_norm.weights'], sd[prefix + 'block2_norm.bias'])
    h = graph_conv(silu(h), g, sd[prefix + 'conv2.weights'], n_node_type)
    key = prefix + 'skip_connection.linear.weight'
    skip = x @ sd[key].t() if key in sd else x            # Conv1x1 has no bias (modules.py:728,334)
    return skip + h


def timestep_embedding(t, dim, max_period=10000):
    """ldm_diffusion_util.py:171-191 (cos first, then sin)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    e = torch.cat([torch.cos(args), torch.sin(args)], -1)
    if dim % 2:
        e = torch.cat([e, torch.zeros_like(e[:, :1])], -1)
    return e


def learned_sinusoidal(t, w):
    """LearnedSinusoidalPosEmb.forward (modules.py:558-563): [t, sin(2 pi t w), cos(2 pi t w)]."""
    f = t[:, None] * w[None, :] * (2 * math.pi)
    return torch.cat([t[:, None], f.sin(), f.cos()], -1)


# ---------------------------------------------------------------------------------------
# dense LR middle U-Net (graph_unet_lr.py) -- channels-first [B, C, D, H, W] as the reference
# ---------------------------------------------------------------------------------------
def group_norm32(x, w, b):
    """GroupNorm32 (modules.py:26-28) built by convnormalization (:34-36): min(C,32) groups."""
    xf = x if x.dtype == torch.float64 else x.float()      # fp64 only for oracle-precision studies
    return F.group_norm(xf, min(x.shape[1], 32), w.to(xf.dtype), b.to(xf.dtype), 1e-5).to(x.dtype)


def qkv_attention(qkv):
    """QKVAttention.forward (modules.py:538-547): legacy head-major q|k|v split, both q and k
    scaled by ch^-1/4, softmax over keys in fp32."""
    ch = qkv.shape[1] // 3
    q, k, v = qkv[:, :ch], qkv[:, ch:2 * ch], qkv[:, 2 * ch:]
    s = 1.0 / math.sqrt(math.sqrt(ch))
    w = torch.einsum('bct,bcs->bts', q * s, k * s)
    w = torch.softmax(w.float(), -1).to(w.dtype)
    return torch.einsum('bts,bcs->bct', w, v)


def attention_block(x, sd, prefix, heads):
    """AttentionBlock.forward (modules.py:527-535)."""
    b, c = x.shape[:2]
    xf = x.reshape(b, c, -1)
    h = group_norm32(xf, sd[prefix + 'norm.weight'], sd[prefix + 'norm.bias'])
    qkv = F.conv1d(h, sd[prefix + 'qkv.weight'], sd[prefix + 'qkv.bias'])
    a = qkv_attention(qkv.reshape(b * heads, -1, qkv.shape[2])).reshape(b, -1, qkv.shape[2])
    a = F.conv1d(a, sd[prefix + 'proj_out.weight'], sd[prefix + 'proj_out.bias'])
    return (xf + a).reshape(x.shape)


def norm_act_attention(x, sd, prefix, heads):
    """nn.Sequential(convnormalization, SiLU, AttentionBlock) (graph_unet_lr.py:128-132)."""
    h = silu(group_norm32(x, sd[prefix + '0.weight'], sd[prefix + '0.bias']))
    return attention_block(h, sd, prefix + '2.', heads)


def resnet_block(x, emb, sd, prefix):
    """ResnetBlock.forward (modules.py:504-513) with use_text_condition=False; eval mode so
    Dropout is the identity."""
    h = silu(group_norm32(x, sd[prefix + 'block1.0.weight'], sd[prefix + 'block1.0.bias']))
    h = F.conv3d(h, sd[prefix + 'block1.2.weight'], sd[prefix + 'block1.2.bias'], padding=1)
    t = F.linear(silu(emb), sd[prefix + 'time_mlp.1.weight'], sd[prefix + 'time_mlp.1.bias'])
    h = h + t[:, :, None, None, None]
    h = silu(group_norm32(h, sd[prefix + 'block2.0.weight'], sd[prefix + 'block2.0.bias']))
    h = F.conv3d(h, sd[prefix + 'block2.3.weight'], sd[prefix + 'block2.3.bias'], padding=1)
    if prefix + 'res_conv.weight' in sd:
        x = F.conv3d(x, sd[prefix + 'res_conv.weight'], sd[prefix + 'res_conv.bias'])
    return h + x


def lr_forward_dense(x, timesteps, sd, cfg, prefix='unet_lr.', label=None, as_middle=True):
    """graph_unet_lr.UNet3DModel.forward (graph_unet_lr.py:184-230); cfg carries
    channel_mult / attention_resolutions / num_heads of the LR entry of the yaml."""
    heads = cfg['num_heads']
    mults = cfg['channel_mult']
    att = cfg['attention_resolutions']
    nres = len(mults)
    if not as_middle:
        x = torch.cat([x, torch.zeros_like(x)], 1)                      # x_self_cond default (:198-199)
        x = F.conv3d(x, sd[prefix + 'input_emb.weight'], sd[prefix + 'input_emb.bias'], padding=1)
    emb = learned_sinusoidal(timesteps.float(), sd[prefix + 'time_pos_emb.weights'])
    emb = F.linear(emb, sd[prefix + 'time_emb.0.weight'], sd[prefix + 'time_emb.0.bias'])
    emb = F.linear(silu(emb), sd[prefix + 'time_emb.2.weight'], sd[prefix + 'time_emb.2.bias'])
    if label is not None:
        emb = emb + sd[prefix + 'label_emb.weight'][label]
    skips = []
    ds = 1
    for i in range(nres):
        p = f'{prefix}downs.{i}.'
        x = resnet_block(x, emb, sd, p + '0.')
        if ds in att:
            x = norm_act_attention(x, sd, p + '1.', heads)
        skips.append(x)
        if i < nres - 1:
            x = F.conv3d(x, sd[p + '2.op.weight'], sd[p + '2.op.bias'], stride=2, padding=1)
            ds *= 2
    x = resnet_block(x, emb, sd, prefix + 'mid_block1.')
    if ds in att:
        x = norm_act_attention(x, sd, prefix + 'mid_self_attn.', heads)
    x = resnet_block(x, emb, sd, prefix + 'mid_block2.')
    for i in range(nres - 1):
        p = f'{prefix}ups.{i}.'
        x = torch.cat([x, skips.pop()], 1)
        x = resnet_block(x, emb, sd, p + '0.')
        if ds in att:
            x = norm_act_attention(x, sd, p + '1.', heads)
        x = F.interpolate(x, scale_factor=2, mode='nearest')            # ConvUpsample (:63-77)
        x = F.conv3d(x, sd[p + '2.conv.weight'], sd[p + '2.conv.bias'], padding=1)
        ds //= 2
    x = silu(group_norm32(x, sd[prefix + 'end.0.weight'], sd[prefix + 'end.0.bias']))
    if as_middle:
        return x
    return F.conv3d(x, sd[prefix + 'out.weight'], sd[prefix + 'out.bias'], padding=1)


def lr_forward_as_middle(h, doctree, timesteps, sd, cfg, prefix='unet_lr.', label=None):
    """graph_unet_lr.UNet3DModel.forward_as_middle (graph_unet_lr.py:175-182): scatter the
    full-layer node features into a [B,C,16,16,16] voxel grid (ocnn.nn.octree2voxel), run the
    dense net, gather back in node order."""
    fd = cfg['full_depth']
    x, y, z, b = doctree.octree.xyzb(fd)
    s = 2 ** fd
    vox = torch.zeros(doctree.batch_size, s, s, s, h.shape[1], dtype=h.dtype)
    vox[b, x, y, z] = h
    out = lr_forward_dense(vox.permute(0, 4, 1, 2, 3).contiguous(), timesteps, sd, cfg, prefix,
                           label, as_middle=True)
    return out.permute(0, 2, 3, 4, 1)[b, x, y, z]


# ---------------------------------------------------------------------------------------
# sparse HR U-Net (graph_unet_hr.py)
# ---------------------------------------------------------------------------------------
def hr_layout(cfg):
    """Module sequence of graph_unet_hr.UNet3DModel.__init__ (graph_unet_hr.py:116-209) as a
    list of (kind, prefix, depth-at-entry, n_node_type)."""
    d = cfg['input_depth']
    seq_in = [('conv', 'input_blocks.0.', d, d - 1)]
    idx = 1
    nlev = len(cfg['channel_mult'])
    for level in range(nlev):
        for _ in range(cfg['num_res_blocks'][level]):
            seq_in.append(('res', f'input_blocks.{idx}.', d, d - 1)); idx += 1
        if level != nlev - 1:
            d -= 1
            seq_in.append(('down', f'input_blocks.{idx}.', d + 1, d - 1)); idx += 1
    mid_depth = d
    seq_out = []
    idx = 0
    for level in reversed(range(nlev)):
        for i in range(cfg['num_res_blocks'][level] + 1):
            seq_out.append(('res', f'output_blocks.{idx}.', d, d - 1)); idx += 1
            if level and i == cfg['num_res_blocks'][level]:
                d += 1
                seq_out.append(('up', f'output_blocks.{idx}.', d - 1, d - 1)); idx += 1
    return seq_in, mid_depth, seq_out


def hr_forward(x, doctree, timesteps, sd, cfg_hr, cfg_lr=None, label=None,
               prefix='unet_hr.', lr_prefix='unet_lr.'):
    """graph_unet_hr.UNet3DModel.forward (graph_unet_hr.py:214-281) with `unet_lr` given as
    (sd, cfg_lr).  `timesteps` are log-SNR floats (octfusion_model_union.py:315-322)."""
    sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    mc = cfg_hr['model_channels']
    emb = timestep_embedding(timesteps, mc).to(x.dtype)
    emb = F.linear(emb, sub['time_embed.0.weight'], sub['time_embed.0.bias'])
    emb = F.linear(silu(emb), sub['time_embed.2.weight'], sub['time_embed.2.bias'])
    if label is not None:
        emb = emb + sub['label_emb.weight'][label]
    seq_in, mid_d, seq_out = hr_layout(cfg_hr)
    hs = []
    h = x
    for kind, p, d, nt in seq_in:
        if kind == 'conv':
            h = graph_conv(h, doctree.graph[d], sub[p + 'weights'], nt)
        elif kind == 'res':
            h = res_block_embed(h, emb, doctree, d, sub, p, nt)
        else:
            h = graph_downsample(h, doctree, d, sub[p + 'downsample.weights'], sub[p + 'conv.weights'], nt)
        hs.append(h)
    d = mid_d
    if cfg_lr is not None:
        h = res_block_embed(h, emb, doctree, d, sub, 'middle_block1.', d - 1)
        h_lr = lr_forward_as_middle(h, doctree, timesteps, sd, cfg_lr, lr_prefix, label)
        h = torch.cat([h, h_lr], 1)
        h = res_block_embed(h, emb, doctree, d, sub, 'middle_block2.', d - 1)
    for kind, p, d, nt in seq_out:
        if kind == 'res':
            h = torch.cat([h, hs.pop()], 1)
            h = res_block_embed(h, emb, doctree, d, sub, p, nt)
        else:
            h = graph_upsample(h, doctree, d, sub[p + 'upsample.weights'], sub[p + 'conv.weights'], nt)
            d = d + 1
    dlast = cfg_hr['input_depth']
    h = silu(doctree_group_norm(h, doctree.batch_id(dlast), doctree.batch_size,
                                sub['end_norm.weights'], sub['end_norm.bias']))
    return graph_conv(h, doctree.graph[dlast], sub['out.weights'], dlast - 1)


def split_cfg(unet_params: dict):
    """Index the stage-list yaml the way graph_unet_union.UNet3DModel.__init__ does
    (graph_unet_union.py:39-77): entry i-1 = 'lr', entry i = 'hr'."""
    types = unet_params['unet_type']
    il, ih = types.index('lr'), types.index('hr')
    common = dict(full_depth=unet_params['full_depth'], num_heads=unet_params['num_heads'],
                  attention_resolutions=unet_params['attention_resolutions'],
                  num_classes=unet_params.get('num_classes'))
    lr = dict(common, model_channels=unet_params['model_channels'][il],
              channel_mult=unet_params['channel_mult'][il],
              in_channels=unet_params['input_channels'][il], out_channels=unet_params['out_channels'][il])
    hr = dict(common, model_channels=unet_params['model_channels'][ih],
              lr_model_channels=unet_params['model_channels'][ih - 1],
              channel_mult=unet_params['channel_mult'][ih], num_res_blocks=unet_params['num_res_blocks'][ih],
              input_depth=unet_params['input_depth'][ih],
              in_channels=unet_params['input_channels'][ih], out_channels=unet_params['out_channels'][ih])
    return lr, hr


# ---------------------------------------------------------------------------------------
# sampler arithmetic (octfusion_model_union.py:300-352, ldm_diffusion_util.py:300-309)
# ---------------------------------------------------------------------------------------
def beta_linear_log_snr(t):
    return -torch.log(torch.special.expm1(1e-4 + 10 * (t ** 2)))


def ddim_eps_update(x, eps, log_snr, log_snr_next):
    """'eps' branch of sample_loop (octfusion_model_union.py:345-350); scalars per step."""
    alpha, sigma = torch.sqrt(torch.sigmoid(log_snr)), torch.sqrt(torch.sigmoid(-log_snr))
    alpha_n, sigma_n = torch.sqrt(torch.sigmoid(log_snr_next)), torch.sqrt(torch.sigmoid(-log_snr_next))
    x0 = (x - eps * sigma) / alpha.clamp(min=1e-8)
    return x0 * alpha_n + eps * sigma_n


def ddpm_x0_update(x, pred, log_snr, log_snr_next, noise, do_sign):
    """'x0' branch of sample_loop (octfusion_model_union.py:324-344); returns (x_next, pred_used)."""
    if do_sign:
        pred = torch.sign(pred)
    alpha = torch.sqrt(torch.sigmoid(log_snr))
    alpha_n, sigma_n = torch.sqrt(torch.sigmoid(log_snr_next)), torch.sqrt(torch.sigmoid(-log_snr_next))
    c = -torch.expm1(log_snr - log_snr_next)
    mean = alpha_n * (x * (1 - c) / alpha + c * pred)
    var = (sigma_n ** 2) * c
    return mean + torch.sqrt(var) * (noise if noise is not None else torch.zeros_like(x)), pred


# ---------------------------------------------------------------------------------------
# dual-octree graph (dual_octree.py) -- restated GEOMETRICALLY: two graph nodes are joined in
# direction `dir` when their cells share a face in that direction.  The reference reaches the
# same edge set hierarchically (dense_graph :124-155, sparse_graph :195-239 with its child
# tables :90-112); tests compare the two edge sets exactly.
# ---------------------------------------------------------------------------------------
_NGH = np.array([[0, 0, 1], [0, 0, -1], [0, 1, 0], [0, -1, 0], [1, 0, 0], [-1, 0, 0]], np.int64)
_OPP = np.array([1, 0, 3, 2, 5, 4], np.int64)                      # dual_octree.py:98-100


# ---------------------------------------------------------------------------------------
# NeuralMPU: models/networks/dualoctree_networks/mpu.py -- oracle only so far
# ---------------------------------------------------------------------------------------
def mpu_eval(pos, reg_voxs, octree, full_depth, depth_stop, depth):
    """NeuralMPU.__call__ (mpu.py:143-155) restated per query point instead of through sparse matrices.
    pos [P, 4] = (x, y, z in [-1, 1], batch index).  For every depth d in [full_depth, D] the 8 cells of depth d whose
    centres surround the point contribute  w * (F . [offset, 1])  with  w = prod(1 - |offset in cells|) * d^2 / 50
    (mpu.py:88-93), F = reg_voxs[D][node] (4 values: gradient + value), offset rescaled to the [-1, 1] frame (:97);
    a cell contributes only if it exists, and for d < D only if it is a leaf (:118-121).  Output for D in
    [depth_stop, depth]: (sum / (sum of weights + 1e-8), point touched by a depth-D cell) (:135-140)."""
    xyz, bid = pos[:, :3], pos[:, 3].long()
    p = pos.shape[0]
    corner = torch.tensor([[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)], dtype=pos.dtype)   # mpu.py:37-41
    nnum_cum = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(octree.nnum, 0)])
    per_depth = {}
    for d in range(full_depth, depth + 1):
        scale = 2 ** d
        xf = (xyz + 1.0) * (scale / 2.0) - 0.5
        xi = torch.floor(xf)
        cells = xi.unsqueeze(1) + corner                                   # [P, 8, 3]
        off = xf.unsqueeze(1) - cells                                      # in [-1, 1]
        inb = ((cells > -1) & (cells < scale)).all(-1)
        c = cells.clamp(0, scale - 1).long()
        key = _xyz_key_t(c[..., 0], c[..., 1], c[..., 2], bid.unsqueeze(1).expand(-1, 8), d)
        keys = octree.keys[d]
        at = torch.searchsorted(keys, key.reshape(-1)).clamp(max=keys.numel() - 1).reshape(p, 8)
        found = (keys[at] == key) & inb
        w = (1.0 - off.abs()).prod(-1) * (d ** 2 / 50.0)
        per_depth[d] = (at, found, w, off * (2.0 / scale))
    out = {}
    for D in range(depth_stop, depth + 1):
        num = torch.zeros(p, dtype=pos.dtype)
        den = torch.zeros(p, dtype=pos.dtype)
        for d in range(full_depth, D + 1):
            at, found, w, off = per_depth[d]
            use = found if d == D else found & (octree.children[d][at] < 0)
            f = reg_voxs[D][at + int(nnum_cum[d] - nnum_cum[full_depth])]   # [P, 8, 4]
            val = (f[..., :3] * off).sum(-1) + f[..., 3]
            num = num + torch.where(use, w * val, torch.zeros_like(w)).sum(1)
            den = den + torch.where(use, w, torch.zeros_like(w)).sum(1)
        out[D] = (num / (den + 1e-8), per_depth[D][1].any(1))
    return out


def _xyz_key_t(x, y, z, b, depth):
    k = torch.zeros_like(x)
    for i in range(depth):
        k = k | (((x >> i) & 1) << (3 * i + 2)) | (((y >> i) & 1) << (3 * i + 1)) | (((z >> i) & 1) << (3 * i))
    return k | (b << 48)


# ---------------------------------------------------------------------------------------
# GraphVAE decoder: models/networks/dualoctree_networks/graph_vae.py
# ---------------------------------------------------------------------------------------
VAE_CHANNELS = [4, 512, 512, 256, 128, 64, 32, 32, 24, 8]      # graph_vae.py:125, channels[depth]


def conv1x1_gn(x, doctree, d, sd, prefix, act=None):
    """Conv1x1Gn / Conv1x1GnGelu / Conv1x1GnGeluSequential (modules.py:343-381): bias-free Linear ->
    DualOctreeGroupNorm -> optional exact (erf) GELU."""
    h = x @ sd[prefix + 'conv.linear.weight'].t()
    h = doctree_group_norm(h, doctree.batch_id(d), doctree.batch_size, sd[prefix + 'gn.weights'], sd[prefix + 'gn.bias'])
    return F.gelu(h) if act == 'gelu' else h


def graph_res_block(x, doctree, d, sd, prefix, n_node_type):
    """GraphResBlock._forward (modules.py:630-648): GN -> swish -> conv1 -> GN -> swish -> dropout(0) -> conv2,
    skip = Conv1x1Gn when the channel count changes."""
    bid, bsz, g = doctree.batch_id(d), doctree.batch_size, doctree.graph[d]
    h = doctree_group_norm(x, bid, bsz, sd[prefix + 'norm1.weights'], sd[prefix + 'norm1.bias'])
    h = graph_conv(silu(h), g, sd[prefix + 'conv1.weights'], n_node_type)
    h = doctree_group_norm(h, bid, bsz, sd[prefix + 'norm2.weights'], sd[prefix + 'norm2.bias'])
    h = graph_conv(silu(h), g, sd[prefix + 'conv2.weights'], n_node_type)
    if prefix + 'conv1x1c.conv.linear.weight' in sd:
        x = conv1x1_gn(x, doctree, d, sd, prefix + 'conv1x1c.')
    return h + x


def graph_res_blocks(x, doctree, d, sd, prefix, num, n_node_type):
    """GraphResBlocks.forward (modules.py:651-666)."""
    for i in range(num):
        x = graph_res_block(x, doctree, d, sd, prefix + 'resblks.%d.' % i, n_node_type)
    return x


def vae_graph_upsample(x, doctree, d, sd, prefix):
    """dualoctree_networks/modules.py:73-91 (the VAE's GraphUpsample: no graph conv): features of the depth-(d-1)
    graph -> depth-d graph; Conv1x1GnGelu when the channel count changes."""
    numd = int(doctree.nnum[d - 1])
    leaf = doctree.node_child(d - 1) < 0
    outd = x[x.shape[0] - numd:]
    out = torch.cat([x[: x.shape[0] - numd], outd[leaf], upsample(outd[~leaf], sd[prefix + 'upsample.weights'])], 0)
    if prefix + 'conv1x1.conv.linear.weight' in sd:
        out = conv1x1_gn(out, doctree, d, sd, prefix + 'conv1x1.', 'gelu')
    return out


def vae_graph_downsample(x, doctree, d, sd, prefix):
    """dualoctree_networks/modules.py:39-66: depth-(d+1) graph features -> depth-d graph (call with the TARGET depth d,
    as graph_vae.py:155 does)."""
    numd, lnumd = int(doctree.nnum[d + 1]), int(doctree.lnum[d])
    leaf = doctree.node_child(d) < 0
    pooled = downsample(x[x.shape[0] - numd:], sd[prefix + 'downsample.weights'])
    out = torch.zeros(leaf.shape[0], x.shape[1], dtype=x.dtype)
    out[leaf] = x[x.shape[0] - lnumd - numd: x.shape[0] - numd]
    out[~leaf] = pooled
    out = torch.cat([x[: x.shape[0] - numd - lnumd], out], 0)
    if prefix + 'conv1x1.conv.linear.weight' in sd:
        out = conv1x1_gn(out, doctree, d, sd, prefix + 'conv1x1.', 'gelu')
    return out


def _vae_head(h, doctree, d, sd, prefix):
    """_make_predict_module (graph_vae.py:127-130): Conv1x1GnGeluSequential(C, 32) -> Conv1x1(32, out, bias)."""
    t = conv1x1_gn(h, doctree, d, sd, prefix + '0.', 'gelu')
    return t @ sd[prefix + '1.linear.weight'].t() + sd[prefix + '1.linear.bias']


def vae_encode(data, doctree, sd, depth, depth_stop, resblk_num):
    """GraphVAE.octree_encoder_step + KL_conv (graph_vae.py:135-168) on given input features `data` [N_depth, 4]:
    returns the posterior moments (mean | logvar) on the depth_stop graph."""
    h = graph_conv(data, doctree.graph[depth], sd['conv1.weights'], depth - 1)
    for i, d in enumerate(range(depth, depth_stop - 1, -1)):
        h = graph_res_blocks(h, doctree, d, sd, 'encoder.%d.' % i, resblk_num - 1, d - 1)
        if d > depth_stop:
            h = vae_graph_downsample(h, doctree, d - 1, sd, 'downsample.%d.' % i)
    h = doctree_group_norm(h, doctree.batch_id(depth_stop), doctree.batch_size, sd['encoder_norm_out.weights'],
                           sd['encoder_norm_out.bias'])
    return F.gelu(h) @ sd['KL_conv.linear.weight'].t() + sd['KL_conv.linear.bias']


def vae_decode(code, doctree, sd, depth_stop, depth_out, resblk_num, update_octree=False, labels=None,
               make_doctree=None):
    """GraphVAE.octree_decoder (graph_vae.py:171-223).  `labels` (dict depth -> int tensor) overrides the argmax of
    the split logits when the octree is grown, so that two implementations can be compared on identical octrees
    even where a logit pair is a near-tie.  Returns (logits, reg_voxs, octree)."""
    ch = VAE_CHANNELS
    make_doctree = make_doctree or DualGraph
    h = code @ sd['post_KL_conv.linear.weight'].t() + sd['post_KL_conv.linear.bias']
    h = graph_res_blocks(h, doctree, depth_stop, sd, 'decoder_mid.block_1.', resblk_num, depth_stop - 1)
    h = graph_res_blocks(h, doctree, depth_stop, sd, 'decoder_mid.block_2.', resblk_num, depth_stop - 1)
    logits, regs = {}, {}
    for i, d in enumerate(range(depth_stop, depth_out + 1)):
        if d > depth_stop:
            h = vae_graph_upsample(h, doctree, d, sd, 'upsample.%d.' % (i - 1))
        assert h.shape[1] == ch[d]
        h = graph_res_blocks(h, doctree, d, sd, 'decoder.%d.' % i, resblk_num, d - 1)
        logit = _vae_head(h, doctree, d, sd, 'predict.%d.' % i)
        nnum = int(doctree.nnum[d])
        logits[d] = logit[logit.shape[0] - nnum:]
        if update_octree:
            label = (labels[d] if labels is not None else logits[d].argmax(1)).to(torch.int32)
            octree = doctree.octree
            octree.octree_split(label, d)
            if d < depth_out:
                octree.octree_grow(d + 1)
                octree.depth += 1
            doctree = make_doctree(octree)
        reg = _vae_head(h, doctree, d, sd, 'regress.%d.' % i)
        mask = doctree.graph[d]['node_mask']
        pad = torch.zeros(mask.shape[0], reg.shape[1], dtype=reg.dtype)
        pad[mask] = reg
        regs[d] = pad
    return logits, regs, doctree.octree


def _facing_children(direction):
    """child octants (4x+2y+z) of a cell that touch its face looking in `direction`."""
    axis = {0: 0, 1: 0, 2: 1, 3: 1, 4: 2, 5: 2}[int(direction)]      # bit position: z=0,y=1,x=2
    want = 1 if direction in (0, 2, 4) else 0
    return np.array([c for c in range(8) if ((c >> axis) & 1) == want], np.int64)


def _key_xyz(key, depth):
    k = key & ((1 << 48) - 1)
    x = np.zeros_like(k); y = np.zeros_like(k); z = np.zeros_like(k)
    for i in range(depth):
        x |= ((k >> (3 * i + 2)) & 1) << i
        y |= ((k >> (3 * i + 1)) & 1) << i
        z |= ((k >> (3 * i)) & 1) << i
    return x, y, z, key >> 48


def _xyz_key(x, y, z, b, depth):
    k = np.zeros_like(x)
    for i in range(depth):
        k |= (((x >> i) & 1) << (3 * i + 2)) | (((y >> i) & 1) << (3 * i + 1)) | (((z >> i) & 1) << (3 * i))
    return k | (b << 48)


# ---------------------------------------------------------------------------------------
# ocnn.nn.OctreeConv (third-party; NOT called by the reference -- SURVEY.md section 0 / Appendix B last row).
# Restated from ocnn-pytorch 2.2.x as recalled: PARITY UNPINNED at the ocnn boundary (no source under /root/reference,
# no reference test fixes the tap order or the weight layout).  Independent anchor used by the tests: on a full octree
# layer it must equal torch.nn.functional.conv3d with zero padding on the [b, x, y, z] voxel grid.
# ---------------------------------------------------------------------------------------
def octree_neigh27(octree, depth, stride=1, nempty=False):
    """`octree.get_neigh(depth, '333', stride, nempty)`: int64 [N', 27]; entry (dx+1)*9 + (dy+1)*3 + (dz+1) = index
    (within `depth`) of the node at (x+dx, y+dy, z+dz), -1 where the cell is outside the volume or absent."""
    from ocnn.octree import key2xyz, xyz2key
    keys = octree.keys[depth].long()
    x, y, z, b = key2xyz(keys, depth)
    n = keys.shape[0]
    out = torch.full((n, 27), -1, dtype=torch.long)
    lim = 1 << depth
    for t in range(27):
        dx, dy, dz = t // 9 - 1, (t // 3) % 3 - 1, t % 3 - 1
        nx, ny, nz = x + dx, y + dy, z + dz
        ok = (nx >= 0) & (ny >= 0) & (nz >= 0) & (nx < lim) & (ny < lim) & (nz < lim)
        k = xyz2key(nx.clamp(0, lim - 1), ny.clamp(0, lim - 1), nz.clamp(0, lim - 1), b, depth)
        pos = torch.searchsorted(keys, k).clamp(max=n - 1)
        hit = ok & (keys[pos] == k)
        out[hit, t] = pos[hit]
    if nempty:
        child = octree.children[depth].long()
        mapped = torch.where(out >= 0, child[out.clamp(min=0)], torch.full_like(out, -1))
        out = mapped[child >= 0]
    if stride == 2:
        out = out[::8]
    return out


def octree_conv(data, octree, depth, weights, stride=1, nempty=False, bias=None):
    """out = gather(data, neigh).flatten(1) @ weights.flatten(0, 1) (+ bias); weights [27, Cin, Cout]."""
    neigh = octree_neigh27(octree, depth, stride, nempty)
    buf = torch.zeros(neigh.shape[0], 27, data.shape[1], dtype=data.dtype)
    valid = neigh >= 0
    buf[valid] = data[neigh[valid]]
    out = buf.flatten(1) @ weights.flatten(0, 1)
    return out + bias if bias is not None else out


class DualGraph:
    """Oracle counterpart of reference `DualOctree` + `post_processing_for_docnn`
    (dual_octree.py:19-63,400-409), exposing the duck-typed surface the U-Net reads
    (SURVEY.md 8b): graph[d]{edge_idx,edge_dir,node_type}, batch_id(d), batch_size, nnum,
    lnum, node_child(d), octree."""

    def __init__(self, octree):
        self.octree = octree
        self.depth, self.full_depth, self.batch_size = octree.depth, octree.full_depth, octree.batch_size
        self.nnum = octree.nnum.clone()
        self.lnum = octree.nnum - octree.nnum_nempty
        fd, dep = self.full_depth, self.depth
        keys = {d: octree.keys[d].cpu().numpy().astype(np.int64) for d in range(fd, dep + 1)}
        child = {d: octree.children[d].cpu().numpy().astype(np.int64) for d in range(fd, dep + 1)}
        self._child = {d: octree.children[d].cpu() for d in range(fd, dep + 1)}
        xyzb = {d: _key_xyz(keys[d], d) for d in range(fd, dep + 1)}
        self.graph = [dict() for _ in range(dep + 1)]
        self._bid = {}
        for D in range(fd, dep + 1):
            # graph index of every octree node that is a graph node at depth D
            gidx, off = {}, 0
            for d in range(fd, D + 1):
                isnode = (child[d] < 0) if d < D else np.ones(len(child[d]), bool)
                gi = np.full(len(child[d]), -1, np.int64)
                gi[isnode] = off + np.arange(int(isnode.sum()))
                off += int(isnode.sum())
                gidx[d] = gi
            rows, cols, dirs = [], [], []
            for d in range(fd, D + 1):
                src = np.nonzero(gidx[d] >= 0)[0]
                x, y, z, b = (a[src] for a in xyzb[d])
                for dr in range(6):
                    nx, ny, nz = x + _NGH[dr, 0], y + _NGH[dr, 1], z + _NGH[dr, 2]
                    bound = 1 << d
                    ok = (nx >= 0) & (nx < bound) & (ny >= 0) & (ny < bound) & (nz >= 0) & (nz < bound)
                    s, nx, ny, nz, nb = src[ok], nx[ok], ny[ok], nz[ok], b[ok]
                    # climb: find the existing cell that contains the neighbour position
                    found_d = np.full(len(s), -1, np.int64)
                    found_i = np.full(len(s), -1, np.int64)
                    for dd in range(d, fd - 1, -1):
                        sh = d - dd
                        k = _xyz_key(nx >> sh, ny >> sh, nz >> sh, nb, dd)
                        pos = np.searchsorted(keys[dd], k)
                        pos = np.minimum(pos, len(keys[dd]) - 1)
                        hit = (keys[dd][pos] == k) & (found_d < 0)
                        found_d[hit] = dd
                        found_i[hit] = pos[hit]
                    assert (found_d >= 0).all()
                    # descend: a found cell that is subdivided contributes its facing descendants
                    cur_s, cur_d, cur_i = s, found_d, found_i
                    face = _facing_children(_OPP[dr])
                    while len(cur_s):
                        isnode = np.zeros(len(cur_s), bool)
                        for dd in range(fd, D + 1):
                            m = cur_d == dd
                            isnode[m] = gidx[dd][cur_i[m]] >= 0
                        for dd in range(fd, D + 1):
                            m = isnode & (cur_d == dd)
                            if m.any():
                                rows.append(gidx[d][cur_s[m]]); cols.append(gidx[dd][cur_i[m]])
                                dirs.append(np.full(int(m.sum()), dr, np.int64))
                        rest = ~isnode
                        if not rest.any():
                            break
                        ns, nd, ni = [], [], []
                        for dd in range(fd, D):
                            m = rest & (cur_d == dd)
                            if m.any():
                                base = child[dd][cur_i[m]] * 8
                                ni.append((base[:, None] + face[None, :]).reshape(-1))
                                ns.append(np.repeat(cur_s[m], 4)); nd.append(np.full(4 * int(m.sum()), dd + 1, np.int64))
                        cur_s, cur_d, cur_i = np.concatenate(ns), np.concatenate(nd), np.concatenate(ni)
            ntot = off
            rows.append(np.arange(ntot)); cols.append(np.arange(ntot)); dirs.append(np.full(ntot, 6, np.int64))
            row, col, edir = np.concatenate(rows), np.concatenate(cols), np.concatenate(dirs)
            order = np.lexsort((col, row * N_DIR + edir))
            ntype = np.concatenate([np.full(int((child[d] < 0).sum()) if d < D else len(child[d]), d - fd, np.int64)
                                    for d in range(fd, D + 1)])
            bid = np.concatenate([(xyzb[d][3][child[d] < 0] if d < D else xyzb[d][3]) for d in range(fd, D + 1)])
            # add_node_mask (dual_octree.py:391-398): over ALL octree nodes of depths fd..D, True = graph node
            nmask = np.concatenate([(child[d] < 0) if d < D else np.ones(len(child[d]), bool) for d in range(fd, D + 1)])
            self.graph[D] = {'edge_idx': torch.from_numpy(np.stack([row[order], col[order]])),
                             'edge_dir': torch.from_numpy(edir[order]),
                             'node_type': torch.from_numpy(ntype),
                             'node_mask': torch.from_numpy(nmask)}
            self._bid[D] = torch.from_numpy(bid)
        self.total_num = int(self._bid[dep].shape[0])

    def batch_id(self, depth, nempty=False):
        return self._bid[depth]

    def node_child(self, depth):
        return self._child[depth]


def edge_set(graph):
    """canonical (row*7+dir, col) sorted edge list for set comparison."""
    r, c = graph['edge_idx'][0].long(), graph['edge_idx'][1].long()
    k = (r * N_DIR + graph['edge_dir'].long())
    key = k * (int(c.max()) + 1 if c.numel() else 1) + c
    order = torch.argsort(key)
    return k[order], c[order]


# ---------------------------------------------------------------------------------------
# deterministic parameters shared by reference / oracle / product
# ---------------------------------------------------------------------------------------
def seeded_state_dict(shapes: dict, seed: int = 0, dtype=torch.float32):
    """Every tensor ~ N(0, s^2) with s chosen per kind so that activations stay O(1) through
    the net (the reference's zero-initialised conv2/out/proj_out tensors -- modules.py:719,525,
    499, graph_unet_hr.py:209 -- would make the whole U-Net output identically 0 and hide
    errors).  Norm scales are 1 + 0.1 N(0,1).  One Generator per key (seeded by a stable hash
    of the key) so that a sub-dict draws the same numbers as the full dict."""
    import zlib
    out = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + 7919 * seed) % (2 ** 31))
        r = torch.randn(shp, generator=g, dtype=torch.float32)
        leaf = k.split('.')[-1]
        if _is_norm_key(k):
            t = 1.0 + 0.1 * r if leaf in ('weights', 'weight') else 0.05 * r
        elif 'time_pos_emb' in k:
            t = r                                         # randn in the reference (modules.py:556)
        elif 'label_emb' in k:
            t = 0.5 * r
        elif len(shp) <= 1:
            t = 0.05 * r                                  # biases
        else:
            if leaf == 'weights' and len(shp) == 2:       # GraphConv [7(Cin+nt), Cout]
                fan_in = shp[0]
            elif leaf == 'weights' and 'upsample' in k:   # Upsample [C, C, 8]: x @ W.flatten(1)
                fan_in = shp[0]
            else:                                         # Linear / ConvNd / Downsample
                fan_in = int(np.prod(shp[1:]))
            t = r / math.sqrt(max(fan_in, 1))
        out[k] = t.to(dtype)
    return out


def _is_norm_key(k: str) -> bool:
    parts = k.split('.')
    if any('norm' in p for p in parts) or (len(parts) >= 2 and parts[-2] == 'gn'):
        return True
    # nn.Sequential(GroupNorm32, SiLU, conv) members: block1.0 / block2.0 / end.0 / <attn seq>.0
    if len(parts) >= 2 and parts[-2] == '0' and parts[-1] in ('weight', 'bias'):
        return parts[-3] in ('block1', 'block2', 'end', '1', 'mid_self_attn') if len(parts) >= 3 else False
    return False
