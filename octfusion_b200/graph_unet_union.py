"""Stage router: drop-in for reference models/networks/diffusion_networks/graph_unet_union.py
`UNet3DModel` (:11-92): builds the lr / hr nets from the stage-list yaml and dispatches by `unet_type`."""
from __future__ import annotations
import torch.nn as nn

from . import graph_unet_hr, graph_unet_lr


class UNet3DModel(nn.Module):
    def __init__(self, stage_flag, image_size, input_depth, unet_type, full_depth, input_channels, out_channels,
                 model_channels, num_res_blocks, attention_resolutions, channel_mult, num_heads, use_checkpoint,
                 dims, num_classes=None, **kwargs):
        super().__init__()
        self.unet_lr = self.unet_hr = self.unet_feature = None
        for i in range(len(unet_type)):
            if unet_type[i] == 'lr':
                self.unet_lr = graph_unet_lr.UNet3DModel(
                    full_depth=full_depth, in_split_channels=input_channels[i], model_channels=model_channels[i],
                    out_split_channels=out_channels[i], attention_resolutions=attention_resolutions,
                    channel_mult=channel_mult[i], use_checkpoint=use_checkpoint, num_heads=num_heads, dims=dims,
                    num_classes=num_classes)
            elif unet_type[i] in ('hr', 'feature'):
                m = graph_unet_hr.UNet3DModel(
                    image_size=image_size[i], input_depth=input_depth[i], full_depth=full_depth,
                    in_channels=input_channels[i], model_channels=model_channels[i],
                    lr_model_channels=model_channels[i - 1], out_channels=out_channels[i],
                    num_res_blocks=num_res_blocks[i], channel_mult=channel_mult[i], dims=dims,
                    use_checkpoint=use_checkpoint, num_heads=num_heads, num_classes=num_classes)
                if unet_type[i] == 'hr':
                    self.unet_hr = m
                else:
                    self.unet_feature = m
            else:
                raise ValueError(unet_type[i])
            if unet_type[i] == stage_flag:
                break

    def forward(self, unet_type=None, **input_data):
        """Inference semantics of reference :80-92 (the training-time random self-conditioning draw of the
        'lr' branch, :82-85, is a training feature and is not reproduced)."""
        if unet_type == 'lr':
            return self.unet_lr(**input_data)
        if unet_type == 'hr':
            return self.unet_hr(**input_data)
        if unet_type == 'feature':
            return self.unet_feature(**input_data)
        raise ValueError(unet_type)
