"""Host-side mirror of models/texture.py for the render path.
    VolumeRadiance ............ models/texture.py:15-41 (`volume-radiance`, the neus config)
"""
import torch
import torch.nn as nn

from .network_utils import Config, get_activation, get_encoding, get_mlp, update_module_step


class VolumeRadiance(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config = Config(config)
        self.n_dir_dims = config.get("n_dir_dims", 3)
        self.n_output_dims = 3
        self.encoding = get_encoding(self.n_dir_dims, config.dir_encoding_config)
        self.n_input_dims = config.input_feature_dim + self.encoding.n_output_dims
        self.network = get_mlp(self.n_input_dims, self.n_output_dims, config.mlp_network_config)

    def forward(self, features, dirs, *args):
        dirs = (dirs + 1.0) / 2.0
        dirs_embd = self.encoding(dirs.view(-1, self.n_dir_dims))
        network_inp = torch.cat([features.view(-1, features.shape[-1]), dirs_embd]
                                + [arg.view(-1, arg.shape[-1]) for arg in args], dim=-1)
        color = self.network(network_inp).view(*features.shape[:-1], self.n_output_dims).float()
        if "color_activation" in self.config:
            color = get_activation(self.config.color_activation)(color)
        return color

    def update_step(self, epoch, global_step):
        update_module_step(self.encoding, epoch, global_step)

    def regularizations(self, out):
        return {}
