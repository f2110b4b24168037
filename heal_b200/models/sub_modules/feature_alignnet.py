"""AlignNet mirror (opencood/models/sub_modules/feature_alignnet.py:12-39): only `identity` (HEAL
base / m1) is on the hot path; the ConvNeXt/SDTA aligners are a 'next' row."""
import torch.nn as nn


class AlignNet(nn.Module):
    def __init__(self, args):
        super().__init__()
        model_name = (args or {}).get('core_method', 'identity')
        if model_name != 'identity':
            raise NotImplementedError(f"aligner '{model_name}' is out of the heal_b200 hot-path scope (identity only)")
        self.channel_align = nn.Identity()

    def forward(self, x):
        return self.channel_align(x)
