import torch.nn as nn

from compressai.entropy_models import EntropyBottleneck


class CompressionModel(nn.Module):
    def __init__(self, entropy_bottleneck_channels=None, init_weights=None):
        super().__init__()
        if entropy_bottleneck_channels is not None:
            self.entropy_bottleneck = EntropyBottleneck(entropy_bottleneck_channels)

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck) and hasattr(m, "loss"))

    def update(self, scale_table=None, force=False):
        updated = False
        for m in self.modules():
            if isinstance(m, EntropyBottleneck):
                updated |= bool(m.update(force=force))
        return updated
