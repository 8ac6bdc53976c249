"""Input encoders -- mirror of the reference's layers/input_encoder.py (outside the hot path, plain PyTorch)."""
import torch
import torch.nn as nn


class EmbeddingEncoder(nn.Module):
    def __init__(self, input_size, hidden_size):
        super(EmbeddingEncoder, self).__init__()
        self.init_proj = nn.Embedding(input_size, hidden_size)

    def reset_parameters(self):
        self.init_proj.reset_parameters()

    def forward(self, data):
        return self.init_proj(data.x)


class LinearEncoder(nn.Module):
    def __init__(self, input_size, hidden_size):
        super(LinearEncoder, self).__init__()
        self.init_proj = nn.Linear(input_size, hidden_size)

    def reset_parameters(self):
        self.init_proj.reset_parameters()

    def forward(self, data):
        return self.init_proj(data.x)


class QM9InputEncoder(nn.Module):
    """19 (or 22 with 3-D positions) continuous features + an 8-wide atomic-number embedding."""

    def __init__(self, hidden_size, use_pos=False):
        super(QM9InputEncoder, self).__init__()
        self.use_pos = use_pos
        self.init_proj = nn.Linear(22 if use_pos else 19, hidden_size)
        self.z_embedding = nn.Embedding(1000, 8)

    def reset_parameters(self):
        self.init_proj.reset_parameters()
        self.z_embedding.reset_parameters()

    def forward(self, data):
        x, z = data.x, data.z
        z_emb = 0
        if z is not None:
            z_emb = self.z_embedding(z)
            if z_emb.ndim == 3:
                z_emb = z_emb.sum(dim=1)
        x = torch.cat([z_emb, x], -1)
        if self.use_pos:
            x = torch.cat([x, data.pos], 1)
        return self.init_proj(x)
