"""Integer-feature encoders -- mirror of the reference's layers/feature_encoder.py (dense embedding + Linear,
called once per forward by models/GNNs.py:172-179; outside the K-hop hot path, plain PyTorch)."""
import torch
import torch.nn as nn


def _embeddings(feature_dims, hidden_size, padding):
    kw = {"padding_idx": 0} if padding else {}
    return nn.ModuleList(nn.Embedding(dim, hidden_size, **kw) for dim in feature_dims)


class FeatureSumEncoder(nn.Module):
    """Sum of one embedding per integer feature column."""

    def __init__(self, feature_dims, hidden_size, padding=False):
        super(FeatureSumEncoder, self).__init__()
        self.embedding_list = _embeddings(feature_dims, hidden_size, padding)

    def reset_parameters(self):
        for emb in self.embedding_list:
            emb.reset_parameters()

    def forward(self, x):
        return sum(emb(x[..., i]) for i, emb in enumerate(self.embedding_list) if i < x.shape[-1])


class FeatureConcatEncoder(nn.Module):
    """Concatenation of one embedding per integer feature column, then a Linear back to hidden_size."""

    def __init__(self, feature_dims, hidden_size, padding=False):
        super(FeatureConcatEncoder, self).__init__()
        self.embedding_list = _embeddings(feature_dims, hidden_size, padding)
        self.proj = nn.Linear(len(feature_dims) * hidden_size, hidden_size)

    def reset_parameters(self):
        for emb in self.embedding_list:
            emb.reset_parameters()
        self.proj.reset_parameters()

    def forward(self, x):
        cols = [self.embedding_list[i](x[..., i]) for i in range(x.shape[-1])]
        return self.proj(torch.cat(cols, dim=-1))
