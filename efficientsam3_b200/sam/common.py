"""Parameter containers mirroring sam3/sam3/sam/common.py (MLPBlock :8-24, LayerNorm2d :27-39)."""
import torch
import torch.nn as nn


class MLPBlock(nn.Module):
    def __init__(self, embedding_dim, mlp_dim, act=nn.GELU):
        super().__init__()
        self.lin1 = nn.Linear(embedding_dim, mlp_dim)
        self.lin2 = nn.Linear(mlp_dim, embedding_dim)
        self.act_name = {nn.ReLU: "relu", nn.GELU: "gelu"}[act]


class LayerNorm2d(nn.Module):
    def __init__(self, num_channels, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps
