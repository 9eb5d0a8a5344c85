"""PromptEncoder, B200-native.  Module tree / keys of sam3/sam3/sam/prompt_encoder.py (PromptEncoder :12-197,
PositionEmbeddingRandom :200-243).  Point prompts and the no-mask dense embedding run natively; box and mask
prompts are not on the hot path (config 3 = one point per image) and raise."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .common import LayerNorm2d


class PositionEmbeddingRandom(nn.Module):
    def __init__(self, num_pos_feats=64, scale=None):
        super().__init__()
        if scale is None or scale <= 0.0:
            scale = 1.0
        self.register_buffer("positional_encoding_gaussian_matrix", scale * torch.randn((2, num_pos_feats)))

    @torch.no_grad()
    def forward(self, size):
        """[C, h, w] fp32 (prompt_encoder.py:222-233)."""
        h, w = size
        tok = ops.dense_pe(self.positional_encoding_gaussian_matrix.float(), h, w)     # [h*w, C]
        return tok.view(h, w, -1).permute(2, 0, 1)

    def tokens(self, size):
        return ops.dense_pe(self.positional_encoding_gaussian_matrix.float(), size[0], size[1])


class PromptEncoder(nn.Module):
    def __init__(self, embed_dim, image_embedding_size, input_image_size, mask_in_chans, activation=nn.GELU):
        super().__init__()
        self.embed_dim = embed_dim
        self.input_image_size = input_image_size
        self.image_embedding_size = image_embedding_size
        self.pe_layer = PositionEmbeddingRandom(embed_dim // 2)
        self.num_point_embeddings = 4
        self.point_embeddings = nn.ModuleList([nn.Embedding(1, embed_dim) for _ in range(4)])
        self.not_a_point_embed = nn.Embedding(1, embed_dim)
        self.mask_input_size = (4 * image_embedding_size[0], 4 * image_embedding_size[1])
        self.mask_downscaling = nn.Sequential(      # parameters kept for checkpoint compatibility
            nn.Conv2d(1, mask_in_chans // 4, kernel_size=2, stride=2), LayerNorm2d(mask_in_chans // 4), activation(),
            nn.Conv2d(mask_in_chans // 4, mask_in_chans, kernel_size=2, stride=2), LayerNorm2d(mask_in_chans), activation(),
            nn.Conv2d(mask_in_chans, embed_dim, kernel_size=1))
        self.no_mask_embed = nn.Embedding(1, embed_dim)

    def get_dense_pe(self):
        return self.pe_layer(self.image_embedding_size).unsqueeze(0)

    @torch.no_grad()
    def forward(self, points, boxes, masks):
        if boxes is not None or masks is not None:
            raise NotImplementedError("native PromptEncoder: box / mask prompts are not built (point prompts only)")
        if points is None:
            raise NotImplementedError("native PromptEncoder: a point prompt is required")
        coords, labels = points
        table = torch.cat([e.weight for e in self.point_embeddings], dim=0).detach().float().contiguous()
        sparse = ops.point_embed(coords.float(), labels, self.pe_layer.positional_encoding_gaussian_matrix.float(),
                                 self.not_a_point_embed.weight.detach().float().reshape(-1), table,
                                 self.input_image_size[1], self.input_image_size[0])
        bs = coords.shape[0]
        dense = self.no_mask_embed.weight.detach().reshape(1, -1, 1, 1).expand(
            bs, -1, self.image_embedding_size[0], self.image_embedding_size[1])
        return sparse, dense
